"""Batch-sharded path on the GPU: the slices of one batch run through separate engines, their result
tables land in the shared-memory ring by DMA, and what rank 0 gathers must equal the unsplit batch bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lightweight_openpose_b200 import checkpoint, engine, replicas, synth  # noqa: E402

pytestmark = pytest.mark.gpu
H = W = 368
B = 7             # odd on purpose: the two slices hold 4 and 3 frames


def _frames():
    return np.stack([synth.synth_image(40 + i, H, W) for i in range(B)])


def _same(a, b):
    assert np.array_equal(np.asarray(a.joint_list), np.asarray(b.joint_list))
    assert np.array_equal(np.asarray(a.person_to_joint_assoc), np.asarray(b.person_to_joint_assoc))
    assert np.array_equal(a.joints, b.joints)
    for la, lb in zip(a.limbs, b.limbs):
        assert np.array_equal(la, lb)


def test_two_way_split_through_the_ring_equals_the_unsplit_batch():
    variables, src = checkpoint.load_or_synthesize()
    frames = _frames()
    full = engine.Engine(H, W, max_batch=B)
    full.load_variables(variables, src)
    want = list(full.infer_host(torch.from_numpy(frames).pin_memory()))
    assert sum(r.counts[0] for r in want) > 0
    name = f"lwop_test_gpu_ring_{os.getpid()}"
    rings = [replicas.SharedResultRing(name, B, 2, r, slots=2) for r in range(2)]          # rank 0 first: it creates
    pools = []
    for r in range(2):
        lo, hi = replicas.shard_bounds(B, 2, r)
        p = engine.EnginePool(H, W, max_batch=hi - lo, lanes=2, depth=2)
        p.load_variables(variables, src)
        pools.append((p, torch.from_numpy(frames[lo:hi].copy()).pin_memory()))
    try:
        for step in range(5):                              # more steps than ring slots and than lanes
            for r, (p, x) in enumerate(pools):
                rings[r].wait_slot_free(step)
                p.submit(x, tables=rings[r].tables(step))
            for r, (p, _) in enumerate(pools):
                p.result()
                rings[r].publish(step)
            views = rings[0].gather(step)
            got = [res for v in views for res in replicas.batch_result_from_tables(v)]
            assert len(got) == B
            for a, b in zip(got, want):
                _same(a, b)
            rings[0].release(step)
    finally:
        for p, _ in pools:
            p.close()
        for ring in reversed(rings):
            ring.close()
        full.close()


def _rank(rank, world, name, out_dir):
    torch.cuda.set_device(rank)
    variables, src = checkpoint.load_or_synthesize()
    frames = _frames()
    lo, hi = replicas.shard_bounds(B, world, rank)
    ring = replicas.SharedResultRing(name, B, world, rank, slots=2)
    pool = engine.EnginePool(H, W, max_batch=hi - lo, lanes=2, device=rank)
    pool.load_variables(variables, src)
    x = torch.from_numpy(frames[lo:hi].copy()).pin_memory()
    for step in range(3):
        ring.wait_slot_free(step)
        pool.submit(x, tables=ring.tables(step))
        pool.result()
        ring.publish(step)
        if rank == 0:
            views = ring.gather(step)
            got = [res for v in views for res in replicas.batch_result_from_tables(v)]
            np.save(os.path.join(out_dir, f"jl_{step}.npy"), np.concatenate([np.asarray(g.joint_list).reshape(-1, 6) for g in got]))
            np.save(os.path.join(out_dir, f"kp_{step}.npy"), np.concatenate([g.joints.reshape(-1, 42) for g in got]))
            ring.release(step)
    if rank != 0:                                  # rank 0 unlinks the segment: leave only after it has read everything
        ring.wait_slot_free(3 + ring.slots - 1)
    pool.close()
    ring.close()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_one_process_each(tmp_path):
    import torch.multiprocessing as mp
    variables, src = checkpoint.load_or_synthesize()
    full = engine.Engine(H, W, max_batch=B)
    full.load_variables(variables, src)
    want = list(full.infer_host(torch.from_numpy(_frames()).pin_memory()))
    full.close()
    mp.spawn(_rank, args=(2, f"lwop_test_gpu_ring2_{os.getpid()}", str(tmp_path)), nprocs=2, join=True)
    for step in range(3):
        assert np.array_equal(np.load(tmp_path / f"jl_{step}.npy"),
                              np.concatenate([np.asarray(r.joint_list).reshape(-1, 6) for r in want]))
        assert np.array_equal(np.load(tmp_path / f"kp_{step}.npy"), np.concatenate([r.joints.reshape(-1, 42) for r in want]))
