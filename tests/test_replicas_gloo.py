"""N > 1 host path on CPU: two gloo ranks shard a batch of scenes, run the (oracle) decode on their slice and
rank 0 gathers the per-image tables in image order -- must equal the single-process result exactly."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lightweight_openpose_b200 import replicas, synth  # noqa: E402

PARAM = {"thre1": 0.1, "thre2": 0.0}
N_SCENES = 5          # odd on purpose: ranks get 3 and 2 images


def _scenes():
    return [synth.synth_scene(900 + i, 1 + i % 3, 128, 128)[:2] for i in range(N_SCENES)]


def _decode_many(batch):
    from oracle import decode_oracle as orc          # CPU stand-in for Engine.infer_host in this test
    return [orc.decode_detail((128, 128), PARAM, h, p) for h, p in batch]


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = replicas.run_sharded(_scenes(), _decode_many)
        if rank == 0:
            np.save(out_path, np.array([r["joint_list"].shape[0] for r in res] +
                                       [np.asarray(r["persons"]).shape[0] for r in res]))
            np.save(out_path + ".jl.npy", np.concatenate([r["joint_list"].reshape(-1, 6) for r in res]))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch_in_order():
    for n in (0, 1, 5, 256, 257):
        for world in (1, 2, 4, 8):
            parts = [replicas.shard_bounds(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_match_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    single = _decode_many(_scenes())
    want = np.array([r["joint_list"].shape[0] for r in single] + [np.asarray(r["persons"]).shape[0] for r in single])
    assert np.array_equal(np.load(out), want)
    assert np.array_equal(np.load(out + ".jl.npy"), np.concatenate([r["joint_list"].reshape(-1, 6) for r in single]))


def _ring_worker(rank, world, port, name, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ring = replicas.SharedResultRing(name, N_SCENES, world, rank, slots=2, pin=False)
        lo, hi = replicas.shard_bounds(N_SCENES, world, rank)
        steps = 5                                   # more steps than slots: the slot hand-back is exercised
        got = []
        for step in range(steps):
            scenes = [synth.synth_scene(900 + step * 10 + i, 1 + i % 3, 128, 128)[:2] for i in range(N_SCENES)]
            ring.wait_slot_free(step)
            replicas.pack_host_results(_decode_many(scenes[lo:hi]), ring._tables[step % ring.slots][rank])
            ring.publish(step)
            if rank == 0:
                views = ring.gather(step)
                res = [r for v in views for r in replicas.batch_result_from_tables(v)]
                got.append(np.concatenate([r.joint_list.reshape(-1, 6) for r in res] +
                                          [np.asarray(r.person_to_joint_assoc).reshape(-1, 16)[:, :6] for r in res]))
                assert len(res) == N_SCENES
                ring.release(step)
        if rank == 0:
            np.save(out_path, np.concatenate(got))
        dist.barrier()
        ring.close()
    finally:
        dist.destroy_process_group()


def test_shared_memory_result_ring_two_ranks(tmp_path):
    """The host-side gather of the batch-sharded job: two ranks write their packed result tables into one shared-memory
    ring, rank 0 reads all of them in image order, 5 steps through 2 slots."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "ring.npy")
    mp.spawn(_ring_worker, args=(2, port, f"lwop_test_ring_{port}", out), nprocs=2, join=True)
    want = []
    for step in range(5):
        scenes = [synth.synth_scene(900 + step * 10 + i, 1 + i % 3, 128, 128)[:2] for i in range(N_SCENES)]
        res = _decode_many(scenes)
        want.append(np.concatenate([r["joint_list"].reshape(-1, 6) for r in res] +
                                   [np.asarray(r["persons"]).reshape(-1, 16)[:, :6] for r in res]))
    assert np.array_equal(np.load(out), np.concatenate(want))
