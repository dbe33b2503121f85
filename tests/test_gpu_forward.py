"""Whole-network parity on the GPU against the CPU oracle (oracle/forward_oracle.py):
heatmaps and PAFs within max-abs 1e-4 in the fp32 check mode and 2e-2 in bf16
(BASELINE.json north_star), at 368x368, 656x368 and 256x256."""
import numpy as np
import pytest
import torch

from lightweight_openpose_b200 import checkpoint, engine, synth
from lightweight_openpose_b200.src.lightweight_openpose import lightweight_openpose
from oracle.forward_oracle import ForwardOracle

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_BF16 = 2e-2


@pytest.fixture(scope="module")
def variables():
    v, _ = checkpoint.load_or_synthesize()
    engine.set_variables(v, "synthetic")
    return v


def _maxabs(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())


@pytest.mark.parametrize("hw,batch", [((368, 368), 2), ((368, 656), 1), ((256, 256), 3), ((250, 362), 1)])
def test_forward_fp32_check_mode(variables, hw, batch):
    x = synth.synth_batch(range(batch), hw[0], hw[1])
    oh, op, oh1, op1 = ForwardOracle(variables)(x, return_stage1=True)
    h, p, h1, p1 = lightweight_openpose(x, 14, 26, is_training=False, mode="fp32", return_stage1=True)
    assert h.shape == oh.shape and p.shape == op.shape and h.dtype == np.float32
    assert _maxabs(h1, oh1) <= TOL_FP32 and _maxabs(p1, op1) <= TOL_FP32
    assert _maxabs(h, oh) <= TOL_FP32 and _maxabs(p, op) <= TOL_FP32


@pytest.mark.parametrize("mode", ["bf16_direct", "bf16"])
@pytest.mark.parametrize("hw,batch", [((368, 368), 2), ((368, 656), 1), ((256, 256), 3), ((250, 362), 2), ((136, 72), 5)])
def test_forward_bf16(variables, hw, batch, mode):
    x = synth.synth_batch(range(10, 10 + batch), hw[0], hw[1])
    oh, op, oh1, op1 = ForwardOracle(variables)(x, return_stage1=True)
    h, p, h1, p1 = lightweight_openpose(x, 14, 26, is_training=False, mode=mode, return_stage1=True)
    errs = (_maxabs(h1, oh1), _maxabs(p1, op1), _maxabs(h, oh), _maxabs(p, op))
    print(mode, hw, "max-abs", errs)
    assert max(errs) <= TOL_BF16, errs


def test_forward_batch64_matches_oracle_and_is_batch_invariant(variables):
    """config[1]: batch 64, bf16 against the oracle; each image's result must not
    depend on its batch mates (replica sharding relies on it)."""
    x = synth.synth_batch(range(64), 368, 368)
    xd = torch.from_numpy(x).cuda()
    h, p = lightweight_openpose(xd, 14, 26, is_training=False)
    h, p = h.cpu().numpy(), p.cpu().numpy()
    oh, op = ForwardOracle(variables)(x[:8])
    assert _maxabs(h[:8], oh) <= TOL_BF16 and _maxabs(p[:8], op) <= TOL_BF16
    h2, p2 = lightweight_openpose(xd[37:41].contiguous(), 14, 26, is_training=False)
    assert np.array_equal(h2.cpu().numpy(), h[37:41]) and np.array_equal(p2.cpu().numpy(), p[37:41])


def test_forward_batch64_fp32_check_and_bf16_on_every_image(variables):
    """BASELINE configs[1] as written: batch 64 at 368x368, forward only, fp32 check mode (<= 1e-4) and bf16 (<= 2e-2)
    against the oracle on ALL 64 images."""
    x = synth.synth_batch(range(64), 368, 368)
    oh, op = ForwardOracle(variables)(x)
    xd = torch.from_numpy(x).cuda()
    h32, p32 = (t.cpu().numpy() for t in lightweight_openpose(xd, 14, 26, is_training=False, mode="fp32"))
    e32 = np.maximum(np.abs(h32 - oh).reshape(64, -1).max(1), np.abs(p32 - op).reshape(64, -1).max(1))
    assert e32.max() <= TOL_FP32, e32
    h16, p16 = (t.cpu().numpy() for t in lightweight_openpose(xd, 14, 26, is_training=False))
    e16 = np.maximum(np.abs(h16 - oh).reshape(64, -1).max(1), np.abs(p16 - op).reshape(64, -1).max(1))
    print("batch 64: fp32-check max-abs", e32.max(), "bf16 max-abs", e16.max())
    assert e16.max() <= TOL_BF16, e16


def test_two_handles_with_different_weights_on_one_device(variables):
    """Every layer's weights belong to the handle (the first conv's taps travel as a kernel parameter): two engines
    with different checkpoints on one device must not see each other's layer 0."""
    other = {k: (v * np.float32(2.0) if k == "BackBone/conv2d/kernel" else v) for k, v in variables.items()}
    assert any(not np.array_equal(other[k], variables[k]) for k in variables)
    x = synth.synth_batch([1, 2], 256, 256)
    xd = torch.from_numpy(x).cuda()
    ea, eb = engine.Engine(256, 256, max_batch=2), engine.Engine(256, 256, max_batch=2)
    ea.load_variables(variables, "a")
    eb.load_variables(other, "b")                     # loaded last: a process-global table would now hold b's taps
    for mode, tol in (("bf16", TOL_BF16), ("fp32", TOL_FP32)):
        ha, pa = (t.clone() for t in ea.forward(xd, mode=mode))
        hb, pb = (t.clone() for t in eb.forward(xd, mode=mode))
        ha2, _ = ea.forward(xd, mode=mode)            # again after b ran
        assert torch.equal(ha, ha2)
        oa, _ = ForwardOracle(variables)(x)
        ob, _ = ForwardOracle(other)(x)
        gap = _maxabs(oa, ob)                         # the two checkpoints differ by ~1 at the output
        assert gap > 0.5
        assert _maxabs(ha.cpu().numpy(), oa) <= tol and _maxabs(hb.cpu().numpy(), ob) <= max(tol, 0.05 * gap)


def test_eager_net_factory_path_matches_executor(variables):
    x = synth.synth_batch([3], 256, 256)
    h, p = lightweight_openpose(x, 14, 26, is_training=False, mode="fp32")
    he, pe = lightweight_openpose(x, 14, 26, is_training=False, mode="fp32", eager=True)
    assert _maxabs(h, he) <= 1e-5 and _maxabs(p, pe) <= 1e-5
    hb, pb = lightweight_openpose(x, 14, 26, is_training=False, mode="bf16")
    hbe, pbe = lightweight_openpose(x, 14, 26, is_training=False, mode="bf16", eager=True)
    assert _maxabs(hb, hbe) <= TOL_BF16 and _maxabs(pb, pbe) <= TOL_BF16


def test_uint8_input_path(variables):
    u8 = np.stack([synth.synth_image(s, 256, 256) for s in (0, 1)])
    xf = (u8 / 255.).astype(np.float32)
    eng = engine.get_engine(256, 256, 2)
    hu, pu = (t.clone() for t in eng.forward(torch.from_numpy(u8).cuda()))
    hf, pf = eng.forward(torch.from_numpy(xf).cuda())
    assert torch.equal(hu, hf) and torch.equal(pu, pf)         # `/ 255.` on the device == on the host, bit for bit
    oh, op = ForwardOracle(variables)(xf)
    assert _maxabs(hu.cpu().numpy(), oh) <= TOL_BF16 and _maxabs(pu.cpu().numpy(), op) <= TOL_BF16
    hu32 = eng.forward(torch.from_numpy(u8).cuda(), mode="fp32")[0].clone()
    hf32, _ = eng.forward(torch.from_numpy(xf).cuda(), mode="fp32")
    assert torch.equal(hu32, hf32)


def test_graph_replay_is_deterministic_and_counts_launches(variables):
    x = torch.from_numpy(synth.synth_batch([5, 6], 368, 368)).cuda()
    eng = engine.get_engine(368, 368, 2)
    a = tuple(t.clone() for t in eng.forward(x))
    eng.launch_count(reset=True)
    b = eng.forward(x)
    n = eng.launch_count()
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert 30 <= n <= 57    # 57 conv steps; fused separable layers and fused head pairs launch one kernel for 2 / 4 steps


def test_async_host_path_matches_synchronous(variables):
    """lwop_infer_host_begin/_end with two batches in flight: same tables as one blocking lwop_infer_host per batch
    and as forward + decode on device-resident frames."""
    eng = engine.Engine(368, 368, max_batch=6, device=0)
    eng.load_variables(variables, "synthetic")
    batches = [np.stack([synth.synth_image(100 * k + s, 368, 368) for s in range(n)]) for k, n in enumerate((6, 4, 6, 1))]
    pinned = [torch.from_numpy(b).pin_memory() for b in batches]
    sync = [eng.infer_host(p) for p in pinned]
    tickets = [eng.infer_host_begin(pinned[0])]
    got = []
    for p in pinned[1:]:
        tickets.append(eng.infer_host_begin(p))
        got.append(eng.infer_host_end(tickets.pop(0)))
    got.append(eng.infer_host_end(tickets.pop(0)))
    for a, b, frames in zip(got, sync, batches):
        assert len(a) == len(b) == frames.shape[0]
        for ra, rb in zip(a, b):
            assert np.array_equal(ra.joint_list, rb.joint_list)
            assert np.array_equal(ra.person_to_joint_assoc, rb.person_to_joint_assoc)
            assert np.array_equal(ra.joints, rb.joints)
        heat, paf = eng.forward(torch.from_numpy(frames).cuda())
        dev = eng.decode(heat, paf)
        for ra, rd in zip(a, dev):
            assert np.array_equal(ra.joint_list, rd.joint_list) and np.array_equal(ra.joints, rd.joints)


def test_forward_is_bitwise_repeatable_over_many_replays(variables):
    """40 graph replays of the same 64-frame batch: maps identical bit for bit (see the fused-separable regression test
    in test_gpu_ops.py for the race this guards against)."""
    eng = engine.Engine(368, 368, max_batch=64)
    eng.load_variables(variables, "test")
    x = torch.from_numpy(np.stack([synth.synth_image(s, 368, 368) for s in range(32)] * 2)).cuda()
    heat, paf = eng.forward(x)
    torch.cuda.synchronize()
    h0, p0 = heat.clone(), paf.clone()
    for run in range(40):
        heat, paf = eng.forward(x)
        torch.cuda.synchronize()
        assert torch.equal(heat, h0) and torch.equal(paf, p0), f"replay {run} differs"
    eng.close()


def test_forward_on_rough_random_weights():
    """Non-benign weights: iid He-normal kernels -- every 3x3 filter a random high-pass, no
    identity path -- with one live-calibrated scalar gain per layer (tools/rough_weights.py).  Such a network amplifies
    perturbations, so absolute tolerances tuned on the benign synthetic checkpoint say little; what is asserted is
    (a) the fp32 check mode still meets 1e-4 against the exact oracle, and (b) the bf16 path is as accurate as the bf16
    FORMAT permits: its error against the fp64 oracle is no larger than that of the CPU oracle run with emulated bf16
    storage of weights and activations (measured: rms 3.58e-3 on the GPU against 3.60e-3 emulated, 1.2 % of the output
    standard deviation)."""
    from tools.rough_weights import errors, make_rough_variables
    v = make_rough_variables()
    H = W = 184
    x = synth.synth_batch([10, 11], H, W)
    o64 = ForwardOracle(v, dtype=torch.float64)(x)
    o32 = ForwardOracle(v)(x)
    emu = ForwardOracle(v, quantize=True)(x)
    eng = engine.Engine(H, W, max_batch=2)
    eng.load_variables(v, "rough")
    xd = torch.from_numpy(x).cuda()
    g16 = [t.cpu().numpy() for t in eng.forward(xd, mode="bf16")]
    g32 = [t.cpu().numpy() for t in eng.forward(xd, mode="fp32")]
    eng.close()
    for k in range(2):
        assert _maxabs(g32[k], o32[k]) <= TOL_FP32
        std = float(np.std(o64[k]))
        gm, gr = errors(g16[k], o64[k])
        em, er = errors(emu[k], o64[k])
        assert er > 1e-4 * std                      # the emulation does perturb: the comparison is not vacuous
        assert gr <= 1.25 * er and gm <= 1.5 * em, (k, gm, gr, em, er)
        assert gr <= 0.02 * std, (k, gr, std)


def test_other_head_widths_run_the_eager_graph(variables):
    """lightweight_openpose(inputs, num_joints, num_pafs) with widths other than the checkpoint's 14 / 26 (reference
    signature, src/lightweight_openpose.py:8): same graph from the net_factory ops on a variable set of matching shapes."""
    from lightweight_openpose_b200 import graph
    nj, npaf = 19, 38
    rng = np.random.default_rng(3)
    v = dict(variables)
    for layer in graph.LAYERS:
        if layer.note in ("heatmaps1", "heatmaps2", "pafs1", "pafs2"):
            width = nj if layer.note.startswith("heat") else npaf
            k = v[layer.name + "/kernel"]
            v[layer.name + "/kernel"] = (rng.standard_normal(k.shape[:3] + (width,)) * k.std()).astype(np.float32)
            v[layer.name + "/bias"] = rng.normal(0, 0.01, width).astype(np.float32)
        elif layer.kind == graph.KIND_CONV and layer.cin == 40:
            k = v[layer.name + "/kernel"]
            v[layer.name + "/kernel"] = (rng.standard_normal((1, 1, nj + npaf, k.shape[3])) * k.std()).astype(np.float32)
    x = synth.synth_batch([5], 128, 128)
    oh, op = ForwardOracle(v)(x, num_joints=nj, num_pafs=npaf)
    assert oh.shape == (1, 16, 16, nj) and op.shape == (1, 16, 16, npaf)
    engine.set_variables(v, "other widths")
    try:
        h, p = lightweight_openpose(x, nj, npaf, is_training=False, mode="fp32")
        assert h.shape == oh.shape and p.shape == op.shape
        assert _maxabs(h, oh) <= TOL_FP32 and _maxabs(p, op) <= TOL_FP32
        hb, pb = lightweight_openpose(x, nj, npaf, is_training=False, mode="bf16")
        assert _maxabs(hb, oh) <= 2 * TOL_BF16 and _maxabs(pb, op) <= 2 * TOL_BF16
    finally:
        engine.set_variables(variables, "synthetic")
