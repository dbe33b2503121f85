"""Every LWOP_* kernel-selection switch still produces the network's maps: the forward of a
small batch is run in a fresh process per switch and compared with the default configuration and the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from lightweight_openpose_b200 import checkpoint, engine, synth
v, src = checkpoint.load_or_synthesize()
eng = engine.Engine(368, 368, max_batch=3)
eng.load_variables(v, src)
x = torch.from_numpy(np.stack([synth.synth_image(s, 368, 368) for s in (3, 7, 11)])).cuda()
heat, paf = eng.forward(x)
heat2, paf2 = (t.clone() for t in eng.forward(x))          # graph replay
torch.cuda.synchronize()
assert torch.equal(heat, heat2) and torch.equal(paf, paf2)
np.savez(sys.argv[1], heat=heat.cpu().numpy(), paf=paf.cpu().numpy())
""" % ROOT

SWITCHES = [
    {},                                       # default
    {"LWOP_HALO": "0"}, {"LWOP_2SM": "0"}, {"LWOP_MT": "1"}, {"LWOP_MT": "2"}, {"LWOP_CLUSTER": "2", "LWOP_2SM": "0"},
    {"LWOP_SEP_PAIR": "0"}, {"LWOP_SEP_PAIR": "0", "LWOP_SEP_NH1": "1"}, {"LWOP_SEP_WIDE64": "0"},
    {"LWOP_HEADS_PAIR": "1"}, {"LWOP_NO_FUSE_SEP": "1"}, {"LWOP_NO_FUSE_HEADS": "1"}, {"LWOP_PDL": "0"},
    {"LWOP_GEMM_PARK": "15", "LWOP_SEP_PARK": "15", "LWOP_HEADS_PARK": "15"},
    {"LWOP_GEMM_PARK": "0", "LWOP_SEP_PARK": "0", "LWOP_HEADS_PARK": "0"},
]


def _run(env_extra, out):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT, out], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (env_extra, r.stderr[-2000:])
    z = np.load(out)
    return z["heat"], z["paf"]


def test_every_kernel_selection_switch_gives_the_same_maps(tmp_path):
    base_h, base_p = _run({}, str(tmp_path / "base.npz"))
    assert np.isfinite(base_h).all() and np.abs(base_h).max() > 1e-3
    worst = {}
    for sw in SWITCHES[1:]:
        h, p = _run(sw, str(tmp_path / "sw.npz"))
        d = max(float(np.abs(h - base_h).max()), float(np.abs(p - base_p).max()))
        worst[" ".join(f"{k}={v}" for k, v in sw.items())] = d
        # different tilings / fusion boundaries may round bf16 intermediates differently; the bound is the same
        # 2e-2 the bf16 path has against the oracle (BASELINE.json)
        assert d <= 2e-2, (sw, d)
    print(worst)
