"""The forward oracle against the reference's OWN model function.

TensorFlow cannot be installed here, so `/root/reference/src/lightweight_openpose.py` cannot run as shipped.  With the
API shim of tests/tf_shim.py it can: the unmodified reference function builds the network (layer order, widths,
strides, dilations, placement of BN / ReLU / ELU / bias, the tf.add and tf.concat edges, the order in which
variables are created and therefore their auto-generated names) and the oracle must agree with it.  What stays a
restatement is the arithmetic of each op (TF 'same' padding, separable = depthwise then pointwise, BN eps)."""
import os

import numpy as np
import pytest

from lightweight_openpose_b200 import checkpoint, synth
from oracle.forward_oracle import ForwardOracle

import tf_shim

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")


@pytest.fixture(scope="module")
def variables():
    v, _ = checkpoint.load_or_synthesize()
    return v


@pytest.mark.parametrize("hw,batch", [((64, 64), 2), ((72, 104), 1), ((50, 66), 1)])
def test_oracle_matches_reference_model_function_on_tf_shim(variables, hw, batch):
    x = synth.synth_batch(range(batch), hw[0], hw[1])
    import torch
    with tf_shim.installed(variables) as state:
        from src.lightweight_openpose import lightweight_openpose as ref_model      # the reference file itself
        with torch.no_grad():
            h_ref, p_ref = ref_model(torch.from_numpy(x), 14, 26, is_training=False)
        used = list(state.used)
    h, p = ForwardOracle(variables)(x)
    assert h.shape == tuple(h_ref.shape) and p.shape == tuple(p_ref.shape)
    assert np.abs(h - h_ref.numpy()).max() <= 1e-5 and np.abs(p - p_ref.numpy()).max() <= 1e-5
    # every inference variable of the checkpoint is consumed exactly once, under the name TensorFlow gave it
    assert len(used) == len(set(used))
    inference_vars = {k for k in variables if not k.endswith(("Adam", "Adam_1")) and "global_step" not in k
                      and "beta1_power" not in k and "beta2_power" not in k}
    assert set(used) == inference_vars, (sorted(set(used) ^ inference_vars)[:8])


def test_same_padding_agrees_with_third_party_tf_ports():
    """TensorFlow's 'SAME' split (pad_before = total // 2, the larger half after) is restated in the oracle, in the shim and
    in the CUDA kernels by one author.  An independent statement of the same rule ships in this image: Hugging Face's
    PyTorch ports of the TF-slim MobileNets pad their convolutions with ``apply_tf_padding`` so that converted TensorFlow
    checkpoints reproduce TF's outputs (v1: stride / kernel; v2: also dilation).  The oracle's ``same_pad`` must pad
    exactly as they do for every conv geometry of the network (3x3 and 1x1, stride 1 and 2, dilation 1 and 2) on the
    frame sizes used here, even and odd."""
    transformers = pytest.importorskip("transformers")
    import torch
    from torch import nn
    from transformers.models.mobilenet_v1.modeling_mobilenet_v1 import apply_tf_padding as pad_v1
    from transformers.models.mobilenet_v2.modeling_mobilenet_v2 import apply_tf_padding as pad_v2
    from oracle.forward_oracle import same_pad

    def hf_pads(fn, h, w, k, s, d):
        conv = nn.Conv2d(1, 1, k, stride=s, dilation=d)
        x = torch.ones(1, 1, h, w)
        y = fn(x, conv)
        rows = y[0, 0].sum(1).nonzero().flatten()
        cols = y[0, 0].sum(0).nonzero().flatten()
        return (int(rows[0]), y.shape[2] - 1 - int(rows[-1])), (int(cols[0]), y.shape[3] - 1 - int(cols[-1]))

    n = 0
    for h, w in ((368, 368), (368, 656), (256, 256), (250, 362), (136, 72), (184, 92), (46, 82), (45, 47), (125, 181), (17, 9)):
        for k in (1, 3):
            for s in (1, 2):
                for d in (1, 2):
                    if s == 2 and d == 2:
                        continue                       # TF rejects strides > 1 with dilation > 1; the network has none
                    (_, pt, pb), (_, pl, pr) = same_pad(h, k, s, d), same_pad(w, k, s, d)
                    fns = (pad_v1, pad_v2) if d == 1 else (pad_v2,)
                    for fn in fns:
                        assert hf_pads(fn, h, w, k, s, d) == ((pt, pb), (pl, pr)), (fn.__module__, h, w, k, s, d)
                        n += 1
    assert n >= 100
