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
