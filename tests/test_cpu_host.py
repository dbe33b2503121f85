"""CPU-side checks: checkpoint index parser, weight folding/packing against the
library's layer table, C-ABI exports, synthetic generators."""
import ctypes as C
import os

import numpy as np
import pytest

from lightweight_openpose_b200 import _lib, checkpoint, graph, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIX = os.path.join(ROOT, "model", "model.ckpt-61236")


def test_index_parses_to_survey_inventory():
    entries, header = checkpoint.read_index(PREFIX + ".index")
    assert header["num_shards"] == 1
    assert len(entries) == 435
    assert max(e.offset + e.size for e in entries.values()) == checkpoint.EXPECTED_SHARD_BYTES
    assert entries["global_step"].dtype == checkpoint.DT_INT64 and entries["global_step"].shape == ()
    assert entries["lr"].offset == 48_706_256 and entries["lr"].size == 4
    inf = checkpoint.inference_entries(entries)
    assert len(inf) == 173
    assert entries["BackBone/conv2d/kernel"].shape == (3, 3, 3, 32)
    assert entries["RefinementStage/conv2d/kernel"].shape == (1, 1, 40, 128)
    assert entries["InitialStage/conv2d_4/kernel"].shape == (1, 1, 512, 14)
    assert entries["InitialStage/conv2d_6/bias"].shape == (26,)
    assert all(e.crc32c is not None for e in entries.values())
    # conv weights only: 4 041 024
    n = sum(int(np.prod(e.shape)) for k, e in inf.items() if k.endswith("kernel"))
    assert n == 4_041_024


def test_layer_table_binds_every_inference_variable_once():
    entries, _ = checkpoint.read_index(PREFIX + ".index")
    inf = set(checkpoint.inference_entries(entries))
    used = []
    for layer in graph.LAYERS:
        used += graph.variable_names(layer)
    assert len(used) == len(set(used))
    assert set(used) == inf
    for layer in graph.LAYERS:
        for name in graph.variable_names(layer):
            shape = entries[name].shape
            leaf = name.rsplit("/", 1)[1]
            if leaf == "kernel":
                assert shape == (layer.k, layer.k, layer.cin, layer.cout), name
            elif leaf == "pointwise_kernel":
                assert shape == (1, 1, layer.cin, layer.cout), name
            elif leaf == "depthwise_kernel":
                assert shape == (3, 3, layer.cin, 1), name
            else:
                assert shape == (layer.cout,), name


def test_synthetic_shard_roundtrip(tmp_path):
    entries, _ = checkpoint.read_index(PREFIX + ".index")
    # write a shard next to a copy of the index and read it back through the loader
    import shutil
    pre = str(tmp_path / "model.ckpt-61236")
    shutil.copy(PREFIX + ".index", pre + ".index")
    n = checkpoint.write_synthetic_shard(PREFIX, checkpoint.data_shard_path(pre))
    assert n == checkpoint.EXPECTED_SHARD_BYTES
    loaded = checkpoint.load_variables(pre, verify_crc=False)
    syn = checkpoint.synthetic_variables(entries)
    assert set(loaded) == set(syn)
    for k in syn:
        assert np.array_equal(loaded[k], syn[k]), k
    with pytest.raises(ValueError):
        checkpoint.load_variables(pre, verify_crc=True)      # synthetic bytes do not match the real CRCs


def test_crc32c_known_answers():
    # RFC 3720 test vectors
    assert checkpoint.crc32c(b"") == 0
    assert checkpoint.crc32c(b"123456789") == 0xE3069283
    assert checkpoint.crc32c(bytes(32)) == 0x8A9136AA
    assert checkpoint.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert checkpoint.mask_crc(0) == 0xA282EAD8


def test_fold_matches_unfolded_arithmetic():
    entries, _ = checkpoint.read_index(PREFIX + ".index")
    v = checkpoint.synthetic_variables(entries)
    layer = graph.LAYERS[25]          # RefinementStage conv with bias and BN
    assert layer.bn and layer.bias
    f = graph.fold_layer(layer, v)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(9 * 128,))
    w = v[layer.name + "/kernel"].astype(np.float64).reshape(9 * 128, 128)
    pre = x @ w + v[layer.name + "/bias"]
    g, b = v[layer.bn + "/gamma"], v[layer.bn + "/beta"]
    m, var = v[layer.bn + "/moving_mean"], v[layer.bn + "/moving_variance"]
    want = g * (pre - m) / np.sqrt(var + 1e-3) + b
    got = f.w.astype(np.float64).reshape(128, 9 * 128) @ x + f.b
    assert np.abs(got - want).max() < 1e-4


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.header_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
    assert lib.lwop_abi_version() == 2
    assert lib.lwop_same_out(368, 2) == 184 and lib.lwop_same_out(45, 2) == 23
    assert lib.lwop_error_string(-4) == b"decode table overflow"


def test_library_blob_size_matches_python_table():
    lib = _lib.load()
    assert lib.lwop_weights_num_floats() == graph.blob_floats()
    entries, _ = checkpoint.read_index(PREFIX + ".index")
    blob = graph.pack_blob(graph.fold_all(checkpoint.synthetic_variables(entries)))
    assert blob.size == graph.blob_floats() and blob.dtype == np.float32
    assert np.isfinite(blob).all()


def test_library_crc32c_matches_python():
    lib = _lib.load()
    data = np.random.default_rng(1).integers(0, 256, size=100003, dtype=np.uint8).tobytes()
    assert lib.lwop_crc32c(data, len(data)) == checkpoint.crc32c(data)
    assert lib.lwop_crc32c(b"123456789", 9) == 0xE3069283


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lightweight_openpose_b200.src.lightweight_openpose import lightweight_openpose
    from lightweight_openpose_b200.src.pose_decode import decode_pose
    x = np.zeros((1, 64, 64, 3), dtype=np.float32)
    with pytest.raises(_lib.LwopError):
        lightweight_openpose(x, 14, 26, is_training=False)
    with pytest.raises(_lib.LwopError):
        decode_pose(np.zeros((64, 64, 3), np.uint8), {"thre1": 0.1, "thre2": 0.0},
                    np.zeros((8, 8, 14), np.float32), np.zeros((8, 8, 26), np.float32))
    h = C.c_void_p()
    rc = _lib.load().lwop_create(0, 1, 64, 64, 0, C.byref(h))
    assert rc == _lib.ERR_NO_DEVICE and not h.value


def test_training_mode_is_rejected():
    from lightweight_openpose_b200.src.lightweight_openpose import lightweight_openpose
    with pytest.raises(NotImplementedError):
        lightweight_openpose(np.zeros((1, 64, 64, 3), np.float32), 14, 26)      # is_training defaults to True


def test_synth_generators_are_deterministic():
    a = synth.synth_image(5, 64, 96)
    b = synth.synth_image(5, 64, 96)
    assert a.dtype == np.uint8 and a.shape == (64, 96, 3) and np.array_equal(a, b)
    assert not np.array_equal(a, synth.synth_image(6, 64, 96))
    h1, p1, k1 = synth.synth_scene(3, 3, 368, 368)
    h2, p2, _ = synth.synth_scene(3, 3, 368, 368)
    assert h1.shape == (46, 46, 14) and p1.shape == (46, 46, 26) and k1.shape == (3, 14, 3)
    assert np.array_equal(h1, h2) and np.array_equal(p1, p2)
    assert h1.max() > 0.9 and np.abs(p1).max() > 0.5


def test_u8_normalisation_is_single_rounding():
    # reference: img_data / 255. in float64, then fed as float32 (model_test.py:65,39)
    u = np.arange(256, dtype=np.uint8)
    ref = (u / 255.).astype(np.float32)
    assert np.array_equal(ref, (u.astype(np.float64) / 255.0).astype(np.float32))
    assert np.array_equal(synth.synth_batch([0], 32, 32)[0], (synth.synth_image(0, 32, 32) / 255.).astype(np.float32))


def test_host_frame_validation():
    """Engine.infer_host / infer_host_begin hand a raw pointer to C: shape, batch, dtype and contiguity are checked
    (and float64 frames -- the reference's own `img / 255.` -- cast to float32) before that."""
    import torch
    from lightweight_openpose_b200.engine import validate_host_images as val
    u8 = np.zeros((2, 64, 48, 3), np.uint8)
    out = val(u8, 64, 48, 4)
    assert out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (2, 64, 48, 3)
    f64 = np.random.default_rng(0).random((1, 64, 48, 3))
    out = val(f64, 64, 48, 4)
    assert out.dtype == torch.float32 and np.array_equal(out.numpy(), f64.astype(np.float32))
    assert val(torch.zeros((1, 64, 48, 3), dtype=torch.float16), 64, 48, 4).dtype == torch.float32
    nc = torch.zeros((2, 64, 48, 6), dtype=torch.float32)[..., ::2]
    assert not nc.is_contiguous() and val(nc, 64, 48, 4).is_contiguous()
    for bad in (np.zeros((2, 48, 64, 3), np.uint8), np.zeros((64, 48, 3), np.uint8), np.zeros((5, 64, 48, 3), np.uint8),
                np.zeros((0, 64, 48, 3), np.uint8), np.zeros((2, 64, 48, 3), np.int32), np.zeros((2, 64, 48, 4), np.uint8)):
        with pytest.raises(ValueError):
            val(bad, 64, 48, 4)


def test_fits_executor_and_general_depthwise_folding():
    """Host logic behind the non-default reference arguments: a variable set with other head widths is recognised as not
    loadable into the 14 / 26 executor, and a separable layer with a 5x5 depthwise kernel folds to [25][cin] taps."""
    from lightweight_openpose_b200 import graph
    entries, _ = checkpoint.read_index(checkpoint.default_index_prefix() + ".index")
    v = checkpoint.synthetic_variables(entries)
    assert graph.fits_executor(v)
    heads = [l for l in graph.LAYERS if l.note in ("heatmaps1", "heatmaps2", "pafs1", "pafs2")]
    assert len(heads) == 4 and sorted(l.cout for l in heads) == [14, 14, 26, 26]
    w = dict(v)
    k = w[heads[0].name + "/kernel"]
    w[heads[0].name + "/kernel"] = np.zeros(k.shape[:3] + (19,), np.float32)
    assert not graph.fits_executor(w)
    missing = dict(v)
    del missing[heads[1].name + "/kernel"]
    assert not graph.fits_executor(missing)

    rng = np.random.default_rng(0)
    cin, cout = 6, 10
    layer = graph.Layer(-1, graph.KIND_SEPARABLE, "T/separable_conv2d", None, cin, cout, 1, 1, 1, graph.ACT_ELU, False, dw_k=5)
    vv = {"T/separable_conv2d/depthwise_kernel": rng.standard_normal((5, 5, cin, 1)).astype(np.float32),
          "T/separable_conv2d/pointwise_kernel": rng.standard_normal((1, 1, cin, cout)).astype(np.float32)}
    f = graph.fold_layer(layer, vv)
    assert f.dw.shape == (25, cin) and f.w.shape == (cout, 1, cin) and f.b.shape == (cout,)
    assert np.array_equal(f.dw[7], vv["T/separable_conv2d/depthwise_kernel"][1, 2, :, 0])      # tap index = dy * 5 + dx
    assert np.array_equal(f.w[:, 0, :], vv["T/separable_conv2d/pointwise_kernel"][0, 0].T)
    with pytest.raises(AssertionError):
        graph.fold_layer(graph.Layer(-1, graph.KIND_SEPARABLE, "T/separable_conv2d", None, cin, cout, 1, 1, 1, graph.ACT_ELU, False), vv)


def test_every_entry_point_is_mapped_in_integration_md():
    """INTEGRATION.md names, for every function include/lwop.h declares, the reference lines it replaces."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "INTEGRATION.md"), encoding="utf-8") as fh:
        text = fh.read()
    missing = [f for f in _lib.header_functions() if f"`{f}`" not in text]
    assert not missing, missing
