"""checkpoints in the reference's HDF5 layout.

CPU part: the pure-Python HDF5 subset (dorknet_b200/minih5.py) round-trips everything a Dorknet checkpoint holds, its
files have the structure the HDF5 specification prescribes for the "earliest" layout, and the LIVE reference
(oracle/_ref) saves and re-loads a trained network through it with its own unmodified save_weights_to_h5 /
load_network_from_json_and_h5 (network/feed_forward_network.py:90-139).
GPU part (-m gpu): a checkpoint written by the reference is loaded by the CUDA layers and scores the same, and the
other way round; optimiser state survives a save / load.

NOT verified here (no libhdf5 / h5py in this image, no network): byte-level interoperability with the HDF5 library
itself.  minih5 is written against the published file-format specification; DESIGN.md says so."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from dorknet_b200 import minih5  # noqa: E402

HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", ".built"))


def _reference():
    sys.path.insert(0, ROOT)
    from oracle import refload
    if not hasattr(np, "string_"):  # NumPy 2 dropped the alias the reference's save_to_h5 uses (convolution.py:242)
        np.string_ = np.bytes_
    return refload.load_reference()


# ------------------------------------------------------------------------------------------------ minih5 alone
def test_minih5_roundtrip_types(tmp_path):
    fn = str(tmp_path / "t.h5")
    rng = np.random.default_rng(0)
    w = rng.standard_normal((7, 3, 5, 5)).astype(np.float32)
    with minih5.File(fn, "w") as f:
        info = f.create_dataset("conv0/layer_info", dtype=np.float32)
        info.attrs["type"] = "ConvLayer"
        info.attrs["with_bias"] = False
        info.attrs["yes"] = np.bool_(True)
        info.attrs["num_filters"] = 64
        info.attrs["eps"] = 1e-5
        info.attrs["run_momentum"] = np.float32(0.95)
        info.attrs["layer_type_list"] = ["DepthwiseConvLayer", "BatchNormLayer", "ReLu"]
        info.attrs["unicode"] = "naïve ✓"
        d = f.create_dataset("conv0/weights", w.shape, dtype=w.dtype)
        d[:] = w
        d.attrs["weight_regulariser_type"] = np.bytes_("l2")
        d.attrs["weight_regulariser_strength"] = np.bytes_(0.0001)
        f.create_dataset("conv0/grads/weights", data=2 * w)
        f.create_dataset("ints", data=np.arange(12, dtype=np.int32).reshape(3, 4))
        f.create_dataset("f64", data=np.linspace(0, 1, 5))
        f.create_dataset("scalar", data=np.float32(3.5))
    with minih5.File(fn, "r") as f:
        a = f["conv0/layer_info"].attrs
        assert a["type"] == "ConvLayer" and isinstance(a["type"], str)
        assert a["with_bias"] == False and a["yes"] == True  # noqa: E712
        assert a["num_filters"] == 64 and isinstance(a["num_filters"], np.integer)
        assert a["eps"] == 1e-5 and a["run_momentum"] == np.float32(0.95)
        assert list(a["layer_type_list"]) == ["DepthwiseConvLayer", "BatchNormLayer", "ReLu"]
        assert a["unicode"] == "naïve ✓"
        assert a.get("missing", None) is None and "type" in a
        assert f["conv0/layer_info"].shape is None
        d = f["conv0/weights"]
        assert d.shape == w.shape and d.dtype == np.float32
        np.testing.assert_array_equal(d[:], w)
        np.testing.assert_array_equal(d[...], w)
        assert d.attrs["weight_regulariser_type"] == b"l2"
        assert float(d.attrs["weight_regulariser_strength"]) == 1e-4
        np.testing.assert_array_equal(f["conv0/grads/weights"][:], 2 * w)
        np.testing.assert_array_equal(f["conv0"]["grads"]["weights"][:], 2 * w)
        np.testing.assert_array_equal(f["ints"][:], np.arange(12, dtype=np.int32).reshape(3, 4))
        np.testing.assert_array_equal(f["f64"][:], np.linspace(0, 1, 5))
        assert f["scalar"][()] == np.float32(3.5)
        assert sorted(f.keys()) == ["conv0", "f64", "ints", "scalar"]
        assert "conv0/grads" in f and "conv0/nope" not in f
        with pytest.raises(KeyError):
            f["conv0/nope"]


def test_minih5_many_groups_and_long_names(tmp_path):
    """ResNet-18-depsep writes ~150 top-level groups: several symbol-table nodes under one B-tree node."""
    fn = str(tmp_path / "many.h5")
    names = ["res%d_dw%d_%s" % (i, j, s) for i in range(1, 9) for j in (1, 2) for s in
             ("dw", "dw_bn", "pw", "pw_bn", "pw_relu", "a_rather_long_layer_name_to_cross_heap_alignment")]
    with minih5.File(fn, "w") as f:
        for i, n in enumerate(names):
            f.create_dataset(n + "/layer_info", dtype=np.float32).attrs["type"] = "ReLu"
            f.create_dataset(n + "/weights", data=np.full((3, 2), i, np.float32))
    with minih5.File(fn, "r") as f:
        assert sorted(f.keys()) == sorted(names)
        for i, n in enumerate(names):
            assert f[n + "/weights"][0, 0] == i and f[n + "/layer_info"].attrs["type"] == "ReLu"


def test_minih5_file_structure_follows_the_specification(tmp_path):
    """Superblock version 0 exactly as HDF5 File Format Specification III.A lays it out, EOF address = file length,
    root symbol-table entry caches the B-tree / heap addresses, which carry the TREE / HEAP signatures."""
    fn = str(tmp_path / "s.h5")
    with minih5.File(fn, "w") as f:
        f.create_dataset("a/b", data=np.arange(4, dtype=np.float32))
    raw = open(fn, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    ver_sb, ver_fs, ver_root, _, ver_shared, size_off, size_len, _ = struct.unpack_from("<8B", raw, 8)
    assert (ver_sb, ver_fs, ver_root, ver_shared, size_off, size_len) == (0, 0, 0, 0, 8, 8)
    leaf_k, int_k, flags = struct.unpack_from("<HHI", raw, 16)
    assert leaf_k == 4 and int_k == 16 and flags == 0
    base, free, eof, driver = struct.unpack_from("<4Q", raw, 24)
    assert base == 0 and free == minih5.UNDEF and driver == minih5.UNDEF and eof == len(raw)
    name_off, hdr, cache_type, _, btree, heap = struct.unpack_from("<QQIIQQ", raw, 56)
    assert name_off == 0 and cache_type == 1
    assert raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    assert raw[hdr] == 1  # object header version 1
    snod = struct.unpack_from("<Q", raw, btree + 24 + 8)[0]
    assert raw[snod:snod + 4] == b"SNOD"
    for off in (hdr, btree, heap, snod):
        assert off % 8 == 0


def test_minih5_rejects_what_it_does_not_implement(tmp_path):
    fn = str(tmp_path / "bad.h5")
    open(fn, "wb").write(b"not an hdf5 file at all" * 10)
    with pytest.raises(Exception):
        minih5.File(fn, "r")
    with pytest.raises(ValueError):
        minih5.File(fn, "a")


# ------------------------------------------------------------------------------------------------ live reference
def _train_one_step(M, net, X, Y, lr=0.05):
    opt = M.SGDMomentum(net, lr, 0.9)
    loss, _ = net.forward(X, Y)
    net.backward()
    opt.update_weights()
    return float(loss), opt


def _mini_batch(seed=3, n=4):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 3, 33, 33)).astype(np.float32)
    Y = np.zeros((n, 5), np.float32)
    Y[np.arange(n), rng.integers(0, 5, n)] = 1
    return X, Y


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_reference_saves_and_reloads_through_minih5(tmp_path):
    """The reference's own, unmodified checkpoint code on top of minih5 (bound as `h5py`)."""
    from net_defs import build_small_net, iter_param_layers
    R = _reference()
    net = build_small_net(R, seed=5)
    X, Y = _mini_batch()
    _train_one_step(R, net, X, Y)
    _, scores = net.forward(X, None, test_mode=True)
    h5, js = str(tmp_path / "ref.h5"), str(tmp_path / "ref.json")
    net.save_weights_to_h5(h5)
    net.save_layer_structure_to_json(js)
    net2 = R.FeedForwardNetwork("empty")
    net2.load_network_from_json_and_h5(js, h5)
    assert net2.name == "mini" and len(net2.layers) == len(net.layers)
    for a, b in zip(iter_param_layers(net), iter_param_layers(net2)):
        assert a.layer_name == b.layer_name
        for k in a.learned_params:
            np.testing.assert_array_equal(a.learned_params[k], b.learned_params[k])
            np.testing.assert_array_equal(a.grads[k], b.grads[k])
    _, scores2 = net2.forward(X, None, test_mode=True)
    np.testing.assert_array_equal(scores, scores2)


# ------------------------------------------------------------------------------------------------ CUDA layers
@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (the compiled reference) did not travel")
def test_checkpoint_compatibility_with_the_reference(tmp_path):
    """reference (CPU) -> file -> CUDA layers, and CUDA layers -> file -> reference: identical parameters, gradients,
    running statistics; test-mode scores agree to the TF32 GEMM tolerance."""
    import dorknet_b200.workloads as wl
    from net_defs import build_small_net, iter_param_layers
    R, M = _reference(), wl.ours()
    X, Y = _mini_batch()
    ref = build_small_net(R, seed=5)
    _train_one_step(R, ref, X, Y)
    _, ref_scores = ref.forward(X, None, test_mode=True)
    h5, js = str(tmp_path / "ref.h5"), str(tmp_path / "ref.json")
    ref.save_weights_to_h5(h5)
    ref.save_layer_structure_to_json(js)

    ours = M.FeedForwardNetwork("empty")
    ours.load_network_from_json_and_h5(js, h5)
    assert [type(l).__name__ for l in ours.layers] == [type(l).__name__ for l in ref.layers]
    for a, b in zip(iter_param_layers(ref), iter_param_layers(ours)):
        assert a.layer_name == b.layer_name and type(a).__name__ == type(b).__name__
        for k in a.learned_params:
            np.testing.assert_array_equal(a.learned_params[k], np.asarray(b.learned_params[k]))
            np.testing.assert_array_equal(a.grads[k], np.asarray(b.grads[k]))
        if type(a).__name__ == "BatchNormLayer":
            for k in ("running_mean", "running_std"):
                np.testing.assert_array_equal(a.non_learned_params[k], np.asarray(b.non_learned_params[k]))
        if getattr(a, "weight_regulariser", None) is not None:
            assert b.weight_regulariser.strength == a.weight_regulariser.strength
    _, s = ours.forward(X, None, test_mode=True)
    s = s.get()
    assert np.abs(s - ref_scores).max() <= 2e-3 * np.abs(ref_scores).max() + 1e-6

    # train ours one step from the loaded state, save, let the reference load it
    loss, opt = _train_one_step(M, ours, X, Y)
    _, s1 = ours.forward(X, None, test_mode=True)
    s1 = s1.get()
    h5b, jsb = str(tmp_path / "ours.h5"), str(tmp_path / "ours.json")
    ours.save_weights_to_h5(h5b, optimiser=opt)
    ours.save_layer_structure_to_json(jsb)
    back = R.FeedForwardNetwork("empty")
    back.load_network_from_json_and_h5(jsb, h5b)
    for a, b in zip(iter_param_layers(ours), iter_param_layers(back)):
        for k in b.learned_params:
            np.testing.assert_array_equal(a.learned_params[k].get(), b.learned_params[k])
            np.testing.assert_array_equal(a.grads[k].get(), b.grads[k])
    _, sb = back.forward(X, None, test_mode=True)
    assert np.abs(s1 - sb).max() <= 2e-3 * np.abs(sb).max() + 1e-6

    # optimiser state (f-4): a fresh optimiser resumed from the file continues exactly like the original
    again = M.FeedForwardNetwork("empty")
    opt2_holder = {}
    again.load_network_from_json_and_h5(jsb, h5b)
    opt2 = M.SGDMomentum(again, 0.05, 0.9)
    with minih5.File(h5b, "r") as f:
        from dorknet_b200.checkpoint import load_optimiser_state
        load_optimiser_state(opt2, f)
    for net, o in ((ours, opt), (again, opt2)):
        net.forward(X, Y)
        net.backward()
        o.update_weights()
    for a, b in zip(iter_param_layers(ours), iter_param_layers(again)):
        for k in a.learned_params:
            np.testing.assert_array_equal(a.learned_params[k].get(), b.learned_params[k].get())
    del opt2_holder
