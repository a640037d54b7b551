"""The oracle (oracle/oracle.py + oracle/dk_oracle.c) against the golden vectors generated
from the live reference (tests/golden/make_golden.py) and, when oracle/_ref is built, against
the reference's own compiled kernels directly.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import refload

TOL = dict(rtol=2e-5, atol=2e-6)


def close(a, b, **kw):
    t = dict(TOL)
    t.update(kw)
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), **t)


@pytest.mark.parametrize("name", ["conv_k3s1p1", "conv_k5s2p1_half", "conv_k4s2p1", "conv_k3s1p0"])
def test_conv(golden, name):
    d = golden(name)
    N, C, H, W, F, k, s, p, bias = d["meta"]
    b = d["b"] if bias else None
    Y, cache = O.conv_fwd(d["X"], d["W"], b, int(s), int(p))
    assert np.array_equal(cache["P"], d["patches"])  # im2col index map: bit-exact
    close(Y, d["Y"])
    dX, g = O.conv_bwd(d["dY"], d["W"], cache, int(s), int(p), float(d["l2"]), bool(bias))
    assert dX.shape == d["dX"].shape
    close(dX, d["dX"])
    close(g["weights"], d["dW"])
    if bias:
        close(g["bias"], d["db"])


@pytest.mark.parametrize("name", ["pw_s1", "pw_s2_even", "pw_s2_odd"])
def test_pointwise(golden, name):
    d = golden(name)
    N, C, H, W, F, s, bias = d["meta"]
    b = d["b"] if bias else None
    Y, cache = O.pointwise_fwd(d["X"], d["W"], b, int(s))
    close(Y, d["Y"])
    dX, g = O.pointwise_bwd(d["dY"], d["W"], cache, int(s), float(d["l2"]), bool(bias))
    assert dX.shape == d["dX"].shape  # zero-stuffed OH*s x OW*s (pw_s2_odd: 8x10, not 7x9)
    close(dX, d["dX"])
    close(g["weights"], d["dW"])
    if bias:
        close(g["bias"], d["db"])


@pytest.mark.parametrize("name", ["dw_k3s1p1", "dw_k3s2p1_half", "dw_k3s2p1_int", "dw_k5s1p2"])
def test_depthwise(golden, name):
    d = golden(name)
    N, C, H, W, k, s, p, bias = d["meta"]
    b = d["b"] if bias else None
    Y, cache = O.depthwise_fwd(d["X"], d["W"], b, int(s), int(p))
    close(Y, d["Y"])
    dX, g = O.depthwise_bwd(d["dY"], d["W"], cache, int(s), int(p), 0.0, bool(bias))
    assert dX.shape == d["dX"].shape
    close(dX, d["dX"])
    close(g["weights"], d["dW"])
    if bias:
        close(g["bias"], d["db"])


@pytest.mark.parametrize("name", ["bn_4d", "bn_2d"])
def test_batchnorm(golden, name):
    d = golden(name)
    Y1, cache, rm, rs = O.bn_fwd_train(d["X1"], d["gamma"], d["beta"], None, None)
    close(Y1, d["Y1"], atol=1e-5)
    close(rm, d["rm1"]); close(rs, d["rs1"])
    dX, g = O.bn_bwd(d["dY1"], d["gamma"], cache)
    close(dX, d["dX1"], atol=1e-5)
    close(g["gamma"], d["dgamma1"], rtol=1e-4); close(g["beta"], d["dbeta1"], rtol=1e-4)
    Y2, _, rm2, rs2 = O.bn_fwd_train(d["X2"], d["gamma"], d["beta"], rm, rs)
    close(Y2, d["Y2"], atol=1e-5); close(rm2, d["rm2"]); close(rs2, d["rs2"])
    close(O.bn_fwd_test(d["X1"], d["gamma"], d["beta"], rm2, rs2), d["Ytest"], atol=1e-5)


@pytest.mark.parametrize("name", ["relu_4d", "relu_2d"])
def test_relu(golden, name):
    d = golden(name)
    Y, mask = O.relu_fwd(d["X"])
    assert np.array_equal(Y, d["Y"]) and np.array_equal(mask, d["mask"])
    assert np.array_equal(O.relu_bwd(d["dY"], mask), d["dX"])
    assert np.array_equal(O.relu_fwd(d["X"], want_mask=False)[0], d["Ytest"])


def test_gap(golden):
    d = golden("gap")
    close(O.gap_fwd(d["X"]), d["Y"])
    close(O.gap_bwd(d["dY"], 8, 6), d["dX"])


@pytest.mark.parametrize("s", [2, 4])
def test_maxpool_bit_exact(golden, s):
    d = golden("maxpool_s%d" % s)
    Y, mask = O.maxpool_fwd(d["X"], s)
    assert np.array_equal(Y, d["Y"]) and np.array_equal(mask, d["mask"])
    assert np.array_equal(O.maxpool_bwd(mask, d["dY"], s), d["dX"])
    assert np.array_equal(O.maxpool_fwd(d["X"], s, train=False)[0], d["Ytest"])
    if s == 2:  # SURVEY A.8 known answer: first max in the row-major window scan wins
        assert Y[0, 0, :2, :2].tolist() == [[1, 3], [5, 0]]
        assert np.argwhere(mask[0, 0, :4, :4] == 1).tolist() == [[0, 0], [1, 3], [2, 0], [2, 2]]


def test_dense_loss_l2(golden):
    d = golden("dense")
    close(O.dense_fwd(d["X"], d["W"], d["b"]), d["Y"])
    dX, g = O.dense_bwd(d["dY"], d["X"], d["W"], float(d["l2"]))
    close(dX, d["dX"]); close(g["weights"], d["dW"]); close(g["bias"], d["db"])
    close(O.l2_fwd(d["W"], float(d["l2"])), d["reg"])
    for nm in ("softmax_hard", "softmax_soft"):
        s = golden(nm)
        loss, p = O.softmax_xent_fwd(s["X"], s["y"])
        close(loss, s["loss"]); close(p, s["p"]); close(O.softmax_xent_bwd(p, s["y"]), s["dX"])


def test_optimisers(golden):
    for nm in ("opt_sgd", "opt_sgdm", "opt_rmsprop"):
        d = golden(nm)
        for key, gk in (("w", "gw"), ("b", "gb")):
            w = d[key + "0"]
            state = np.zeros_like(w)
            for i in range(3):
                g = d["%s%d" % (gk, i)]
                if nm == "opt_sgd":
                    w = O.sgd_update(w, g, 0.1)
                elif nm == "opt_sgdm":
                    w, state = O.sgdm_update(w, g, state, 0.1, 0.9)
                else:
                    w, state = O.rmsprop_update(w, g, state, 0.01, 0.9)
                close(w, d["%s%d" % (key, i + 1)])


def test_input_pipeline_golden(golden):
    """uint8 NHWC -> fp32 NCHW - 128 (+ mixup) against the live reference's ImagePreprocessor / loader arithmetic."""
    d = golden("input_pipeline")
    assert np.array_equal(O.input_u8_nhwc(d["img_a"]), d["Xa"])
    close(O.input_u8_nhwc(d["img_a"], d["img_b"], float(d["lam"])), d["X_mixed"], rtol=1e-6, atol=1e-5)


def test_synthetic_batches_are_the_oracle_of_their_uint8_form():
    """bench.py's device-resident fp32 batches and its uint8 host batches (e2e) describe the same images."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dorknet_b200 import workloads as W
    for mix in (False, True):
        raw = W.synthetic_batch_u8(3, 3, 11, 7, seed=4, mixup=mix)
        X, y, Y = W.synthetic_batch(3, 3, 11, 7, seed=4, mixup=mix)
        close(X, O.input_u8_nhwc(raw["img"], raw.get("img_b"), raw.get("lam", 0.0)), rtol=1e-6, atol=1e-5)
        assert np.array_equal(Y, raw["Y"]) and np.array_equal(y, raw["y"])


def test_mixup_identity():
    g = np.random.default_rng(0)
    a, b = g.standard_normal((2, 3)), g.standard_normal((2, 3))
    X, y = O.mixup(a, b, a, b, 0.25)
    close(X, 0.25 * b + 0.75 * a)


@pytest.mark.skipif(not refload.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_live_reference_kernels():
    """Restated C kernels vs the reference's compiled Cython, larger random shapes."""
    R = refload.load_reference()
    g = np.random.default_rng(5)
    X = g.standard_normal((3, 6, 13, 11)).astype(np.float32)
    for (k, s) in ((3, 1), (5, 2), (4, 2)):
        P_ref, ohf, owf = R.im2col.im2col_cy(X, k, k, s)
        P = O.im2col(X, k, k, s)
        assert np.array_equal(P, P_ref)
        rows = g.standard_normal(P.shape).astype(np.float32)
        ref = np.asarray(R.im2col.row2im_cy(rows, 3, ohf, owf, k, k, 6, s, 1))
        close(O.row2im(rows, 3, 6, 13, 11, k, k, s, 1), ref)
    m_ref, v_ref = R.batch_norm_stats_cy.channelwise_mean_and_var_4d(X)
    m, v = O.bn_stats(X)
    close(m, m_ref, atol=1e-6); close(v, v_ref, rtol=1e-5)
