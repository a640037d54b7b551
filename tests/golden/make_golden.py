#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference CPU path.

Run in the dev container (where /root/reference exists) after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these fixtures,
produced by importing the reference's own classes (oracle/_ref, compiled from the sources in
/root/reference), are what pins the oracle.  Inputs are seeded; every .npz stores inputs,
parameters, outputs and gradients.  The reference CPU path is run-to-run deterministic.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

OUT = os.path.dirname(os.path.abspath(__file__))
R = load_reference()


def rng(seed):
    return np.random.default_rng(seed)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def conv_case(name, N, C, H, W, F, k, s, p, bias, l2s, seed):
    g = rng(seed)
    L = R.ConvLayer(name, filter_block_shape=(F, C, k, k), stride=s, padding=p, with_bias=bias,
                    weight_regulariser=R.l2(l2s) if l2s else None)
    L.learned_params["weights"] = f32(g.standard_normal((F, C, k, k)) * 0.2)
    if bias:
        L.learned_params["bias"] = f32(g.standard_normal(F))
    X = f32(g.standard_normal((N, C, H, W)))
    Y = f32(L.forward(X))
    dY = f32(g.standard_normal(Y.shape))
    dX = f32(L.backward(dY))
    out = dict(X=X, W=L.learned_params["weights"], Y=Y, dY=dY, dX=dX, dW=f32(L.grads["weights"]),
               patches=f32(L.patches), meta=np.array([N, C, H, W, F, k, s, p, int(bias)]), l2=np.float32(l2s))
    if bias:
        out.update(b=L.learned_params["bias"], db=f32(L.grads["bias"]))
    save(name, **out)


def pointwise_case(name, N, C, H, W, F, s, bias, l2s, seed):
    g = rng(seed)
    L = R.PointwiseConvLayer(name, stride=s, filter_block_shape=(F, C), with_bias=bias,
                             weight_regulariser=R.l2(l2s) if l2s else None)
    L.learned_params["weights"] = f32(g.standard_normal((F, C)) * 0.3)
    if bias:
        L.learned_params["bias"] = f32(g.standard_normal(F))
    X = f32(g.standard_normal((N, C, H, W)))
    Y = f32(L.forward(X))
    dY = f32(g.standard_normal(Y.shape))
    dX = f32(L.backward(dY))
    out = dict(X=X, W=L.learned_params["weights"], Y=Y, dY=dY, dX=dX, dW=f32(L.grads["weights"]),
               meta=np.array([N, C, H, W, F, s, int(bias)]), l2=np.float32(l2s))
    if bias:
        out.update(b=L.learned_params["bias"], db=f32(L.grads["bias"]))
    save(name, **out)


def depthwise_case(name, N, C, H, W, k, s, p, bias, seed):
    g = rng(seed)
    L = R.DepthwiseConvLayer(name, filter_block_shape=(C, k, k), stride=s, padding=p, with_bias=bias)
    L.learned_params["weights"] = f32(g.standard_normal((C, k, k)) * 0.3)
    if bias:
        L.learned_params["bias"] = f32(g.standard_normal(C))
    X = f32(g.standard_normal((N, C, H, W)))
    Y = f32(L.forward(X))
    dY = f32(g.standard_normal(Y.shape))
    dX = f32(L.backward(dY))
    out = dict(X=X, W=L.learned_params["weights"], Y=Y, dY=dY, dX=dX, dW=f32(L.grads["weights"]),
               meta=np.array([N, C, H, W, k, s, p, int(bias)]))
    if bias:
        out.update(b=L.learned_params["bias"], db=f32(L.grads["bias"]))
    save(name, **out)


def bn_case(name, shape, seed):
    g = rng(seed)
    dim = len(shape)
    C = shape[1]
    L = R.BatchNormLayer(name, input_dimension=dim, incoming_chans=C)
    # The reference's CPU branch of BatchNorm backward is 4-D only (einsum "ijkl",
    # layers/batch_norm.py:145,162); for 2-D inputs only its `is_on_gpu` branch (xp.sum over
    # axis 0, :147,164) works, and it runs fine on NumPy arrays -- use that branch for 2-D.
    L.is_on_gpu = (dim == 2)
    gam = f32(1.0 + 0.3 * g.standard_normal(C))
    bet = f32(0.2 * g.standard_normal(C))
    if dim == 4:
        gam, bet = gam[None, :, None, None], bet[None, :, None, None]
    L.learned_params["gamma"], L.learned_params["beta"] = gam, bet
    X1 = f32(g.standard_normal(shape) * 2.0 + 0.7)
    X2 = f32(g.standard_normal(shape) * 0.5 - 1.1)
    Y1 = f32(L.forward(X1))
    rm1, rs1 = f32(L.non_learned_params["running_mean"]), f32(L.non_learned_params["running_std"])
    dY1 = f32(g.standard_normal(shape))
    dX1 = f32(L.backward(dY1))
    dg1, db1 = f32(L.grads["gamma"]), f32(L.grads["beta"])
    Y2 = f32(L.forward(X2))
    rm2, rs2 = f32(L.non_learned_params["running_mean"]), f32(L.non_learned_params["running_std"])
    Yt = f32(L.forward(X1, test_mode=True))
    save(name, X1=X1, X2=X2, gamma=gam, beta=bet, Y1=Y1, rm1=rm1, rs1=rs1, dY1=dY1, dX1=dX1,
         dgamma1=dg1, dbeta1=db1, Y2=Y2, rm2=rm2, rs2=rs2, Ytest=Yt)


def relu_case(name, shape, seed):
    g = rng(seed)
    X = f32(g.standard_normal(shape))
    X.flat[::7] = 0.0  # exact zeros: gradient must be 0 there
    L = R.ReLu(name)
    Y = f32(L.forward(X))
    mask = f32(L.positive_locs)
    dY = f32(g.standard_normal(shape))
    dX = f32(L.backward(dY))
    Yt = f32(L.forward(X, test_mode=True))
    save(name, X=X, Y=Y, mask=mask, dY=dY, dX=dX, Ytest=Yt)


def pool_cases():
    g = rng(40)
    X = f32(g.standard_normal((2, 3, 8, 6)))
    L = R.GlobalAveragePoolingLayer("gap")
    Y = f32(L.forward(X))
    dY = f32(g.standard_normal(Y.shape))
    dX = f32(L.backward(dY))
    save("gap", X=X, Y=Y, dY=dY, dX=dX)
    # max pool: includes the SURVEY A.8 tie pattern in plane (0,0)
    Xm = f32(g.integers(-3, 4, size=(2, 3, 8, 12)))  # many ties
    Xm[0, 0, :4, :4] = np.array([[1, 1, 2, 2], [1, 1, 2, 3], [5, 4, 0, 0], [4, 5, 0, -1]], np.float32)
    for s in (2, 4):
        M = R.MaxPoolLayer("mp", None, stride=s)
        Yp = f32(M.forward(Xm))
        mask = np.asarray(M.max_locations).astype(np.int32)
        dYp = f32(g.standard_normal(Yp.shape))
        dXp = f32(M.backward(dYp))
        Yt = f32(M.forward(Xm, test_mode=True))
        save("maxpool_s%d" % s, X=Xm, Y=Yp, mask=mask, dY=dYp, dX=dXp, Ytest=Yt)


def dense_loss_cases():
    g = rng(50)
    L = R.DenseLayer("d", incoming_chans=12, output_dim=7, with_bias=True, weight_regulariser=R.l2(0.01))
    L.learned_params["weights"] = f32(g.standard_normal((12, 7)) * 0.4)
    L.learned_params["bias"] = f32(g.standard_normal(7))
    X = f32(g.standard_normal((5, 12)))
    Y = f32(L.forward(X))
    dY = f32(g.standard_normal(Y.shape))
    dX = f32(L.backward(dY))
    save("dense", X=X, W=L.learned_params["weights"], b=L.learned_params["bias"], Y=Y, dY=dY, dX=dX,
         dW=f32(L.grads["weights"]), db=f32(L.grads["bias"]), l2=np.float32(0.01),
         reg=np.float32(L.regulariser_forward()))
    S = R.SoftmaxWithCrossEntropy("s")
    logits = f32(g.standard_normal((6, 9)) * 2)
    hard = np.eye(9, dtype=np.float32)[g.integers(0, 9, 6)]
    soft = f32(0.8 * hard + 0.2 * np.eye(9, dtype=np.float32)[g.integers(0, 9, 6)])  # mixup-style labels
    for nm, y in (("softmax_hard", hard), ("softmax_soft", soft)):
        loss, p = S.forward(logits, y)
        d = f32(S.backward())
        _, pt = S.forward(logits, None, test_mode=True)
        save(nm, X=logits, y=y, loss=np.float32(loss), p=f32(p), dX=d, ptest=f32(pt))


def optimiser_cases():
    g = rng(60)

    class FakeLayer:
        def __init__(self):
            self.learned_params = {"weights": f32(g.standard_normal((4, 5))), "bias": f32(g.standard_normal(5))}
            self.grads = {k: np.zeros_like(v) for k, v in self.learned_params.items()}

    class FakeNet:
        pass

    grads_seq = [{"weights": f32(g.standard_normal((4, 5))), "bias": f32(g.standard_normal(5))} for _ in range(3)]
    for nm, mk in (("opt_sgd", lambda n: R.SGD(n, 0.1)), ("opt_sgdm", lambda n: R.SGDMomentum(n, 0.1, 0.9)),
                   ("opt_rmsprop", lambda n: R.RMSProp(n, 0.01, 0.9))):
        net = FakeNet()
        lay = FakeLayer()
        w0 = {k: v.copy() for k, v in lay.learned_params.items()}
        net.layers = [lay]
        opt = mk(net)
        out = {"w0": w0["weights"], "b0": w0["bias"]}
        for i, gr in enumerate(grads_seq):
            lay.grads = {k: v.copy() for k, v in gr.items()}
            opt.update_weights()
            out["gw%d" % i], out["gb%d" % i] = gr["weights"], gr["bias"]
            out["w%d" % (i + 1)] = f32(lay.learned_params["weights"]).copy()
            out["b%d" % (i + 1)] = f32(lay.learned_params["bias"]).copy()
        save(nm, **out)


from net_defs import build_small_net, iter_param_layers  # noqa: E402


def _tf32(a, mode):
    """fp32 -> TF32 (10-bit mantissa) as the tensor core sees its operands: "rz" drops the 13 low mantissa bits,
    "rn" rounds to nearest (ties away from zero)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    if mode == "rn":
        u = u + np.uint32(0x1000)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


class _Tf32Numpy:
    """numpy / cupy-stub stand-in whose dot() feeds TF32-rounded operands to the fp32 GEMM: patched into the
    reference's convolution / pointwise modules to model the sm_100a tensor-core path (kind::tf32)."""

    def __init__(self, mode):
        self._mode = mode

    def __getattr__(self, name):
        return getattr(np, name)

    def dot(self, a, b):
        return np.dot(_tf32(a, self._mode), _tf32(b, self._mode))

    def get_array_module(self, *a):
        return self

    def asarray(self, a, *k, **kw):
        return np.asarray(a, *k, **kw)


def net_case(tf32_mode=None):
    """3 SGDMomentum steps on the miniature net at 33x33 (the reference's odd-size geometry:
    33 -(5x5 s2 p1)-> 16 -(pw s2)-> 8 -> 8 -(dw s2)-> 4).  With tf32_mode the reference's conv / pointwise GEMMs
    (and only those: cp.dot / np.dot / xp.dot in layers/convolution.py:75-120, pointwise_convolution.py:51-65) see
    TF32-rounded operands -- the arithmetic model of the tcgen05 kind::tf32 path."""
    import importlib
    patched = []
    if tf32_mode:
        proxy = _Tf32Numpy(tf32_mode)
        for mn in ("layers.convolution", "layers.pointwise_convolution"):
            mod = importlib.import_module(mn)
            patched.append((mod, mod.cp, mod.np))
            mod.cp = proxy
            mod.np = proxy
    try:
        _net_case_body("mini_net" if not tf32_mode else "mini_net_tf32" + tf32_mode)
    finally:
        for mod, cp_, np_ in patched:
            mod.cp, mod.np = cp_, np_


def _net_case_body(name):
    g = rng(70)
    net = build_small_net(R, seed=123)
    out = {}
    for l in iter_param_layers(net):
        for k, v in l.learned_params.items():
            out["init/%s/%s" % (l.layer_name, k)] = f32(v).copy()
    opt = R.SGDMomentum(net, 0.02, 0.9)
    X = f32(g.uniform(0, 255, (8, 3, 33, 33)) - 128.0)
    y = np.eye(5, dtype=np.float32)[g.integers(0, 5, 8)]
    out["X"], out["y"] = X, y
    losses = []
    for step in range(3):
        loss, scores = net.forward(X, y)
        losses.append(loss)
        net.backward()
        if step == 0:
            out["scores0"] = f32(scores)
            for l in iter_param_layers(net):
                for k, v in l.grads.items():
                    out["grad0/%s/%s" % (l.layer_name, k)] = f32(v).copy()
        opt.update_weights()
    out["losses"] = np.array(losses, np.float64)
    for l in iter_param_layers(net):
        for k, v in l.learned_params.items():
            out["final/%s/%s" % (l.layer_name, k)] = f32(v).copy()
    _, st = net.forward(X, None, test_mode=True)
    out["scores_test"] = f32(st)
    _, lt = net.forward(X, None, test_mode=True, terminal_layer_name="dense1")
    out["logits_test"] = f32(lt)
    save(name, **out)


def input_case():
    """The reference loader's per-image arithmetic after decoding (data_loading/image_preprocessor.py:16-39 with
    crop_mode=None on images that already have the target size: cv2.resize to the same size is the identity, then
    astype(float32).transpose(2,0,1) - 128) and its mixup (image_data_loader.py:100-110), from the live reference."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_image_preprocessor", "/root/reference/data_loading/image_preprocessor.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = rng(90)
    N, S, C = 3, 17, 3
    pre = mod.ImagePreprocessor(image_size=(S, S), crop_mode=None)
    img_a = g.integers(0, 256, size=(N, S, S, C), dtype=np.uint8)
    img_b = g.integers(0, 256, size=(N, S, S, C), dtype=np.uint8)
    Xa = np.stack([pre.preprocess_image(im.copy()) for im in img_a], axis=0)
    Xb = np.stack([pre.preprocess_image(im.copy()) for im in img_b], axis=0)
    lam = np.float32(0.2173)
    mixed = lam * Xb + (1 - lam) * Xa  # image_data_loader.py:105 with mixup_prop = lam
    save("input_pipeline", img_a=img_a, img_b=img_b, Xa=f32(Xa), lam=np.float32(lam), X_mixed=f32(mixed))


def main():
    input_case()
    conv_case("conv_k3s1p1", 2, 3, 9, 10, 4, 3, 1, 1, False, 0.0, 1)
    conv_case("conv_k5s2p1_half", 2, 3, 10, 12, 5, 5, 2, 1, True, 0.01, 2)   # (10+2-5)/2 = 3.5 -> floor
    conv_case("conv_k4s2p1", 2, 4, 8, 8, 6, 4, 2, 1, False, 1e-4, 3)         # MNIST 4x4 s2
    conv_case("conv_k3s1p0", 1, 2, 6, 7, 3, 3, 1, 0, True, 0.0, 4)
    pointwise_case("pw_s1", 2, 6, 5, 7, 9, 1, True, 0.01, 10)
    pointwise_case("pw_s2_even", 2, 4, 6, 8, 5, 2, False, 1e-3, 11)
    pointwise_case("pw_s2_odd", 2, 4, 7, 9, 5, 2, False, 0.0, 12)            # dX is 8x10, not 7x9
    depthwise_case("dw_k3s1p1", 2, 5, 7, 9, 3, 1, 1, False, 20)
    depthwise_case("dw_k3s2p1_half", 2, 4, 8, 10, 3, 2, 1, True, 21)         # (8+2-3)/2 = 3.5
    depthwise_case("dw_k3s2p1_int", 2, 4, 7, 9, 3, 2, 1, False, 22)
    depthwise_case("dw_k5s1p2", 1, 3, 9, 8, 5, 1, 2, False, 23)
    bn_case("bn_4d", (3, 5, 6, 7), 30)
    bn_case("bn_2d", (16, 6), 31)
    relu_case("relu_4d", (2, 3, 5, 7), 35)
    relu_case("relu_2d", (6, 11), 36)
    pool_cases()
    dense_loss_cases()
    optimiser_cases()
    net_case()
    net_case("rz")
    net_case("rn")


if __name__ == "__main__":
    main()
