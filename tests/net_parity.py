"""One training step of a BASELINE network on the CUDA path against the golden vectors of the live reference
(tests/golden/make_golden_r18.py): shared by tests/test_gpu_network.py (in process, the repo's own container) and by
the same test run in a subprocess on top of the REFERENCE'S UNCHANGED container:

    python tests/net_parity.py --net r18 --container ref      # oracle/_ref/network/feed_forward_network (compiled
                                                             # from /root/reference/network/feed_forward_network.py,
                                                             # not one line changed) driving dorknet_b200's layers
    python tests/net_parity.py --net r18 --container ours

What is compared with what.  The forward pass (loss, scores, BatchNorm running statistics) is compared with the
reference directly.  The reference's fp32 GRADIENTS are themselves up to 9.5 % away from the exact gradients of the same
step (float64 evaluation, tests/golden/fp64_net.py; measured per tensor by tests/golden/make_golden_r18.py and stored
as referr/*: its per-channel fp32 sums lose the terms that cancel), so every gradient is held to the EXACT gradient
(grad64/*) with the gates below, and in addition must be at least as close to it as the reference is (x1.5 + gate
floor).  The distance to the reference's own gradient is printed next to it.

Gates (normalised max-abs error max|a-b| / max|b| per tensor; gradients whose exact value cancels to ~0 --
dbeta of a BatchNorm that feeds a linear layer + BatchNorm -- are normalised by 1 % of the largest gradient of the same
kind in the net instead of by their own noise).  The product default runs conv / pointwise / dense GEMMs on tcgen05
kind::tf32 (operands truncated to 10 mantissa bits, fp32 accumulation), everything else in fp32.

TF32 and these networks.  At these configurations (random init, batch 8 / 16) the parameter gradients are sums of
N*H*W terms that almost cancel, and every ReLU whose input changes sign under a perturbation adds or removes a whole
term: the gradient is not a smooth function of the arithmetic.  Measured on the CPU with the exact float64 evaluation
and ONLY the GEMM operands rounded (tests/golden/fp64_net.py TF32_OPERANDS; MNIST net, global relative L2 error of the
gradient vs exact / worst per-tensor max-abs): operands kept to 19 mantissa bits 1.4e-6 / 3.5e-6, 16 bits 1.9e-3 /
1.5e-2, 13 bits 9e-3 / 4.8e-2, 10 bits (TF32, truncated or rounded alike) 2.7e-2 / 1.0e-1; ResNet-18-depsep: 1.5e-1
global.  A 1e-6 relative perturbation of the input moves the reference's own gradients by 0.9 % (median per tensor).
That is a property of the network, not of a kernel, so the TF32 product path is held
  (a) at network level to the forward quantities tightly and to the gradient as a direction (global cosine, relative
      L2 per tensor) with the gates below, and
  (b) LAYER BY LAYER INSIDE THE REAL STEP, tightly: every conv / pointwise / dense GEMM the step launched is recomputed
      in fp32 (SIMT backend) from the very operands the TF32 kernel consumed -- forward and dgrad <= 2e-3, wgrad <= 5e-3
      (layerwise_check), i.e. the per-layer TF32 gates of SURVEY A.12 at the real shapes on the real data.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# network-level gates, product default backend (TF32 tcgen05 GEMMs + fused cluster BatchNorm)
TOL = {
    "loss": 2e-4,         # relative, loss + l2 terms, vs the reference
    "scores": 5e-3,       # softmax probabilities of the batch, vs the reference
    "grad_gemm": 0.35,    # per tensor, max-abs vs the exact gradient (see "TF32 and these networks": measured 0.25)
    "grad_dw": 0.35,
    "grad_bn": 0.35,
    "grad_l2": 0.30,      # per tensor, relative L2 error vs the exact gradient
    "grad_cos": 2e-2,     # 1 - cosine between the whole gradient (all tensors) and the exact one
    "running": 1.5e-2,    # BatchNorm batch mean / std of the step (running statistics after the first batch), vs float64
    # test-mode scores after the SGDMomentum update, vs the reference.  They inherit the gradient's TF32 noise through the
    # update (see above).  Measured on ResNet-18-depsep: 1.4e-2 with every BatchNorm as its own kernel, 2.4e-2 with the
    # depthwise BatchNorms folded into the pointwise GEMMs (bn_fold.cu: the GEMMs then truncate un-centred activations,
    # whose TF32 rounding noise is larger by sqrt(1 + mean^2/var) -- the output MEAN is kept exact, the per-layer gates
    # below are unchanged)
    "scores_test": 3e-2,
    "layer_fwd": 2e-3, "layer_dgrad": 2e-3, "layer_wgrad": 5e-3,  # (b): each GEMM of the step vs fp32 on the same operands
}
# fp32 SIMT GEMM backend (GPU-side cross-check): fp32 against float64.  Measured: MNIST 1.9e-6 from the exact gradient
# (the reference: 6.0e-3), ResNet-18-depsep 9.7e-3 (the reference: 9.4e-2) -- cancellation again, ten times less of it
TOL_FP32 = {"loss": 2e-6, "scores": 2e-5, "grad_gemm": 2e-2, "grad_dw": 2e-2, "grad_bn": 2e-2, "grad_l2": 2e-2,
            "grad_cos": 1e-4, "running": 1e-4, "scores_test": 5e-3}


def nerr(a, b, floor=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor, 1e-30))


def _gemm_layers(net):
    from dorknet_b200 import workloads
    return [l for l in workloads.iter_param_layers(net)
            if type(l).__name__ in ("ConvLayer", "PointwiseConvLayer", "DenseLayer")]


def _capture_upstream(net):
    """Wrap backward() of every GEMM layer to remember the upstream-gradient array it was called with (every buffer on
    this path is persistent and owned by one layer, so it still holds the step's values afterwards)."""
    rec = {}
    for l in _gemm_layers(net):
        orig = l.backward

        def wrapped(dY, *a, _l=l, _orig=orig, **kw):
            rec[id(_l)] = (_l, dY)
            return _orig(dY, *a, **kw)
        l.backward = wrapped
    return rec


def layerwise_check(captured, check):
    """(b) of the module docstring: recompute every GEMM of the step that just ran with the fp32 SIMT kernels from the
    operands the tcgen05 kernels consumed (layer._x / downstream_X, the captured upstream gradient, the weights) through
    the C ABI into scratch buffers, and compare with what the step produced."""
    from dorknet_b200 import api, runtime, empty, asarray
    st = runtime.stream()
    api.dk_set_gemm_backend(1)
    try:
        for l, dY in captured.values():
            t, nm = type(l).__name__, l.layer_name
            dY = asarray(dY)
            w = l._param("weights")
            bias = l._param("bias").ptr if getattr(l, "with_bias", False) else None
            dwt, l2s = empty(w.shape), l._l2_strength()
            dbt = empty((w.shape[0],)) if bias is not None else None
            if t == "PointwiseConvLayer" and getattr(l, "_folded_bn", None) is not None:
                # BatchNorm folded into this layer (bn_fold.cu): the GEMMs consumed the BatchNorm's INPUT, the folded
                # weights / bias, and produced the raw wgrad G and the BatchNorm's input gradient
                x, bn = l._x, l._folded_bn
                N, C, H, W = x.shape
                F = l.num_filters
                wf, bf, y = l._bufs["w_fold"], l._bufs["b_fold"], l._bufs["y"]
                ws, wsn = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, H, W, F, 1))
                yt = empty(y.shape)
                api.dk_pwconv_fwd(x.ptr, wf.ptr, bf.ptr, yt.ptr, N, C, H, W, F, 1, ws, wsn, st)
                check("layer_fwd", "fwd " + nm + " (folded BN)", nerr(y.get(), yt.get()))
                api.dk_pwconv_wgrad(dY.ptr, x.ptr, w.ptr, dwt.ptr, None, 0.0, N, C, H, W, F, 1, ws, wsn, st)
                check("layer_wgrad", "wgrad " + nm + " (raw, folded BN)", nerr(l._bufs["g_raw"].get(), dwt.get()))
                coef = l._bufs["bn_coef"]
                dx = bn._bufs["dx"]
                dxt = empty(dx.shape)
                api.dk_pwconv_dgrad_affine(dY.ptr, wf.ptr, x.ptr, coef.ptr, coef.ptr + 4 * C, dxt.ptr, N, C, H, W, F, ws, wsn, st)
                check("layer_dgrad", "dgrad " + nm + " (+ BN backward epilogue)", nerr(dx.get(), dxt.get()))
            elif t == "PointwiseConvLayer":
                x, (xh, xw, xs) = l._x, l._xgeom
                N, C = x.shape[0], x.shape[1]
                F = l.num_filters
                y = l._bufs["y"]
                ws, wsn = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, max(xh, y.shape[2] * xs), max(xw, y.shape[3] * xs), F, xs))
                yt = empty(y.shape)
                api.dk_pwconv_fwd(x.ptr, w.ptr, bias, yt.ptr, N, C, xh, xw, F, xs, ws, wsn, st)
                check("layer_fwd", "fwd " + nm, nerr(y.get(), yt.get()))
                api.dk_pwconv_wgrad(dY.ptr, x.ptr, w.ptr, dwt.ptr, dbt.ptr if dbt is not None else None, l2s, N, C, xh, xw, F, xs, ws, wsn, st)
                check("layer_wgrad", "wgrad " + nm, nerr(l.grads["weights"].get(), dwt.get()))
                OH, OW = y.shape[2], y.shape[3]
                if int(l.stride) == 1:
                    dx = l._bufs["dx"]
                elif "dx_sub" in l._bufs:  # stride 2, consumed in compact form (the non-zero entries)
                    dx = l._bufs["dx_sub"]
                else:
                    dx = None
                if dx is not None:
                    dxt = empty(dx.shape)
                    api.dk_pwconv_dgrad(dY.ptr, w.ptr, dxt.ptr, N, C, OH, OW, F, 1, ws, wsn, st)
                    check("layer_dgrad", "dgrad " + nm, nerr(dx.get(), dxt.get()))
            elif t == "ConvLayer":
                x = l._x
                N, C, H, W = x.shape
                F, kh, kw, s, p = l.num_filters, l.f_rows, l.f_cols, int(l.stride), int(l.padding)
                ws, wsn = runtime.scratch(api.dk_conv2d_ws_bytes(N, C, H, W, F, kh, kw, s, p))
                y = l._bufs["y"]
                yt = empty(y.shape)
                api.dk_conv2d_fwd(x.ptr, w.ptr, bias, yt.ptr, N, C, H, W, F, kh, kw, s, p, ws, wsn, st)
                check("layer_fwd", "fwd " + nm, nerr(y.get(), yt.get()))
                api.dk_conv2d_wgrad(dY.ptr, x.ptr, w.ptr, dwt.ptr, dbt.ptr if dbt is not None else None, l2s, N, C, H, W, F, kh, kw, s, p, ws, wsn, st)
                check("layer_wgrad", "wgrad " + nm, nerr(l.grads["weights"].get(), dwt.get()))
                dxt = empty(x.shape)
                api.dk_conv2d_dgrad(dY.ptr, w.ptr, dxt.ptr, N, C, H, W, F, kh, kw, s, p, ws, wsn, st)
                api.dk_set_gemm_backend(0)
                dx0 = empty(x.shape)
                api.dk_conv2d_dgrad(dY.ptr, w.ptr, dx0.ptr, N, C, H, W, F, kh, kw, s, p, ws, wsn, st)
                api.dk_set_gemm_backend(1)
                check("layer_dgrad", "dgrad " + nm, nerr(dx0.get(), dxt.get()))
            else:  # DenseLayer
                x = l.downstream_X
                B, D = x.shape
                K = l.output_dim
                ws, wsn = runtime.scratch(api.dk_dense_ws_bytes(B, D, K))
                y = l._bufs["y"]
                yt, dxt = empty(y.shape), empty(x.shape)
                api.dk_dense_fwd(x.ptr, w.ptr, bias, yt.ptr, B, D, K, ws, wsn, st)
                check("layer_fwd", "fwd " + nm, nerr(y.get(), yt.get()))
                dbd = empty((K,)) if bias is not None else None
                api.dk_dense_bwd(dY.ptr, x.ptr, w.ptr, dxt.ptr, dwt.ptr, dbd.ptr if dbd is not None else None, l2s, B, D, K, ws, wsn, st)
                check("layer_dgrad", "dgrad " + nm, nerr(l._bufs["dx"].get(), dxt.get()))
                check("layer_wgrad", "wgrad " + nm, nerr(l.grads["weights"].get(), dwt.get()))
    finally:
        api.dk_set_gemm_backend(0)


def container_namespace(kind):
    """Layer classes + container: "ours" = dorknet_b200.network.FeedForwardNetwork, "ref" = the reference's own
    compiled container module on top of our layers after dropin.install()."""
    from dorknet_b200 import workloads
    M = workloads.ours()
    if kind == "ref":
        from dorknet_b200 import dropin
        dropin.install()
        ref_dir = os.path.join(ROOT, "oracle", "_ref")
        if not os.path.isdir(os.path.join(ref_dir, "network")):
            raise RuntimeError("oracle/_ref/network is not built (python oracle/build_ref.py where /root/reference exists)")
        sys.path.append(ref_dir)  # appended: `layers`, `optimisers`, `regularisers` are already bound to the product
        import importlib
        mod = importlib.import_module("network.feed_forward_network")
        assert "oracle/_ref/network" in mod.__file__.replace("\\", "/"), mod.__file__
        import dorknet_b200.layers.dense_layer as ours_dense
        assert mod.DenseLayer is ours_dense.DenseLayer, "the reference container did not bind to the CUDA layers"
        M.FeedForwardNetwork = mod.FeedForwardNetwork
    return M


# fp32 SIMT backend WITH the depthwise BatchNorms folded into the pointwise GEMMs (bn_fold.cu).  The fold is the same
# mathematics in another operation order: y = W'x + b' cancels W'.mean against b' in fp32, which perturbs the activations at
# the 1e-6 level instead of 1e-7 -- and this network turns a 1e-6 perturbation into per-cent changes of the few gradients
# that are sums of nearly cancelling terms (module docstring; the reference's own fp32 gradients are 9.4e-2 from exact on
# the same tensors: measured 9.4e-2 on res7_dw2_pw/weights for both).  Forward quantities and the direction stay tight.
TOL_FP32_FOLDED = dict(TOL_FP32, grad_gemm=0.12, grad_dw=0.12, grad_bn=0.12, grad_l2=0.05)


def run(net_name="r18", container="ours", backend=0, verbose=True, fold=None):
    """fold: None = product default (fold where the statistics ride on the depthwise kernel); False = every BatchNorm as
    its own layer (the reference's operation order); "always" / True as PointwiseConvLayer.fold_bn_input."""
    from dorknet_b200 import api, workloads, launch_count
    d = np.load(os.path.join(ROOT, "tests", "golden", {"r18": "r18_b8", "mnist": "mnist_b16"}[net_name] + ".npz"))
    M = container_namespace(container)
    api.dk_set_gemm_backend(backend)
    tol = TOL if backend == 0 else (TOL_FP32 if fold is False else TOL_FP32_FOLDED)
    try:
        if net_name == "r18":
            net = workloads.build_resnet18_depsep(M, classes=120, conv0_padding=1, seed=0)
            X, _, Y = workloads.synthetic_batch(8, 3, 225, 120, seed=7)
            lr = 0.05 * 8 / 200.0
        else:
            net = workloads.build_mnist_convnet(M, seed=0)
            X, Y, lr = d["X"], d["Y"], 0.01
        if fold is not None:
            for l in workloads.iter_param_layers(net):
                if hasattr(l, "fold_bn_input"):
                    l.fold_bn_input = fold
        # same seed, same draw order as the reference build: check before comparing anything
        for l in workloads.iter_param_layers(net):
            for k, v in l.learned_params.items():
                s = float(np.sum(np.abs(np.asarray(v, np.float64))))
                assert abs(s - float(d["initsum/%s/%s" % (l.layer_name, k)])) <= 1e-9 * max(s, 1.0), (l.layer_name, k)
        opt = M.SGDMomentum(net, lr, 0.9)
        n0 = launch_count()
        captured = _capture_upstream(net) if backend == 0 else None
        loss, scores = net.forward(X, Y)
        net.backward()
        worst, failures, ref_worst, vs_ref, dots = {}, [], [0.0], [0.0], [0.0, 0.0, 0.0]

        def check(cat, what, e):
            if e >= worst.get(cat, (-1.0, ""))[0]:
                worst[cat] = (e, what)
            if not e <= tol[cat]:
                failures.append("%s: %s error %.3e > %.1e" % (what, cat, e, tol[cat]))

        check("loss", "loss", abs(float(loss) - float(d["loss"])) / abs(float(d["loss"])))
        check("scores", "scores", nerr(scores.get(), d["scores"]))
        # gradients that cancel to ~0 in exact arithmetic (dbeta / dgamma-free terms of a BN feeding a BN, bias-like sums)
        # are compared against the scale of the same KIND of gradient in the whole net, not against their own noise
        kinds = {}
        for l in workloads.iter_param_layers(net):
            for k in l.grads.keys():
                g = d["grad64/%s/%s" % (l.layer_name, k)]
                kk = (type(l).__name__, k)
                kinds[kk] = max(kinds.get(kk, 0.0), float(np.max(np.abs(g))))
        for l in workloads.iter_param_layers(net):
            tname = type(l).__name__
            cat = {"BatchNormLayer": "grad_bn", "DepthwiseConvLayer": "grad_dw"}.get(tname, "grad_gemm")
            for k in l.grads.keys():
                nm = "%s/%s" % (l.layer_name, k)
                g, gref = d["grad64/" + nm], d["grad/" + nm]
                floor = 1e-2 * kinds[(tname, k)]
                mine = l.grads[k].get()
                e = nerr(mine, g, floor)
                check(cat, "grad " + nm, e)
                eref = nerr(gref, g, floor)  # the reference's own distance to the exact gradient
                a64, b64 = np.asarray(mine, np.float64).ravel(), np.asarray(g, np.float64).ravel()
                if float(np.max(np.abs(b64))) >= floor:  # (tensors that are pure cancellation noise carry no direction)
                    check("grad_l2", "grad " + nm, float(np.linalg.norm(a64 - b64) / np.linalg.norm(b64)))
                    dots[0] += float(a64 @ b64); dots[1] += float(a64 @ a64); dots[2] += float(b64 @ b64)
                ref_worst[0] = max(ref_worst[0], eref)
                vs_ref[0] = max(vs_ref[0], nerr(mine, gref, floor))
                if backend != 0 and not e <= 1.5 * eref + tol[cat]:
                    failures.append("grad %s: %.3e from the exact gradient, the reference is at %.3e" % (nm, e, eref))
            nl = getattr(l, "non_learned_params", None)
            if nl and "rm/%s" % l.layer_name in d.files:
                # first training batch: running statistics are assigned the batch mean / std (batch_norm.py:76-89)
                check("running", "running_mean %s" % l.layer_name,
                      nerr(np.asarray(nl["running_mean"].get()).reshape(-1), d["mean64/%s" % l.layer_name], 1e-3))
                check("running", "running_std %s" % l.layer_name,
                      nerr(np.asarray(nl["running_std"].get()).reshape(-1), d["std64/%s" % l.layer_name]))
        check("grad_cos", "whole gradient", 1.0 - dots[0] / np.sqrt(dots[1] * dots[2]))
        if captured is not None:
            layerwise_check(captured, check)
        opt.update_weights()
        _, st = net.forward(X, None, test_mode=True)
        check("scores_test", "scores_test", nerr(st.get(), d["scores_test"]))
        launches = launch_count() - n0
        assert launches > 0, "no CUDA kernel was launched"
        from dorknet_b200._lib import gemm_call_counts
        tc, simt = gemm_call_counts()
        if verbose:
            print("net=%s container=%s (%s) backend=%d: %d kernel launches, GEMM calls tcgen05=%d simt=%d" % (
                net_name, container, type(net).__module__, backend, launches, tc, simt))
            for cat in sorted(worst):
                print("  %-12s worst %.3e (gate %.1e) at %s" % (cat, worst[cat][0], tol[cat], worst[cat][1]))
            print("  gradients: the reference's own worst distance to the exact gradient %.3e; ours to the reference's %.3e"
                  % (ref_worst[0], vs_ref[0]))
        assert not failures, "\n".join(failures)
        return worst
    finally:
        api.dk_set_gemm_backend(0)


def autograph_check(M, steps=6):
    """dorknet_b200.dropin.accelerate on container namespace M (the reference's own FeedForwardNetwork for --container
    ref): the same loop as an eager twin -- forward / backward / update_weights called one after the other, the loss kept
    as a running average like examples/imagenet_dogs_225_resnet_18_depsep.py:222-226, a test-mode forward in between, a
    smaller last batch -- must produce bit-identical weights, and really replay graphs."""
    from dorknet_b200 import dropin, workloads
    rng = np.random.default_rng(5)
    batches = [(rng.standard_normal((16 if i != 4 else 8, 1, 28, 28)).astype(np.float32), None) for i in range(steps)]
    batches = [(x, np.eye(10, dtype=np.float32)[rng.integers(0, 10, x.shape[0])]) for x, _ in batches]
    results = []
    for accelerated in (False, True):
        net = workloads.build_mnist_convnet(M, seed=3)
        opt = M.SGDMomentum(net, 0.01, 0.9)
        ag = dropin.accelerate(net, opt, warmup=2) if accelerated else None
        running, test_scores = 0.0, None
        for i, (x, y) in enumerate(batches):
            loss, _ = net.forward(x, y)
            net.backward()
            opt.update_weights()
            running = 0.9 * running + 0.1 * loss
            if i == 3:
                _, sc = net.forward(batches[0][0], None, test_mode=True)
                test_scores = sc.get().copy()
        weights = [l.learned_params[k].get().copy() for l in workloads.iter_param_layers(net) for k in sorted(l.learned_params)]
        results.append((float(running), test_scores, weights, ag.num_graphs if ag else 0))
    # and the container's own checkpoint round trip on top of the CUDA layers (feed_forward_network.py:90-139; `h5py` is
    # dorknet_b200.minih5 under dropin.install() when the real package is missing): save, rebuild from JSON + HDF5, same scores
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        h5, js = os.path.join(tmp, "net.h5"), os.path.join(tmp, "net.json")
        net.save_weights_to_h5(h5)
        net.save_layer_structure_to_json(js)
        again = M.FeedForwardNetwork("empty")
        again.load_network_from_json_and_h5(js, h5)
        _, a = net.forward(batches[0][0], None, test_mode=True)
        a = a.get().copy()
        _, b = again.forward(batches[0][0], None, test_mode=True)
        assert np.array_equal(a, b.get()), "checkpoint round trip changed the test-mode scores"
    (r0, t0, w0, _), (r1, t1, w1, ng) = results
    assert ng >= 3, "no CUDA graph was captured (%d)" % ng
    assert r0 == r1, (r0, r1)
    assert np.array_equal(t0, t1)
    for a, b in zip(w0, w1):
        assert np.array_equal(a, b)
    return ng


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="r18", choices=["r18", "mnist"])
    ap.add_argument("--container", default="ours", choices=["ours", "ref"])
    ap.add_argument("--backend", type=int, default=0)
    ap.add_argument("--fold", default="default", choices=["default", "off", "always"])
    a = ap.parse_args()
    run(a.net, a.container, a.backend, fold={"default": None, "off": False, "always": "always"}[a.fold])
    if a.net == "mnist":
        print("autograph ok: %d graphs behind the unchanged loop (%s container)" % (autograph_check(container_namespace(a.container)), a.container))
    print("net_parity ok")
