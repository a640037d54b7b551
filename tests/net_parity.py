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
kind::tf32 (operands truncated to 10 mantissa bits, fp32 accumulation), everything else in fp32:
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
    "grad_gemm": 1.5e-2,  # conv / pointwise / dense weight gradients vs the exact gradient
    "grad_dw": 1.5e-2,    # depthwise weight gradients (fp32 kernels fed by TF32-perturbed activations / gradients)
    "grad_bn": 1.5e-2,    # gamma / beta gradients
    "running": 1e-3,      # BatchNorm batch mean / std of the step (running statistics after the first batch), vs float64
    "scores_test": 2e-2,  # test-mode scores after the SGDMomentum update, vs the reference
}
# fp32 SIMT GEMM backend (GPU-side cross-check): fp32 against float64
TOL_FP32 = {"loss": 2e-6, "scores": 2e-5, "grad_gemm": 2e-4, "grad_dw": 2e-4, "grad_bn": 2e-4, "running": 2e-5,
            "scores_test": 5e-3}


def nerr(a, b, floor=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor, 1e-30))


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


def run(net_name="r18", container="ours", backend=0, verbose=True):
    from dorknet_b200 import api, workloads, launch_count
    d = np.load(os.path.join(ROOT, "tests", "golden", {"r18": "r18_b8", "mnist": "mnist_b16"}[net_name] + ".npz"))
    M = container_namespace(container)
    api.dk_set_gemm_backend(backend)
    tol = TOL if backend == 0 else TOL_FP32
    try:
        if net_name == "r18":
            net = workloads.build_resnet18_depsep(M, classes=120, conv0_padding=1, seed=0)
            X, _, Y = workloads.synthetic_batch(8, 3, 225, 120, seed=7)
            lr = 0.05 * 8 / 200.0
        else:
            net = workloads.build_mnist_convnet(M, seed=0)
            X, Y, lr = d["X"], d["Y"], 0.01
        # same seed, same draw order as the reference build: check before comparing anything
        for l in workloads.iter_param_layers(net):
            for k, v in l.learned_params.items():
                s = float(np.sum(np.abs(np.asarray(v, np.float64))))
                assert abs(s - float(d["initsum/%s/%s" % (l.layer_name, k)])) <= 1e-9 * max(s, 1.0), (l.layer_name, k)
        opt = M.SGDMomentum(net, lr, 0.9)
        n0 = launch_count()
        loss, scores = net.forward(X, Y)
        net.backward()
        worst, failures, ref_worst, vs_ref = {}, [], [0.0], [0.0]

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


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="r18", choices=["r18", "mnist"])
    ap.add_argument("--container", default="ours", choices=["ours", "ref"])
    ap.add_argument("--backend", type=int, default=0)
    a = ap.parse_args()
    run(a.net, a.container, a.backend)
    print("net_parity ok")
