"""Diagnostic (run on a B200): per-tensor error of the mini ResNet-depsep training step for each GEMM backend
against the golden vectors of the live reference -- max-abs and relative-L2."""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dorknet_b200 import api, workloads as W
    spec = importlib.util.spec_from_file_location("net_defs", os.path.join(ROOT, "tests", "golden", "net_defs.py"))
    defs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(defs)
    d = np.load(os.path.join(ROOT, "tests", "golden", "mini_net.npz"))
    gold = {k: np.load(os.path.join(ROOT, "tests", "golden", "mini_net%s.npz" % k)) for k in ("", "_tf32rz", "_tf32rn")}
    L = W.ours()
    for backend, mask, rows, bnf in ((1, 0, 1, 0), (1, 0, 1, 1), (0, 0, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (0, 0, 1, 1)):
        api.dk_set_gemm_backend(backend)
        api.dk_tc_debug_set(0, mask)
        api.dk_tc_debug_set(8, rows)  # conv_rows.cu on/off
        api.dk_tc_debug_set(9, bnf)   # bn_fused.cu on/off
        net = defs.build_small_net(L, seed=123)
        for l in defs.iter_param_layers(net):
            for k in list(l.learned_params.keys()):
                l.learned_params[k] = d["init/%s/%s" % (l.layer_name, k)].copy()
        loss, scores = net.forward(d["X"], d["y"])
        net.backward()
        print("== backend %d tc-disable-mask %d conv_rows %d bn_fused %d  X%s loss %.7f (ref %.7f)" % (
            backend, mask, rows, bnf, d["X"].shape, float(loss), float(d["losses"][0])))
        for gname, gd in gold.items():
            gscale = max(float(np.max(np.abs(gd[k]))) for k in gd.files if k.startswith("grad0/"))
            worst = []
            for l in defs.iter_param_layers(net):
                for k in l.grads.keys():
                    a = l.grads[k].get().astype(np.float64)
                    b = gd["grad0/%s/%s" % (l.layer_name, k)].astype(np.float64)
                    worst.append((np.max(np.abs(a - b)) / gscale, l.layer_name + "/" + k))
            worst.sort(reverse=True)
            print("   vs golden%-8s loss %.8f  worst grad errors / max-grad-of-net: %s" % (
                gname or "(fp32)", float(gd["losses"][0]), ", ".join("%s %.2e" % (n, e) for e, n in worst[:3])))
    api.dk_set_gemm_backend(0)
    api.dk_tc_debug_set(0, 0)
    api.dk_tc_debug_set(8, 1)
    api.dk_tc_debug_set(9, 1)


if __name__ == "__main__" and "--layerwise" not in sys.argv:
    main()


def layerwise(backend=0):
    """Per-layer outputs / input gradients of the mini net with bn_fused off vs on (same GEMM backend): the first layer
    whose difference is not rounding noise names the kernel that disagrees."""
    from dorknet_b200 import api, workloads as W
    spec = importlib.util.spec_from_file_location("net_defs", os.path.join(ROOT, "tests", "golden", "net_defs.py"))
    defs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(defs)
    d = np.load(os.path.join(ROOT, "tests", "golden", "mini_net.npz"))
    L = W.ours()
    rec = {}
    for bnf in (0, 1):
        api.dk_set_gemm_backend(backend)
        api.dk_tc_debug_set(9, bnf)
        net = defs.build_small_net(L, seed=123)
        for l in defs.iter_param_layers(net):
            for k in list(l.learned_params.keys()):
                l.learned_params[k] = d["init/%s/%s" % (l.layer_name, k)].copy()
        log = []

        def wrap(layer):
            f0, b0 = layer.forward, layer.backward

            def fwd(X, *a, **k):
                xin = X.get().copy() if hasattr(X, "get") else np.asarray(X)
                y = f0(X, *a, **k)
                yh = y.get().copy()
                log.append(("fwd " + layer.layer_name, yh))
                if type(layer).__name__ == "BatchNormLayer":
                    sv = layer._bufs["saved"].get()
                    x64 = xin.astype(np.float64)
                    m, v = x64.mean(axis=(0, 2, 3)), x64.var(axis=(0, 2, 3))
                    ref = (x64 - m[None, :, None, None]) / np.sqrt(v + 1e-5)[None, :, None, None]
                    print("      [bn_fused %d] %-14s saved mean err %.2e  invstd rel err %.2e  y err vs float64(x_in) %.2e" % (
                        bnf, layer.layer_name, np.max(np.abs(sv[0] - m)) / max(np.max(np.abs(m)), 1e-30),
                        np.max(np.abs(sv[1] * np.sqrt(v + 1e-5) - 1)), np.max(np.abs(yh - ref)) / np.max(np.abs(ref))))
                return y

            def bwd(g, *a, **k):
                r = b0(g, *a, **k)
                if r is not None:
                    log.append(("bwd " + layer.layer_name, r.get().copy()))
                return r
            layer.forward, layer.backward = fwd, bwd
        for l in net.layers:
            if hasattr(l, "layer_list"):
                for m in l.layer_list:
                    wrap(m)
                if l.skip_projection is not None:
                    wrap(l.skip_projection)
            wrap(l)
        net.forward(d["X"], d["y"])
        net.backward()
        rec[bnf] = log
    print("== layerwise, backend %d: bn_fused 0 vs 1" % backend)
    for (n0, a), (n1, b) in zip(rec[0], rec[1]):
        assert n0 == n1 and a.shape == b.shape, (n0, n1)
        print("   %-28s max|a| %.3e  max diff %.3e  (rel %.2e)  differing elements %d / %d" % (
            n0, np.max(np.abs(a)), np.max(np.abs(a - b)), np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-30),
            int(np.sum(a != b)), a.size))
    api.dk_tc_debug_set(9, 1)


if __name__ == "__main__" and "--layerwise" in sys.argv:
    layerwise(0)
    layerwise(1)
