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
    for backend, mask in ((1, 0), (0, 0)):
        api.dk_set_gemm_backend(backend)
        api.dk_tc_debug_set(0, mask)
        net = defs.build_small_net(L, seed=123)
        for l in defs.iter_param_layers(net):
            for k in list(l.learned_params.keys()):
                l.learned_params[k] = d["init/%s/%s" % (l.layer_name, k)].copy()
        loss, scores = net.forward(d["X"], d["y"])
        net.backward()
        print("== backend %d tc-disable-mask %d  X%s loss %.7f (ref %.7f)" % (backend, mask, d["X"].shape, float(loss), float(d["losses"][0])))
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


if __name__ == "__main__":
    main()
