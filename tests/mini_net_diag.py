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
    L = W.ours()
    for backend, mask in ((1, 0), (0, 0), (0, 6), (0, 5), (0, 3)):
        api.dk_set_gemm_backend(backend)
        api.dk_tc_debug_set(0, mask)
        net = defs.build_small_net(L, seed=123)
        for l in defs.iter_param_layers(net):
            for k in list(l.learned_params.keys()):
                l.learned_params[k] = d["init/%s/%s" % (l.layer_name, k)].copy()
        loss, scores = net.forward(d["X"], d["y"])
        net.backward()
        print("== backend %d tc-disable-mask %d  X%s loss %.7f (ref %.7f)" % (backend, mask, d["X"].shape, float(loss), float(d["losses"][0])))
        worst = []
        for l in defs.iter_param_layers(net):
            for k in l.grads.keys():
                a = l.grads[k].get().astype(np.float64)
                b = d["grad0/%s/%s" % (l.layer_name, k)].astype(np.float64)
                ma = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)
                l2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
                worst.append((ma, l2, l.layer_name + "/" + k, np.max(np.abs(b))))
        worst.sort(reverse=True)
        for ma, l2, nm, mx in worst[:6]:
            print("   %-28s max-abs/max %.3e   rel-L2 %.3e   max|ref| %.2e" % (nm, ma, l2, mx))
    api.dk_set_gemm_backend(0)
    api.dk_tc_debug_set(0, 0)


if __name__ == "__main__":
    main()
