"""Diagnostic (B200): one training step of ResNet-18-depsep (batch 16, 225x225) with the fused BatchNorm cluster kernels
and the row-staged conv0 kernels on vs off, same TF32 GEMM backend: loss and per-layer gradient differences
(relative L2).  At full size BatchNorm averages over >= 784 samples per channel, so rounding-level changes must stay
rounding-level here (the 8x2x2-sample miniature of test_gpu_golden.py amplifies them through ReLU mask flips)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dorknet_b200 import api, workloads as W
    M = W.ours()
    X, _, Y = W.synthetic_batch(16, 3, 225, 120, seed=5, mixup=False)
    backend = int(sys.argv[1]) if len(sys.argv) > 1 else 0  # 1: fp32 SIMT GEMMs (no TF32 truncation anywhere)
    api.dk_set_gemm_backend(backend)
    print("GEMM backend %d" % backend)
    res = {}
    for key, (rows, bnf) in {"split": (0, 0), "split again": (0, 0), "rows only": (1, 0), "bn only": (0, 1), "fused": (1, 1)}.items():
        api.dk_tc_debug_set(8, rows)
        api.dk_tc_debug_set(9, bnf)
        net = W.build_resnet18_depsep(M, classes=120, seed=0)
        loss, _ = net.forward(X, Y)
        net.backward()
        grads = {}
        for l in W.iter_param_layers(net) if hasattr(W, "iter_param_layers") else []:
            for k, v in l.grads.items():
                grads[l.layer_name + "/" + k] = v.get().astype(np.float64)
        if not grads:
            for l in net.layers:
                subs = [l] + list(getattr(l, "layer_list", []) or []) + ([l.skip_projection] if getattr(l, "skip_projection", None) is not None else [])
                for m in subs:
                    if getattr(m, "grads", None):
                        for k, v in m.grads.items():
                            grads[m.layer_name + "/" + k] = v.get().astype(np.float64)
        res[key] = (float(loss), grads)
    l0, g0 = res["split"]
    for key in ("split again", "rows only", "bn only", "fused"):
        l1, g1 = res[key]
        worst = sorted(((np.linalg.norm(g1[k] - g0[k]) / max(np.linalg.norm(g0[k]), 1e-30), k) for k in g0
                        if not k.endswith("_dw_bn/beta")), reverse=True)
        print("%-12s loss %.7f (split %.7f); worst rel-L2 grad diffs: %s; median %.2e over %d tensors" % (
            key, l1, l0, ", ".join("%s %.2e" % (k, e) for e, k in worst[:4]), float(np.median([e for e, _ in worst])), len(worst)))
    api.dk_tc_debug_set(8, 1)
    api.dk_tc_debug_set(9, 1)
    api.dk_set_gemm_backend(0)


if __name__ == "__main__":
    main()
