"""bring-up probe for conv_tma.cu: forward / dgrad / wgrad separately against the oracle, blocking launches"""
import os, sys
import numpy as np
os.environ.setdefault("CUDA_LAUNCH_BLOCKING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
import torch
from dorknet_b200 import api, runtime, _lib
from dorknet_b200.array import asarray, empty

def err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))

def run(case, which):
    N, C, H, W, F, k, s, p = case
    rng = np.random.default_rng(1)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = (rng.standard_normal((F, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    Yo, cache = O.conv_fwd(X, Wt, None, s, p)
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dXo, g = O.conv_bwd(dY, Wt, cache, s, p)
    x, w, dy = asarray(X), asarray(Wt), asarray(dY)
    y, dx, dw = empty(Yo.shape), empty(X.shape), empty(Wt.shape)
    ws, wsn = runtime.scratch(api.dk_conv2d_ws_bytes(N, C, H, W, F, k, k, s, p))
    st = runtime.stream()
    tc0 = _lib.gemm_call_counts()
    if "f" in which:
        api.dk_conv2d_fwd(x.ptr, w.ptr, None, y.ptr, N, C, H, W, F, k, k, s, p, ws, wsn, st)
        torch.cuda.synchronize()
        print(case, "fwd err %.3e" % err(y.get(), Yo), flush=True)
    if "d" in which:
        api.dk_conv2d_dgrad(dy.ptr, w.ptr, dx.ptr, N, C, H, W, F, k, k, s, p, ws, wsn, st)
        torch.cuda.synchronize()
        print(case, "dgrad err %.3e" % err(dx.get(), dXo), flush=True)
    if "w" in which:
        api.dk_conv2d_wgrad(dy.ptr, x.ptr, w.ptr, dw.ptr, None, 0.0, N, C, H, W, F, k, k, s, p, ws, wsn, st)
        torch.cuda.synchronize()
        print(case, "wgrad err %.3e" % err(dw.get(), g["weights"]), flush=True)
    print("  calls (tc, simt):", tuple(b - a for a, b in zip(tc0, _lib.gemm_call_counts())), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "fdw"
    cases = [(2, 64, 56, 56, 64, 3, 1, 1), (3, 32, 28, 28, 32, 3, 1, 1), (2, 40, 12, 36, 48, 3, 1, 1), (2, 16, 10, 16, 24, 3, 1, 0),
             (2, 8, 9, 20, 8, 5, 1, 2), (1, 24, 8, 12, 16, 3, 1, 2), (2, 128, 12, 12, 128, 3, 1, 1), (3, 64, 20, 8, 200, 3, 1, 1)]
    if len(sys.argv) > 2:
        cases = [cases[int(sys.argv[2])]]
    for c in cases:
        run(c, which)
