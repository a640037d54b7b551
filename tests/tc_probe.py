"""Bring-up probe for the tcgen05 pointwise GEMMs (run on a B200: python tests/tc_probe.py).

Runs fwd / dgrad / wgrad of PointwiseConvLayer through the C ABI on a few shapes and prints the normalised
max-abs error against a float64 NumPy evaluation, plus which backend served each call.  Optional
`--variants` sweeps the MN-major descriptor knobs (dk_tc_debug_set) to find the layout the hardware expects.
Each variant runs in this process; a trapped kernel aborts the process, so risky sweeps are run one per
process by the caller.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def run_case(N, C, F, H, W, s=1, seed=0):
    from dorknet_b200 import _lib
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = (rng.standard_normal((F, C)) / np.sqrt(C)).astype(np.float32)
    OH, OW = (H - 1) // s + 1, (W - 1) // s + 1
    dY = rng.standard_normal((N, F, OH, OW)).astype(np.float32)
    lay = PointwiseConvLayer("p", stride=s, filter_block_shape=(F, C), with_bias=False)
    lay.learned_params["weights"] = Wt
    t0, s0 = _lib.gemm_call_counts()
    Y = lay.forward(X).get()
    dX = lay.backward(dY).get()
    dW = lay.grads["weights"].get()
    t1, s1 = _lib.gemm_call_counts()
    X64, W64, dY64 = X[:, :, ::s, ::s].astype(np.float64), Wt.astype(np.float64), dY.astype(np.float64)
    Yr = np.einsum("fc,nchw->nfhw", W64, X64)
    dXs = np.einsum("fc,nfhw->nchw", W64, dY64)
    dXr = np.zeros((N, C, OH * s, OW * s))
    dXr[:, :, ::s, ::s] = dXs
    dWr = np.einsum("nfhw,nchw->fc", dY64, X64)
    return dict(shape=(N, C, F, H, W, s), fwd=nerr(Y, Yr), dgrad=nerr(dX, dXr), wgrad=nerr(dW, dWr), tc=t1 - t0,
                simt=s1 - s0)


def run_conv(N, C, H, W, F, k, s, p, seed=0):
    from dorknet_b200 import _lib
    from dorknet_b200.layers.convolution import ConvLayer
    from oracle import oracle as O  # checker
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = (rng.standard_normal((F, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    lay = ConvLayer("c", (F, C, k, k), stride=s, padding=p, with_bias=False)
    lay.needs_input_grad = True
    lay.learned_params["weights"] = Wt
    t0, s0 = _lib.gemm_call_counts()
    Y = lay.forward(X).get()
    Yo, cache = O.conv_fwd(X, Wt, None, s, p)
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dX = lay.backward(dY).get()
    dW = lay.grads["weights"].get()
    t1, s1 = _lib.gemm_call_counts()
    dXo, g = O.conv_bwd(dY, Wt, cache, s, p)
    return dict(shape=(N, C, H, W, F, k, s, p), fwd=nerr(Y, Yo), dgrad=nerr(dX, dXo), wgrad=nerr(dW, g["weights"]),
                tc=t1 - t0, simt=s1 - s0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mask", type=int, default=0, help="dk_tc_debug_set(0, mask): bit0 fwd, bit1 dgrad, bit2 wgrad off")
    ap.add_argument("--knobs", default="", help="comma list key=value for dk_tc_debug_set keys 1..5")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--only", default="", choices=["", "pw", "conv"])
    a = ap.parse_args()
    from dorknet_b200 import api, runtime
    runtime.ensure_init()
    api.dk_tc_debug_set(0, a.mask)
    for kv in [x for x in a.knobs.split(",") if x]:
        k, v = kv.split("=")
        api.dk_tc_debug_set(int(k), int(v))
    cases = [(2, 64, 64, 8, 16)] if a.small else [
        (2, 64, 64, 8, 16), (2, 32, 64, 8, 8), (3, 64, 128, 28, 28), (2, 128, 64, 12, 12), (2, 256, 256, 14, 14),
        (4, 64, 64, 56, 56), (2, 40, 48, 10, 10), (1, 8, 16, 4, 8), (2, 512, 512, 8, 8),
        (2, 64, 64, 14, 14, 2), (2, 64, 128, 15, 15, 2), (2, 256, 512, 7, 7), (3, 512, 512, 7, 7), (3, 256, 512, 14, 14, 2),
        (2, 64, 64, 112, 112, 2), (2, 8, 8, 5, 5), (2, 24, 40, 9, 7, 3)]
    if a.only != "conv":
        for c in cases:
            r = run_case(*c)
            print("pw  N,C,F,H,W,s=%-26s fwd %.2e  dgrad %.2e  wgrad %.2e   calls tc=%d simt=%d" % (
                r["shape"], r["fwd"], r["dgrad"], r["wgrad"], r["tc"], r["simt"]), flush=True)
    convs = [(2, 3, 33, 33, 8, 5, 2, 1), (2, 3, 225, 225, 64, 5, 2, 1), (2, 32, 14, 14, 64, 4, 2, 1),
             (2, 64, 16, 16, 64, 3, 1, 1), (2, 1, 28, 28, 32, 3, 1, 1), (2, 5, 10, 10, 7, 3, 2, 1), (1, 16, 9, 9, 300, 3, 1, 0)]
    if a.only != "pw" and not a.small:
        for c in convs:
            r = run_conv(*c)
            print("conv N,C,H,W,F,k,s,p=%-30s fwd %.2e  dgrad %.2e  wgrad %.2e   calls tc=%d simt=%d" % (
                r["shape"], r["fwd"], r["dgrad"], r["wgrad"], r["tc"], r["simt"]), flush=True)


if __name__ == "__main__":
    main()
