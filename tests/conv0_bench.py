"""First-layer convolution micro-benchmark (ResNet-18-depsep conv0: 3 -> 64, 5x5, stride 2, pad 1 at 225x225) -- run on a B200.

    python tests/conv0_bench.py [--batch 64] [--hw 225] [--filters 64] [--k 5] [--stride 2] [--pad 1] [--knobs 8=0]

Times dk_conv2d_fwd and dk_conv2d_wgrad alone (CUDA graph of `iters` launches over rotating operands, CUDA events) and
checks both against the oracle on the first image pair.  --knobs 8=0 selects the generic gather-loader kernels
instead of conv_rows.cu for an A/B comparison.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chans", type=int, default=3)
    ap.add_argument("--hw", type=int, default=225)
    ap.add_argument("--filters", type=int, default=64)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--stride", type=int, default=2)
    ap.add_argument("--pad", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--knobs", default="")
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty
    from oracle import oracle as O  # checker only
    runtime.ensure_init()
    for kv in [x for x in a.knobs.split(",") if x]:
        k, v = kv.split("=")
        api.dk_tc_debug_set(int(k), int(v))
    st = runtime.stream
    N, C, H, W, F, k, s, p = a.batch, a.chans, a.hw, a.hw, a.filters, a.k, a.stride, a.pad
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    rng = np.random.default_rng(0)
    nbuf = 3
    xh = [rng.standard_normal((N, C, H, W)).astype(np.float32) for _ in range(nbuf)]
    dyh = [rng.standard_normal((N, F, OH, OW)).astype(np.float32) for _ in range(nbuf)]
    xs, dys = [asarray(v) for v in xh], [asarray(v) for v in dyh]
    wh = (rng.standard_normal((F, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    w = asarray(wh)
    y, dw = empty((N, F, OH, OW)), empty((F, C, k, k))
    ws_ptr, ws_n = runtime.scratch(api.dk_conv2d_ws_bytes(N, C, H, W, F, k, k, s, p))
    K = {
        "conv0_fwd": (lambda i: api.dk_conv2d_fwd(xs[i].ptr, w.ptr, None, y.ptr, N, C, H, W, F, k, k, s, p, ws_ptr, ws_n, st()),
                      4 * (N * C * H * W + N * F * OH * OW), 2 * N * OH * OW * F * C * k * k),
        "conv0_wgrad": (lambda i: api.dk_conv2d_wgrad(dys[i].ptr, xs[i].ptr, w.ptr, dw.ptr, None, 0.0, N, C, H, W, F, k, k, s, p,
                                                      ws_ptr, ws_n, st()),
                        4 * (N * C * H * W + N * F * OH * OW), 2 * N * OH * OW * F * C * k * k),
    }
    # parity on two images (oracle im2col of the full batch would take a while)
    K["conv0_fwd"][0](0)
    K["conv0_wgrad"][0](0)
    torch.cuda.synchronize()
    nb = min(N, 2)
    Yo, cache = O.conv_fwd(xh[0][:nb], wh, None, s, p)
    e_fwd = float(np.max(np.abs(y.get()[:nb] - Yo)) / np.max(np.abs(Yo)))
    _, g = O.conv_bwd(dyh[0], wh, O.conv_fwd(xh[0], wh, None, s, p)[1], s, p) if N <= 8 else (None, None)
    e_w = float(np.max(np.abs(dw.get() - g["weights"])) / np.max(np.abs(g["weights"]))) if g else None
    print("parity: fwd err %.3e, wgrad err %s" % (e_fwd, "%.3e" % e_w if e_w is not None else "n/a (batch > 8)"), flush=True)
    results = {"parity": {"fwd": e_fwd, "wgrad": e_w}}
    for name, (fn, nbytes, flops) in K.items():
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        if a.once:
            fn(0)
            torch.cuda.synchronize()
            continue
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(a.iters):
                fn(i % nbuf)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / a.iters)
        ts.sort()
        med = ts[len(ts) // 2]
        results[name] = {"ms_median": med, "GBps": nbytes / (med * 1e-3) / 1e9, "TFLOPs": flops / (med * 1e-3) / 1e12}
        print("%-12s %8.1f us  %7.1f GB/s  %.1f TFLOP/s" % (name, 1e3 * med, results[name]["GBps"], results[name]["TFLOPs"]),
              flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
