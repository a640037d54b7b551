"""Single-layer sweep (BASELINE.json config[1]) and per-kernel roofline micro-benchmark -- run on a B200.

    python tests/kernel_bench.py [--only NAME[,NAME]] [--batch 128] [--chans 64] [--hw 56] [--iters 20] [--out f.json]

Times each C-ABI entry point alone with CUDA events on the launching stream (>= 3 warm-up launches; the
operands rotate over enough distinct buffers to exceed the 126 MB L2, so every launch reads from HBM) and
reports achieved algorithmic GB/s (SURVEY.md 8(d) byte counts) against MEASURED_PEAKS.json.  With --once it
launches each selected kernel exactly once after warm-up (the command line ncu captures).
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
    ap.add_argument("--only", default="")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--chans", type=int, default=64)
    ap.add_argument("--filters", type=int, default=0)
    ap.add_argument("--hw", type=int, default=56)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--knobs", default="", help="comma list key=value for dk_tc_debug_set")
    ap.add_argument("--dw", type=int, default=-1, help="dk_dw_debug_set value")
    a = ap.parse_args()
    import torch
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty, zeros
    runtime.ensure_init()
    for kv in [x for x in a.knobs.split(",") if x]:
        k, v = kv.split("=")
        api.dk_tc_debug_set(int(k), int(v))
    if a.dw >= 0:
        api.dk_dw_debug_set(a.dw)
    st = runtime.stream
    N, C, H, W = a.batch, a.chans, a.hw, a.hw
    F = a.filters or C
    s = a.stride
    n_in = N * C * H * W
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rng = np.random.default_rng(0)
    nbuf = max(2, int(np.ceil(160e6 / (4.0 * n_in))))  # rotate over > L2 worth of distinct inputs
    xs = [asarray(rng.standard_normal((N, C, H, W)).astype(np.float32)) for _ in range(nbuf)]
    OH, OW = (H - 1) // s + 1, (W - 1) // s + 1
    dys_pw = [asarray(rng.standard_normal((N, F, OH, OW)).astype(np.float32)) for _ in range(nbuf)]
    dys_c = [asarray(rng.standard_normal((N, C, H, W)).astype(np.float32)) for _ in range(nbuf)]
    w_pw = asarray((rng.standard_normal((F, C)) / 8).astype(np.float32))
    w_dw = asarray(rng.standard_normal((C, 3, 3)).astype(np.float32))
    w_cv = asarray((rng.standard_normal((F, C, 3, 3)) / 24).astype(np.float32))
    gamma, beta = asarray(np.ones(C, np.float32)), asarray(np.zeros(C, np.float32))
    y_pw = empty((N, F, OH, OW))
    y_c = empty((N, C, H, W))
    y_cv = empty((N, F, H, W))
    dx = empty((N, C, H, W))
    dw_pw, dw_dw, dw_cv = empty((F, C)), empty((C, 3, 3)), empty((F, C, 3, 3))
    dg, db = empty((C,)), empty((C,))
    rm, rs = empty((C,)), empty((C,))
    saved = empty((4, C))
    bn_ws = runtime.zeroed_workspace(api.dk_bn_ws_bytes(C))
    ws_ptr, ws_n = runtime.scratch(max(api.dk_pwconv_ws_bytes(N, C, H, W, F, s), api.dk_conv2d_ws_bytes(N, C, H, W, F, 3, 3, 1, 1),
                                       api.dk_dwconv_ws_bytes(N, C, H, W, 3, 3, s, 1), 1 << 20))
    sv = saved.ptr
    # make the saved BN statistics valid once
    api.dk_bn_fwd_train(xs[0].ptr, y_c.ptr, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, 1, 0.95, 1e-5, sv, sv + 4 * C, sv + 8 * C,
                        sv + 12 * C, 0, N, C, H * W, bn_ws.data_ptr(), bn_ws.numel(), st())
    n_out_pw = N * F * OH * OW
    K = {
        "relu_fwd": (lambda i: api.dk_relu_fwd(xs[i].ptr, y_c.ptr, None, n_in, st()), 4 * 2 * n_in, 0),
        "relu_bwd": (lambda i: api.dk_relu_bwd(dys_c[i].ptr, xs[i].ptr, dx.ptr, n_in, st()), 4 * 3 * n_in, 0),
        "add_relu_fwd": (lambda i: api.dk_add_relu_fwd(xs[i].ptr, dys_c[i].ptr, y_c.ptr, n_in, st()), 4 * 3 * n_in, 0),
        "bn_fwd_train": (lambda i: api.dk_bn_fwd_train(xs[i].ptr, y_c.ptr, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, 0, 0.95, 1e-5,
                                                       sv, sv + 4 * C, sv + 8 * C, sv + 12 * C, 0, N, C, H * W,
                                                       bn_ws.data_ptr(), bn_ws.numel(), st()), 4 * 3 * n_in, 0),
        "bn_bwd": (lambda i: api.dk_bn_bwd(dys_c[i].ptr, xs[i].ptr, gamma.ptr, sv, sv + 4 * C, sv + 8 * C, sv + 12 * C, dx.ptr,
                                           dg.ptr, db.ptr, 0, N, C, H * W, bn_ws.data_ptr(), bn_ws.numel(), st()), 4 * 5 * n_in, 0),
        "dw_fwd": (lambda i: api.dk_dwconv_fwd(xs[i].ptr, w_dw.ptr, None, y_c.ptr, None, None, 0, N, C, H, W, 3, 3, s, 1, st()),
                   4 * (n_in + N * C * OH * OW), 2 * N * C * OH * OW * 9),
        "dw_bwd": (lambda i: api.dk_dwconv_bwd((dys_c[i] if s == 1 else dys_pw[i]).ptr, xs[i].ptr, w_dw.ptr, dx.ptr, dw_dw.ptr,
                                               None, None, None, 0, None, 0.0, N, C, H, W, 3, 3, s, 1, ws_ptr, ws_n, st()),
                   4 * (2 * n_in + N * C * OH * OW), 4 * N * C * OH * OW * 9),
        "pw_fwd": (lambda i: api.dk_pwconv_fwd(xs[i].ptr, w_pw.ptr, None, y_pw.ptr, N, C, H, W, F, s, ws_ptr, ws_n, st()),
                   4 * (N * C * OH * OW + n_out_pw), 2 * N * OH * OW * F * C),
        "pw_dgrad": (lambda i: api.dk_pwconv_dgrad(dys_pw[i].ptr, w_pw.ptr, dx.ptr, N, C, OH, OW, F, s, ws_ptr, ws_n, st()),
                     4 * (n_out_pw + N * C * OH * s * OW * s), 2 * N * OH * OW * F * C),
        "pw_wgrad": (lambda i: api.dk_pwconv_wgrad(dys_pw[i].ptr, xs[i].ptr, w_pw.ptr, dw_pw.ptr, None, 1e-4, N, C, H, W, F, s,
                                                   ws_ptr, ws_n, st()), 4 * (n_out_pw + N * C * OH * OW), 2 * N * OH * OW * F * C),
        "conv3x3_fwd": (lambda i: api.dk_conv2d_fwd(xs[i].ptr, w_cv.ptr, None, y_cv.ptr, N, C, H, W, F, 3, 3, 1, 1, ws_ptr, ws_n,
                                                    st()), 4 * (n_in + N * F * H * W), 2 * N * H * W * F * C * 9),
        "conv3x3_dgrad": (lambda i: api.dk_conv2d_dgrad(dys_pw[i].ptr if (s == 1 and F == C) else dys_c[i].ptr, w_cv.ptr, dx.ptr,
                                                        N, C, H, W, F, 3, 3, 1, 1, ws_ptr, ws_n, st()),
                          4 * (n_in + N * F * H * W), 2 * N * H * W * F * C * 9),
        "conv3x3_wgrad": (lambda i: api.dk_conv2d_wgrad(dys_pw[i].ptr if (s == 1 and F == C) else dys_c[i].ptr, xs[i].ptr,
                                                        w_cv.ptr, dw_cv.ptr, None, 1e-4, N, C, H, W, F, 3, 3, 1, 1, ws_ptr, ws_n,
                                                        st()), 4 * (n_in + N * F * H * W), 2 * N * H * W * F * C * 9),
    }
    only = [x for x in a.only.split(",") if x] or list(K)
    results = {}
    for name in only:
        fn, nbytes, flops = K[name]
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        if a.once:
            fn(0)
            torch.cuda.synchronize()
            continue
        # a CUDA graph of `iters` back-to-back launches (rotating operands) takes the Python/ctypes launch cost out
        # of the measurement; the graph is replayed 5 times and each replay timed with events on its stream
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
        r = {"shape": [N, C, H, W, F, s], "ms_median": med, "ms_min": ts[0], "alg_bytes": nbytes,
             "GBps": nbytes / (med * 1e-3) / 1e9, "frac_hbm_peak": nbytes / (med * 1e-3) / 1e9 / peak,
             "TFLOPs": flops / (med * 1e-3) / 1e12 if flops else None}
        results[name] = r
        print("%-14s %8.1f us  %7.1f GB/s (%.2f of %.0f)  %s" % (name, 1e3 * med, r["GBps"], r["frac_hbm_peak"], peak,
                                                               ("%.1f TFLOP/s" % r["TFLOPs"]) if flops else ""), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
