"""GPU micro-benchmark: pointwise conv fwd / dgrad / wgrad at the ResNet-18-depsep shapes, one process, graph-replayed
launches, optionally under several values of a dk_tc_debug_set knob.
Usage: python tests/pw_sweep.py [batch] [key=v1,v2,...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # C, F, H (input), stride
    (64, 64, 56, 1), (128, 128, 28, 1), (256, 256, 14, 1), (512, 512, 7, 1), (256, 512, 7, 1),
    (64, 64, 112, 2), (64, 128, 56, 2), (128, 256, 28, 2), (256, 512, 14, 2), (64, 128, 28, 1), (128, 256, 14, 1),
]


def main():
    import torch
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty
    runtime.ensure_init()
    st = runtime.stream
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    key, vals = -1, [None]
    if len(sys.argv) > 2:
        k, v = sys.argv[2].split("=")
        key, vals = int(k), [int(t) for t in v.split(",")]
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    rng = np.random.default_rng(0)
    iters = 20
    for (C, F, H, s) in SHAPES:
        W = H
        OH = (H - 1) // s + 1
        P = OH * OH
        n_in, n_out = N * C * H * W, N * F * P
        nbuf = max(2, int(np.ceil(160e6 / (4.0 * min(n_in, 4 * n_out)))))
        nbuf = min(nbuf, 6)
        xs = [asarray(rng.standard_normal((N, C, H, W)).astype(np.float32)) for _ in range(nbuf)]
        dys = [asarray(rng.standard_normal((N, F, OH, OH)).astype(np.float32)) for _ in range(nbuf)]
        w = asarray((rng.standard_normal((F, C)) / 8).astype(np.float32))
        y, dx, dw = empty((N, F, OH, OH)), empty((N, C, OH * s, OH * s)), empty((F, C))
        ws_ptr, ws_n = runtime.scratch(max(api.dk_pwconv_ws_bytes(N, C, max(H, OH * s), max(W, OH * s), F, s), 1 << 20))
        used_in = N * C * P  # pixels the GEMM reads
        K = {
            "fwd": (lambda i: api.dk_pwconv_fwd(xs[i].ptr, w.ptr, None, y.ptr, N, C, H, W, F, s, ws_ptr, ws_n, st()),
                    4 * (used_in + n_out)),
            "dgrad": (lambda i: api.dk_pwconv_dgrad(dys[i].ptr, w.ptr, dx.ptr, N, C, OH, OH, F, s, ws_ptr, ws_n, st()),
                      4 * (n_out + N * C * OH * s * OH * s)),
            "wgrad": (lambda i: api.dk_pwconv_wgrad(dys[i].ptr, xs[i].ptr, w.ptr, dw.ptr, None, 1e-4, N, C, H, W, F, s, ws_ptr,
                                                    ws_n, st()), 4 * (n_out + used_in)),
        }
        for name, (fn, nbytes) in K.items():
            if only and name not in only:
                continue
            line = "pw_%-5s N=%d C=%d F=%d HW=%d s=%d:" % (name, N, C, F, H, s)
            for v in vals:
                if v is not None:
                    api.dk_tc_debug_set(key, v)
                for i in range(3):
                    fn(i % nbuf)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(iters):
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
                    ts.append(e0.elapsed_time(e1) / iters)
                med = sorted(ts)[2]
                line += "  [%s] %6.1f us %6.0f GB/s (%.2f)" % ("-" if v is None else str(v), 1e3 * med, nbytes / med / 1e6,
                                                              nbytes / med / 1e6 / peak)
            print(line, flush=True)
        del xs, dys


if __name__ == "__main__":
    main()
