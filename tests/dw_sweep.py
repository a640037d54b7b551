"""GPU micro-benchmark: depthwise 3x3 forward/backward at the ResNet-18-depsep shapes under each dispatch mode of
dk_dw_debug_set (2 = planes in shared memory, 3 = register windows / tiles), one process, graph-replayed launches.
Usage: python tests/dw_sweep.py [batch]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty
    runtime.ensure_init()
    st = runtime.stream
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3]
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    rng = np.random.default_rng(0)
    iters = 20
    for (H, C, s) in [(56, 64, 1), (28, 128, 1), (14, 256, 1), (7, 512, 1), (56, 64, 2), (28, 128, 2), (14, 256, 2)]:
        W = H
        OH = (H - 1) // s + 1
        n_in, n_out = N * C * H * W, N * C * OH * OH
        nbuf = max(2, int(np.ceil(160e6 / (4.0 * n_in))))
        xs = [asarray(rng.standard_normal((N, C, H, W)).astype(np.float32)) for _ in range(nbuf)]
        dys = [asarray(rng.standard_normal((N, C, OH, OH)).astype(np.float32)) for _ in range(nbuf)]
        w = asarray(rng.standard_normal((C, 3, 3)).astype(np.float32))
        y, dx, dw = empty((N, C, OH, OH)), empty((N, C, H, W)), empty((C, 3, 3))
        ws_ptr, ws_n = runtime.scratch(max(api.dk_dwconv_ws_bytes(N, C, H, W, 3, 3, s, 1), 1 << 20))
        K = {
            "fwd": (lambda i: api.dk_dwconv_fwd(xs[i].ptr, w.ptr, None, y.ptr, None, None, 0, N, C, H, W, 3, 3, s, 1, st()),
                    4 * (n_in + n_out)),
            "bwd": (lambda i: api.dk_dwconv_bwd(dys[i].ptr, xs[i].ptr, w.ptr, dx.ptr, dw.ptr, None, None, None, 0, None, 0.0,
                                                N, C, H, W, 3, 3, s, 1, ws_ptr, ws_n, st()), 4 * (2 * n_in + n_out)),
        }
        for name, (fn, nbytes) in K.items():
            line = "dw_%s N=%d C=%d HW=%d s=%d:" % (name, N, C, H, s)
            for mode in modes:
                api.dk_dw_debug_set(mode)
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
                line += "  mode %d: %6.1f us %6.0f GB/s (%.2f)" % (mode, 1e3 * med, nbytes / med / 1e6, nbytes / med / 1e6 / peak)
            print(line, flush=True)
        api.dk_dw_debug_set(1)
        del xs, dys


if __name__ == "__main__":
    main()
