"""GPU micro-benchmark: the BatchNorm statistics pass alone (dk_bn_fwd_train with y = NULL) at the ResNet-18-depsep
shapes under several values of dk_tc_debug_set key 24 (split-kernel CTAs per SM).  Usage: python tests/bn_stats_sweep.py 2,4,8"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty, zeros
    runtime.ensure_init()
    vals = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8]
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    rng = np.random.default_rng(0)
    N, iters = 64, 20
    for (H, C) in [(112, 64), (56, 64), (28, 128), (14, 256)]:
        HW = H * H
        nbuf = min(6, max(2, int(np.ceil(160e6 / (4.0 * N * C * HW)))))
        xs = [asarray(rng.standard_normal((N, C, H, H)).astype(np.float32)) for _ in range(nbuf)]
        g, b = asarray(np.ones(C, np.float32)), asarray(np.zeros(C, np.float32))
        rm, rs, sv = empty((C,)), empty((C,)), empty((4, C))
        ws = zeros((api.dk_bn_ws_bytes(C) // 4 + 16,))
        line = "bn_stats N=%d C=%d HW=%d:" % (N, C, HW)
        for v in vals:
            api.dk_tc_debug_set(24, v)

            def fn(i):
                api.dk_bn_fwd_train(xs[i % nbuf].ptr, None, g.ptr, b.ptr, rm.ptr, rs.ptr, 0, 0.95, 1e-5, sv.ptr, sv.ptr + 4 * C,
                                    sv.ptr + 8 * C, sv.ptr + 12 * C, 0, N, C, HW, ws.ptr, ws.size * 4, runtime.stream())
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for i in range(iters):
                    fn(i)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / iters)
            med = sorted(ts)[2]
            line += "  [%d] %6.1f us %5.0f GB/s (%.2f)" % (v, 1e3 * med, 4.0 * N * C * HW / med / 1e6, 4.0 * N * C * HW / med / 1e6 / peak)
        print(line, flush=True)
        api.dk_tc_debug_set(24, 8)


if __name__ == "__main__":
    main()
