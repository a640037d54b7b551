"""GPU micro-benchmark: BatchNorm forward (train) / backward at the ResNet-18-depsep shapes under several values of
dk_tc_debug_set key 9 (0 split kernels, 1 default = pipelined where it applies, 2 one-shot cluster kernels).
Usage: python tests/bn_sweep.py [batch] [modes]"""
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
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2]
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    rng = np.random.default_rng(0)
    iters = 20
    for (H, C) in [(112, 64), (56, 64), (28, 128), (14, 256), (7, 512)]:
        HW = H * H
        n_in = N * C * HW
        nbuf = min(6, max(2, int(np.ceil(160e6 / (4.0 * n_in)))))
        xs = [asarray(rng.standard_normal((N, C, H, H)).astype(np.float32)) for _ in range(nbuf)]
        dys = [asarray(rng.standard_normal((N, C, H, H)).astype(np.float32)) for _ in range(nbuf)]
        gamma, beta = asarray(np.ones(C, np.float32)), asarray(np.zeros(C, np.float32))
        y, dx = empty((N, C, H, H)), empty((N, C, H, H))
        dg, db, rm, rs, saved = empty((C,)), empty((C,)), empty((C,)), empty((C,)), empty((4, C))
        bn_ws = runtime.zeroed_workspace(api.dk_bn_ws_bytes(C))
        sv = saved.ptr
        for relu in (0, 1):
            K = {
                "fwd": (lambda i: api.dk_bn_fwd_train(xs[i].ptr, y.ptr, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, 0, 0.95, 1e-5, sv,
                                                      sv + 4 * C, sv + 8 * C, sv + 12 * C, relu, N, C, HW, bn_ws.data_ptr(),
                                                      bn_ws.numel(), st()), 4 * 3 * n_in, 4 * 2 * n_in),
                "bwd": (lambda i: api.dk_bn_bwd(dys[i].ptr, xs[i].ptr, gamma.ptr, sv, sv + 4 * C, sv + 8 * C, sv + 12 * C, dx.ptr,
                                                dg.ptr, db.ptr, relu, N, C, HW, bn_ws.data_ptr(), bn_ws.numel(), st()),
                        4 * 5 * n_in, 4 * 3 * n_in),
            }
            for name, (fn, nbytes, moved) in K.items():
                line = "bn_%s relu=%d N=%d C=%d HW=%d:" % (name, relu, N, C, H)
                for mode in modes:
                    api.dk_tc_debug_set(9, mode)
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
                    line += "  [%d] %6.1f us alg %5.0f GB/s (%.2f) fused-traffic %5.0f GB/s" % (
                        mode, 1e3 * med, nbytes / med / 1e6, nbytes / med / 1e6 / peak, moved / med / 1e6)
                print(line, flush=True)
            api.dk_tc_debug_set(9, 1)
        del xs, dys


if __name__ == "__main__":
    main()
