"""Diagnostic (B200): BatchNorm of a tensor produced by a tcgen05 kernel, fused vs split kernels, on the producer's own
buffer and on a fresh copy of it, against a float64 NumPy evaluation of the same input."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bn64(x, gamma, beta, eps=1e-5):
    x = x.astype(np.float64)
    m = x.mean(axis=(0, 2, 3), keepdims=True)
    v = x.var(axis=(0, 2, 3), keepdims=True)
    return gamma * (x - m) / np.sqrt(v + eps) + beta


def main():
    import torch
    from dorknet_b200 import api
    from dorknet_b200.array import asarray, DeviceArray
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    rng = np.random.default_rng(3)
    for backend in (0, 1):
        api.dk_set_gemm_backend(backend)
        for (N, C, H, W, F, s) in ((8, 8, 16, 16, 8, 2), (8, 8, 8, 8, 8, 1), (16, 64, 28, 28, 64, 1)):
            X = rng.standard_normal((N, C, H, W)).astype(np.float32) * 3 + 1
            pw = PointwiseConvLayer("p", stride=s, filter_block_shape=(F, C), with_bias=False)
            pw.learned_params["weights"] = (rng.standard_normal((F, C)) * 0.01).astype(np.float32)
            y = pw.forward(X)
            torch.cuda.synchronize()
            yh = y.get().copy()
            gamma = np.ones((1, F, 1, 1), np.float32)
            beta = np.zeros((1, F, 1, 1), np.float32)
            ref = bn64(yh, gamma, beta)
            y_copy = asarray(yh)  # fresh buffer, written by a host->device copy
            out = []
            for name, src in (("producer buffer", y), ("fresh copy", y_copy)):
                for fused in (1, 0):
                    api.dk_tc_debug_set(9, fused)
                    bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=F)
                    r = bn.forward(src).get()
                    out.append("%s/%s %.2e" % (name, "fused" if fused else "split", np.max(np.abs(r - ref)) / np.max(np.abs(ref))))
            print("backend %d shape %s -> max rel err vs float64: %s" % (backend, (N, C, H, W, F, s), "; ".join(out)), flush=True)
    api.dk_set_gemm_backend(0)
    api.dk_tc_debug_set(9, 1)


if __name__ == "__main__":
    main()
