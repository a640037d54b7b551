"""One forward + backward of the depthwise-separable unit dw3x3 -> BatchNorm -> pointwise (56 x 56 x 64, batch 64) through
the product default path (statistics in the depthwise kernel, BatchNorm folded into the pointwise GEMMs), after two warm-up
passes: the command line `ncu --set full` captures for the kernels of bn_fold.cu / the TMA epilogue ring / dw_chan_bwd.
    ncu --set full --clock-control none --import-source on --launch-skip-before-match 0 -k regex:"dw3x3_rows_fwd|dw_stats_finalize|bn_fold|tc_gemm_kernel|dw_chan_bwd" \
        -o gpurun_out/r03e_fold_unit python tools/fold_once.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from dorknet_b200.array import ZeroSumGrad, asarray
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.depthwise_convolution import DepthwiseConvLayer
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    N, C, H, F = 64, 64, 56, 64
    rng = np.random.default_rng(0)
    dw = DepthwiseConvLayer("dw", filter_block_shape=(C, 3, 3), stride=1, padding=1, with_bias=False)
    bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
    pw = PointwiseConvLayer("pw", filter_block_shape=(F, C), with_bias=False)
    X = asarray(np.maximum(rng.standard_normal((N, C, H, H)), 0).astype(np.float32))
    g = rng.standard_normal((N, F, H, H)).astype(np.float32)
    g -= g.mean(axis=(0, 2, 3), keepdims=True)
    g = asarray(g)
    dY = ZeroSumGrad(g.t, g.shape)
    reps = int(os.environ.get("REPS", "3"))
    for _ in range(reps):
        pw.forward(bn.forward(dw.forward(X)))
        dw.backward(bn.backward(pw.backward(dY))).ptr
    torch.cuda.synchronize()
    assert pw._folded_bn is bn
    print("fold_once ok")


if __name__ == "__main__":
    main()
