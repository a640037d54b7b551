"""-m gpu: the CUDA kernels against the oracle (oracle/oracle.py) on seeded inputs, at shapes chosen to hit every
kernel variant: register-window vs tile depthwise (vector widths 4/2/1, row bands, stride 2), TMA vs software-gather
tcgen05 GEMM operands (stride 2, 7x7 planes, odd channel counts), implicit-GEMM convolutions, split-K plans.
All calls go through the C ABI via the layer classes."""
import numpy as np
import pytest

from gpu_util import FP32, FP32_RED, GEMM, GEMM_W, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


DW_CASES = [
    # N, C, H, W, k, s, p, bias
    (3, 8, 56, 56, 3, 1, 1, False),   # vec 4, bands
    (2, 16, 28, 28, 3, 1, 1, True),   # vec 4
    (2, 24, 14, 14, 3, 1, 1, False),  # vec 2
    (3, 40, 7, 7, 3, 1, 1, True),     # vec 1
    (1, 5, 9, 112, 3, 1, 1, False),   # 28 strips/row
    (1, 3, 12, 224, 3, 1, 1, False),  # a row spans two warps (56 strips)
    (2, 8, 11, 13, 3, 1, 1, False),   # odd sizes -> scalar strips
    (2, 8, 56, 56, 3, 2, 1, False),   # stride 2 (x.5 patch count): tile / planes kernel
    (3, 16, 28, 28, 3, 2, 1, True),   # stride 2, vec 4 output rows
    (3, 24, 14, 14, 3, 2, 1, False),  # stride 2, scalar items, several planes per CTA
    (5, 40, 7, 7, 3, 2, 1, True),     # stride 2, odd plane (7 -> 4), ragged last group
    (2, 5, 13, 9, 3, 2, 1, False),    # stride 2, odd non-square
    (70, 4, 14, 14, 3, 1, 1, True),   # per-channel backward: several stages of 16 images, ragged last stage, cluster of 5
    (19, 3, 28, 28, 3, 1, 1, False),  # per-channel backward: cluster of 5 ranks, ragged image ranges
    (9, 2, 56, 56, 3, 1, 1, True),    # per-channel backward: more stage items than ring slots
    (2, 6, 15, 15, 5, 1, 2, True),    # 5x5: generic tile kernel
    # channel-group kernels (depthwise_group.cu; rows=1): 4 channels of 7x7 / one of 14x14 per CTA, all images resident
    (64, 512, 7, 7, 3, 1, 1, False), (5, 296, 7, 7, 3, 1, 1, True), (64, 256, 14, 14, 3, 1, 1, True), (3, 80, 14, 14, 3, 1, 1, False),
]


@pytest.mark.parametrize("rows", [1, 0, 2, 3, 4, 5])  # default dispatch (channel-group / per-channel kernels), tiles, planes-in-smem, register windows, default without per-channel, default without channel-group
@pytest.mark.parametrize("case", DW_CASES)
def test_depthwise_vs_oracle(O, case, rows):
    from dorknet_b200 import api
    from dorknet_b200.layers.depthwise_convolution import DepthwiseConvLayer
    N, C, H, W, k, s, p, bias = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = rng.standard_normal((C, k, k)).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32) if bias else None
    api.dk_dw_debug_set(rows)
    try:
        lay = DepthwiseConvLayer("dw", (C, k, k), stride=s, padding=p, with_bias=bias)
        lay.learned_params["weights"] = Wt
        if bias:
            lay.learned_params["bias"] = b
        Y = lay.forward(X)
        Yo, cache = O.depthwise_fwd(X, Wt, b, s, p)
        assert Y.shape == Yo.shape
        assert_close(Y.get(), Yo, FP32, "Y")
        dY = rng.standard_normal(Yo.shape).astype(np.float32)
        dXo, g = O.depthwise_bwd(dY, Wt, cache, s, p, 0.0, bias)
        dX = lay.backward(dY)
        assert_close(dX.get(), dXo, FP32, "dX")
        assert_close(lay.grads["weights"].get(), g["weights"], 5 * FP32_RED, "dW")
        if bias:
            assert_close(lay.grads["bias"].get(), g["bias"], 5 * FP32_RED, "db")
        # residual join folded into the backward
        add = rng.standard_normal(X.shape).astype(np.float32)
        from dorknet_b200.array import asarray
        dX2 = lay.backward(dY, dx_add=asarray(add))
        assert_close(dX2.get(), dXo + add, FP32, "dX + skip")
    finally:
        api.dk_dw_debug_set(1)


def test_depthwise_backward_is_deterministic():
    from dorknet_b200.layers.depthwise_convolution import DepthwiseConvLayer
    rng = np.random.default_rng(5)
    X = rng.standard_normal((4, 32, 56, 56)).astype(np.float32)
    dY = rng.standard_normal((4, 32, 56, 56)).astype(np.float32)
    lay = DepthwiseConvLayer("dw", (32, 3, 3), stride=1, padding=1, with_bias=True)
    lay.forward(X)
    lay.backward(dY)
    a = lay.grads["weights"].get().copy(), lay.grads["bias"].get().copy()
    for _ in range(3):
        lay.backward(dY)
        assert np.array_equal(lay.grads["weights"].get(), a[0]) and np.array_equal(lay.grads["bias"].get(), a[1])


PW_CASES = [
    # N, C, F, H, W, s
    (2, 64, 64, 8, 16, 1), (3, 64, 128, 28, 28, 1), (2, 256, 256, 14, 14, 1), (2, 40, 48, 10, 10, 1),
    (2, 64, 64, 14, 14, 2), (2, 64, 128, 15, 15, 2), (2, 256, 512, 7, 7, 1), (3, 512, 512, 7, 7, 1),
    (2, 64, 64, 112, 112, 2), (2, 24, 40, 9, 7, 3), (2, 6, 10, 5, 5, 1),
    # cfg5 (MobileNet stack) channel counts: 512 -> 1024 and 1024 -> 1024 at 7x7, and 1024 at 14x14 planes
    (2, 512, 1024, 7, 7, 1), (3, 1024, 1024, 7, 7, 1), (2, 1024, 512, 14, 14, 1),
]


@pytest.mark.parametrize("case", PW_CASES)
def test_pointwise_tcgen05_vs_oracle(O, case):
    from dorknet_b200 import _lib
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    from dorknet_b200.regularisers.l2 import l2
    N, C, F, H, W, s = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = (rng.standard_normal((F, C)) / np.sqrt(C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    lay = PointwiseConvLayer("p", stride=s, filter_block_shape=(F, C), with_bias=True, weight_regulariser=l2(1e-2))
    lay.learned_params["weights"], lay.learned_params["bias"] = Wt, b
    tc0, _ = _lib.gemm_call_counts()
    Y = lay.forward(X)
    Yo, cache = O.pointwise_fwd(X, Wt, b, s)
    assert_close(Y.get(), Yo, GEMM, "Y")
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dXo, g = O.pointwise_bwd(dY, Wt, cache, s, 1e-2, True)
    dX = lay.backward(dY)
    assert dX.shape == dXo.shape
    assert_close(dX.get(), dXo, GEMM, "dX")
    assert_close(lay.grads["weights"].get(), g["weights"], GEMM_W, "dW")
    assert_close(lay.grads["bias"].get(), g["bias"], FP32_RED, "db")
    tc1, _ = _lib.gemm_call_counts()
    if C % 4 == 0:
        assert tc1 - tc0 == 3, "expected fwd, dgrad and wgrad on the tcgen05 path"


CONV_CASES = [
    # N, C, H, W, F, k, s, p
    (2, 3, 33, 33, 8, 5, 2, 1), (2, 3, 65, 65, 64, 5, 2, 1), (2, 32, 14, 14, 64, 4, 2, 1), (2, 64, 16, 16, 64, 3, 1, 1),
    (2, 1, 28, 28, 32, 3, 1, 1), (2, 5, 10, 10, 7, 3, 2, 1), (1, 16, 9, 9, 300, 3, 1, 0),
    # small-K row-staged kernels (conv_rows.cu): conv0 at full width (OW = 112: partial last pixel chunk), several
    # tiles per output row, MobileNet conv0, many rows per CTA
    (2, 3, 225, 225, 64, 5, 2, 1), (1, 3, 40, 300, 16, 3, 1, 1), (3, 3, 224, 224, 32, 3, 2, 1), (40, 3, 33, 33, 24, 5, 2, 1),
    # span staging (bulk-copied row spans + 16-byte expansion stores; needs N*C*H*W % 4 == 0): conv0 itself, padding 2 and 0,
    # stride 1 with 5 columns, clipped top / bottom rows, OW not a multiple of 4 (forward only takes it)
    (4, 3, 225, 225, 64, 5, 2, 1), (4, 3, 224, 224, 16, 5, 2, 2), (4, 2, 30, 30, 8, 5, 1, 2), (4, 3, 19, 19, 8, 3, 1, 0),
    (8, 3, 31, 31, 40, 3, 2, 1),
    # cfg2's layer at its real plane size (3x3 64 -> 64 at 56x56), the MNIST stack's layers at theirs (cfg1)
    (2, 64, 56, 56, 64, 3, 1, 1), (3, 32, 28, 28, 32, 3, 1, 1), (3, 32, 28, 28, 64, 4, 2, 1), (3, 64, 14, 14, 64, 3, 1, 1),
    (3, 64, 14, 14, 128, 4, 2, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tcgen05_vs_oracle(O, case):
    _conv_case(O, case)


CONV_TMA_CASES = [
    # N, C, H, W, F, k, s, p -- stride 1, W % 4 == 0: the 4-D tensor-map kernels of conv_tma.cu (forward, dgrad, wgrad)
    (2, 64, 56, 56, 64, 3, 1, 1),    # cfg2: two 32-column segments per row (24 valid in the second), resident filters
    (3, 32, 28, 28, 32, 3, 1, 1),    # MNIST conv2: one partial column segment
    (2, 40, 12, 36, 48, 3, 1, 1),    # channel counts that are not multiples of 32, H not a multiple of 4, rectangular
    (2, 16, 10, 16, 24, 3, 1, 0),    # no padding: output 8 x 14 (OW % 4 != 0: wgrad falls back, forward / dgrad stay)
    (2, 8, 9, 20, 8, 5, 1, 2),       # 5 x 5, padding 2, 8 channels (one partial k step)
    (1, 24, 8, 12, 16, 3, 1, 2),     # padding larger than (k-1)/2: output grows to 10 x 14
    (2, 128, 12, 12, 128, 3, 1, 1),  # filters too large to stay resident: streamed next to the input boxes
    (3, 64, 20, 8, 200, 3, 1, 1),    # F > 128: wgrad falls back, N = 224 forward
]


@pytest.mark.parametrize("case", CONV_TMA_CASES)
def test_conv_tma_stride1_vs_oracle(O, case):
    from dorknet_b200 import _lib
    tc0, s0 = _lib.gemm_call_counts()
    _conv_case(O, case)
    tc1, s1 = _lib.gemm_call_counts()
    assert tc1 - tc0 == 3 and s1 == s0, "expected forward, dgrad and wgrad on the tensor-core path"


@pytest.mark.parametrize("case", [(2, 64, 56, 56, 64, 3, 1, 1), (2, 40, 13, 36, 48, 3, 1, 1), (3, 32, 28, 28, 32, 3, 1, 1)])
def test_conv_tma_wgrad_one_row_per_step(O, case):
    """conv_s1_wgrad2_kernel<1, 3> (one X row per pipeline step; dk_tc_debug_set key 26 = 1).  The default for 3 x 3 is two
    rows per step (an odd row count per unit then ends in a one-row step: the 13-row case of CONV_TMA_CASES' sibling here)."""
    from dorknet_b200 import _lib
    _lib.api.dk_tc_debug_set(26, 1)
    try:
        _conv_case(O, case)
    finally:
        _lib.api.dk_tc_debug_set(26, 0)


def test_conv_tma_wgrad_channels_not_multiple_of_four(O):
    """C = 10: the wgrad kernel's scalar partial-sum stores and cw2_reduce_kernel<1> (the 16-byte paths need C % 4 == 0)"""
    from dorknet_b200 import _lib
    tc0, _ = _lib.gemm_call_counts()
    _conv_case(O, (2, 10, 12, 16, 12, 3, 1, 1))
    tc1, _ = _lib.gemm_call_counts()
    assert tc1 - tc0 >= 1, "expected the weight gradient on the tensor-core path"


CONV_MAT_CASES = [
    # N, C, H, W, F, k, s, p -- shapes outside conv_tma.cu / conv_rows.cu: materialised transposed patches + the all-TMA
    # pointwise GEMMs (gemm_tcgen05.cu cvm_*), dX by gather-form col2im
    (3, 32, 28, 28, 64, 4, 2, 1),    # MNIST conv3: 4x4 stride 2, 14 x 14 output (P = 196)
    (3, 64, 14, 14, 128, 4, 2, 1),   # MNIST conv5: 7 x 7 output (P = 49: padded patch pitch, dY repacked for wgrad)
    (3, 64, 14, 14, 64, 3, 1, 1),    # MNIST conv4: stride 1 but 56-byte rows (conv_tma needs 16-byte pitches)
    (2, 32, 14, 14, 64, 4, 2, 1),
    (2, 16, 15, 17, 24, 3, 2, 1),    # x.5 patch counts (floor), odd rectangular planes, P = 8 * 9
    (2, 20, 11, 11, 12, 5, 3, 2),    # stride 3, 5 x 5: taps that never reach some dX elements
    (1, 16, 9, 9, 300, 3, 1, 0),     # F > 256: two n-blocks
    (2, 64, 30, 30, 64, 3, 2, 0),    # no padding, stride 2
]


@pytest.mark.parametrize("case", CONV_MAT_CASES)
def test_conv_materialised_patches_vs_oracle(O, case):
    """every pass stays on the tensor-core path, and agrees with the gather variants it replaces"""
    from dorknet_b200 import _lib, api
    tc0, s0 = _lib.gemm_call_counts()
    _conv_case(O, case)
    tc1, s1 = _lib.gemm_call_counts()
    assert tc1 - tc0 == 3 and s1 == s0, "expected forward, dgrad and wgrad on the tensor-core path"
    api.dk_tc_debug_set(20, 0)
    try:
        _conv_case(O, case)
    finally:
        api.dk_tc_debug_set(20, 1)


@pytest.mark.parametrize("case", [(4, 3, 225, 225, 64, 5, 2, 1), (2, 1, 28, 28, 32, 3, 1, 1)])
def test_conv_rows_generic_staging_vs_oracle(O, case):
    """the same small-K kernels with span staging switched off (4-byte cp.async staging of any geometry)"""
    from dorknet_b200 import api
    api.dk_tc_debug_set(8, 2)
    try:
        _conv_case(O, case)
    finally:
        api.dk_tc_debug_set(8, 1)


def _conv_case(O, case):
    from dorknet_b200.layers.convolution import ConvLayer
    N, C, H, W, F, k, s, p = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = (rng.standard_normal((F, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    bias = (N + F) % 2 == 1
    b = rng.standard_normal(F).astype(np.float32) if bias else None
    lay = ConvLayer("c", (F, C, k, k), stride=s, padding=p, with_bias=bias)
    lay.learned_params["weights"] = Wt
    if bias:
        lay.learned_params["bias"] = b
    Y = lay.forward(X)
    Yo, cache = O.conv_fwd(X, Wt, b, s, p)
    assert np.array_equal(lay.im2col_materialise(X).get(), cache["P"])  # bit-exact im2col index map
    assert_close(Y.get(), Yo, GEMM, "Y")
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dXo, g = O.conv_bwd(dY, Wt, cache, s, p)
    dX = lay.backward(dY)  # lazy: computed on first read
    assert not dX.is_materialised
    assert_close(dX.get(), dXo, GEMM, "dX")
    assert dX.is_materialised
    assert_close(lay.grads["weights"].get(), g["weights"], GEMM_W, "dW")


def test_full_size_properties_resnet_shapes():
    """BASELINE full sizes (batch 64): size-independent properties instead of an oracle run --
    linearity of the pointwise GEMM, dX of depthwise against a finite-difference probe along one direction,
    BN output statistics (zero mean / unit variance per channel)."""
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.depthwise_convolution import DepthwiseConvLayer
    rng = np.random.default_rng(11)
    N, C, H, W = 64, 64, 56, 56
    X1 = rng.standard_normal((N, C, H, W)).astype(np.float32)
    X2 = rng.standard_normal((N, C, H, W)).astype(np.float32)
    pw = PointwiseConvLayer("p", filter_block_shape=(64, 64), with_bias=False)
    y1 = pw.forward(X1).get().copy()
    y2 = pw.forward(X2).get().copy()
    y12 = pw.forward(X1 + 2 * X2).get()
    assert_close(y12, y1 + 2 * y2, 3 * GEMM, "pointwise linearity")
    bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
    yb = bn.forward(3.0 * X1 + 1.5).get()
    m, v = yb.mean(axis=(0, 2, 3)), yb.var(axis=(0, 2, 3))
    assert np.max(np.abs(m)) < 1e-4 and np.max(np.abs(v - 1.0)) < 1e-3
    dw = DepthwiseConvLayer("d", (C, 3, 3), stride=1, padding=1, with_bias=False)
    ya = dw.forward(X1).get().copy()
    dY = rng.standard_normal(ya.shape).astype(np.float32)
    dX = dw.backward(dY).get()
    yb2 = dw.forward(X1 + 1e-2 * X2).get()
    lhs = float(np.sum((yb2.astype(np.float64) - ya) * dY))  # <J dx, dY>
    rhs = float(np.sum(1e-2 * X2.astype(np.float64) * dX))    # <dx, J^T dY>
    assert abs(lhs - rhs) <= 2e-3 * abs(rhs)


@pytest.mark.parametrize("case", [(64, 512, 120), (8, 128, 12), (130, 64, 300), (16, 1024, 120), (64, 128, 10), (5, 64, 33)])
def test_dense_tcgen05_vs_oracle(O, case):
    """DenseLayer on the tensor-core path: fwd, dX, dW (+l2), db.  The last two cases have out_dim % 4 != 0 (MNIST's 10
    classes): W and dY are re-pitched into the workspace."""
    from dorknet_b200 import _lib
    from dorknet_b200.layers.dense_layer import DenseLayer
    from dorknet_b200.regularisers.l2 import l2
    B, D, K = case
    rng = np.random.default_rng(B + D + K)
    X = rng.standard_normal((B, D)).astype(np.float32)
    Wt = (rng.standard_normal((D, K)) / np.sqrt(D)).astype(np.float32)
    b = rng.standard_normal(K).astype(np.float32)
    dY = rng.standard_normal((B, K)).astype(np.float32)
    lay = DenseLayer("d", D, K, weight_regulariser=l2(1e-2))
    lay.learned_params["weights"], lay.learned_params["bias"] = Wt, b
    tc0, _ = _lib.gemm_call_counts()
    Y = lay.forward(X)
    assert_close(Y.get(), O.dense_fwd(X, Wt, b), GEMM, "Y")
    dX = lay.backward(dY)
    dXo, g = O.dense_bwd(dY, X, Wt, 1e-2, True)
    assert_close(dX.get(), dXo, GEMM, "dX")
    assert_close(lay.grads["weights"].get(), g["weights"], GEMM_W, "dW")
    assert_close(lay.grads["bias"].get(), g["bias"], FP32_RED, "db")
    tc1, _ = _lib.gemm_call_counts()
    assert tc1 - tc0 == 2  # forward + backward calls served by the tcgen05 kernels


BN_CASES = [
    # N, C, H, W : cluster size / load path the shape selects in bn_fused.cu
    (4, 8, 56, 56),     # S = 8, bulk copies, slices cut inside planes
    (64, 64, 28, 28),   # S = 8: every CTA owns whole planes
    (8, 256, 14, 14),   # S = 2
    (6, 512, 7, 7),     # S = 1, HW % 4 != 0: plain loads
    (3, 5, 9, 11),      # odd everything, tails
    (2, 3, 300, 300),   # slices too large for shared memory in backward: split kernels
    (5, 16, 1, 1),      # one value per (n, c)
    # channel-group kernels (bn_group.cu; fused=1): 4 channels of 7x7 / one of 14x14 per CTA, all images resident
    (64, 512, 7, 7), (64, 256, 14, 14), (3, 300, 6, 6), (5, 296, 5, 6), (32, 512, 1, 1), (70, 592, 3, 3),
]


@pytest.mark.parametrize("fused", [1, 2, 0])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("case", BN_CASES)
def test_batchnorm_vs_oracle(O, case, relu, fused):
    """BatchNorm forward / backward (with and without the fused ReLU) against the oracle, through the channel-group / cluster
    kernels (fused=1), the cluster kernels alone (fused=2) and through the split statistics / apply / reduce / dx kernels (fused=0)."""
    from dorknet_b200 import api
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.activations import ReLu
    N, C, H, W = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = (rng.standard_normal((N, C, H, W)) * rng.uniform(0.5, 3.0, (1, C, 1, 1)) + rng.uniform(-2, 2, (1, C, 1, 1))).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, (1, C, 1, 1)).astype(np.float32)
    dY = rng.standard_normal((N, C, H, W)).astype(np.float32)
    api.dk_tc_debug_set(9, fused)
    try:
        bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
        bn.learned_params["gamma"], bn.learned_params["beta"] = gamma, beta
        act = ReLu("r")
        Yo, cache, rm, rs = O.bn_fwd_train(X, gamma, beta, None, None)
        dXo_in = dY
        if relu:
            Y = act.forward(bn.forward(X))
            Yo_r, mask = O.relu_fwd(Yo)
            # the fused kernel evaluates x*scale + shift, the oracle gamma*x_hat + beta: values within rounding of 0
            # may land on either side of the ReLU
            safe = np.abs(Yo) > 1e-5 * np.max(np.abs(Yo))
            assert_close(np.where(safe, Y.get(), 0), np.where(safe, Yo_r, 0), FP32, "relu(Y)")
            # backward parity uses OUR mask: both directions of the kernel evaluate the same x*scale + shift > 0
            dXo_in = O.relu_bwd(dY, (Y.get() > 0).astype(np.float32))
        else:
            Y = bn.forward(X)
            assert_close(Y.get(), Yo, FP32, "Y")
        assert_close(bn.non_learned_params["running_mean"].get(), rm, FP32, "running_mean")
        assert_close(bn.non_learned_params["running_std"].get(), rs, FP32, "running_std")
        dXo, g = O.bn_bwd(dXo_in, gamma, cache)
        dX = bn.backward(act.backward(dY) if relu else dY)
        tol = 5 * FP32_RED if relu else FP32_RED
        assert_close(dX.get(), dXo, tol, "dX", atol=1e-6)
        assert_close(bn.grads["gamma"].get(), g["gamma"], tol, "dgamma", atol=1e-5)
        assert_close(bn.grads["beta"].get(), g["beta"], tol, "dbeta", atol=1e-5)
        # second batch: running statistics EMA
        X2 = (X * 0.5 + 1.0).astype(np.float32)
        _, _, rm2, rs2 = O.bn_fwd_train(X2, gamma, beta, rm, rs)
        bn.forward(X2)
        assert_close(bn.non_learned_params["running_mean"].get(), rm2, FP32, "running_mean 2")
        assert_close(bn.non_learned_params["running_std"].get(), rs2, FP32, "running_std 2")
    finally:
        api.dk_tc_debug_set(9, 1)


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("case", [(8, 64, 56, 56), (64, 128, 28, 28), (8, 256, 14, 14), (6, 512, 7, 7), (3, 5, 9, 11), (2, 3, 300, 300)])
def test_residual_join_folded_into_batchnorm(O, case, fuse):
    """ResidualBlock whose branch ends in a BatchNorm: relu(bn(x) + skip) as ONE pass of the BatchNorm kernel
    (cluster / channel-group kernels; split-kernel fallback for odd shapes) against the oracle, forward and backward."""
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.activations import ReLu
    from dorknet_b200.layers.residual_block import ResidualBlock
    N, C, H, W = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = (rng.standard_normal((N, C, H, W)) * 2.0 + 0.5).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, (1, C, 1, 1)).astype(np.float32)
    dY = rng.standard_normal((N, C, H, W)).astype(np.float32)
    bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
    bn.learned_params["gamma"], bn.learned_params["beta"] = gamma, beta
    blk = ResidualBlock("blk", [bn], None, ReLu("r"))
    blk.fuse_join = fuse
    Y = blk.forward(X)
    Yo, cache, rm, rs = O.bn_fwd_train(X, gamma, beta, None, None)
    pre = Yo + X
    safe = np.abs(pre) > 1e-5 * np.max(np.abs(pre))
    assert_close(np.where(safe, Y.get(), 0), np.where(safe, np.maximum(pre, 0), 0), FP32, "relu(bn(x) + x)")
    assert_close(bn.non_learned_params["running_mean"].get(), rm, FP32, "running_mean")
    d = dY * (Y.get() > 0)
    dXo, g = O.bn_bwd(d.astype(np.float32), gamma, cache)
    dX = blk.backward(dY)
    assert_close(dX.get(), dXo + d, FP32_RED, "dX", atol=1e-6)
    assert_close(bn.grads["gamma"].get(), g["gamma"], FP32_RED, "dgamma", atol=1e-5)
    assert_close(bn.grads["beta"].get(), g["beta"], FP32_RED, "dbeta", atol=1e-5)


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("case", [(4, 16, 16, 16, 24, 2), (64, 64, 112, 112, 64, 2), (3, 8, 15, 15, 12, 2), (2, 8, 12, 12, 8, 3)])
def test_batchnorm_relu_into_strided_pointwise(O, case, fuse):
    """BatchNorm -> ReLU -> PointwiseConvLayer(stride s): only the pixels the pointwise layer reads are normalised
    (dk_bn_apply_strided into a compact operand); forward, backward and a late read of the full-size ReLU output."""
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.activations import ReLu
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    N, C, H, W, F, s = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = (rng.standard_normal((N, C, H, W)) * 1.5 + 0.3).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, (1, C, 1, 1)).astype(np.float32)
    Wt = (rng.standard_normal((F, C)) / np.sqrt(C)).astype(np.float32)
    bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
    bn.learned_params["gamma"], bn.learned_params["beta"] = gamma, beta
    act = ReLu("r")
    pw = PointwiseConvLayer("p", stride=s, filter_block_shape=(F, C), with_bias=False)
    pw.learned_params["weights"] = Wt
    pw.fuse_strided_input = fuse
    A = act.forward(bn.forward(X))
    Y = pw.forward(A)
    Yb, cache, rm, rs = O.bn_fwd_train(X, gamma, beta, None, None)
    Ao, _ = O.relu_fwd(Yb)
    Yo, pcache = O.pointwise_fwd(Ao, Wt, None, s)
    assert_close(Y.get(), Yo, GEMM, "Y")
    assert_close(bn.non_learned_params["running_std"].get(), rs, FP32, "running_std")
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dAo, g = O.pointwise_bwd(dY, Wt, pcache, s, 0.0, False)
    pw.lazy_strided_dx = fuse
    dA = pw.backward(dY)
    assert_close(pw.grads["weights"].get(), g["weights"], GEMM_W, "dW")
    if dAo.shape == X.shape:  # (odd H: the reference's zero-stuffed dX is larger than X and its own backward fails)
        # BatchNorm backward FIRST: with fuse it takes the gradient in its compact form (dk_bn_bwd_strided, W % 8 == 0)
        dX = bn.backward(act.backward(dA))
        if fuse and s == 2 and W % 8 == 0:
            assert not dA.is_materialised
        assert_close(dA.get(), dAo, GEMM, "dA (zero-stuffed, read late)")
        ours_mask = (np.asarray(A.get()) > 0).astype(np.float32)  # late read of the full-size output
        # (the oracle's BatchNorm backward is fed OUR dA: the TF32 rounding of the dgrad GEMM is checked above, not here)
        dXo, gb = O.bn_bwd(O.relu_bwd(np.asarray(dA.get()), ours_mask), gamma, cache)
        assert_close(dX.get(), dXo, 5 * FP32_RED, "dX", atol=1e-6)
        assert_close(bn.grads["gamma"].get(), gb["gamma"], 5 * FP32_RED, "dgamma", atol=1e-5)
        assert_close(bn.grads["beta"].get(), gb["beta"], 5 * FP32_RED, "dbeta", atol=1e-5)
    else:
        assert_close(dA.get(), dAo, GEMM, "dA")
    safe = np.abs(Yb) > 1e-5 * np.max(np.abs(Yb))
    assert_close(np.where(safe, A.get(), 0), np.where(safe, Ao, 0), FP32, "full-size relu(bn(x)) read late")


@pytest.mark.parametrize("mixup", [False, True])
def test_input_pipeline_u8_nhwc(mixup):
    """uint8 NHWC batches -> fp32 NCHW - 128 (+ mixup) on the device (dk_input_u8_nhwc through HostBatchUploader)
    against the host arithmetic of the reference's loader (image_preprocessor.py:36-37, image_data_loader.py:100-110)."""
    from dorknet_b200 import workloads as W
    from dorknet_b200.input_pipeline import HostBatchUploader
    N, C, S, K = 5, 3, 33, 7
    raw = W.synthetic_batch_u8(N, C, S, K, seed=9, mixup=mixup)
    Xref, _, Yref = W.synthetic_batch(N, C, S, K, seed=9, mixup=mixup)
    up = HostBatchUploader((N, C, S, S), (N, K), slots=2)
    ha, hb, hy = up.pin_u8(raw["img"], raw["Y"], raw.get("img_b"))
    nbytes = up.submit_u8_from_pinned(ha, hb, hy, raw.get("lam", 0.0))
    Xd, Yd = up.get()
    assert nbytes == (2 if mixup else 1) * N * C * S * S + 4 * N * K
    assert_close(Xd.get(), Xref, 1e-6 if mixup else 0.0, "X")
    assert np.array_equal(Yd.get(), Yref)
    up.release()


FOLD_CASES = [
    # N, C, H, W, F, bias, zero_sum upstream
    (4, 64, 56, 56, 64, False, True),     # the 56 x 56 unit of ResNet-18-depsep
    (8, 128, 28, 28, 128, False, True),
    (8, 256, 14, 14, 512, False, True),   # F = 512: two n-blocks in wgrad, 256-column accumulators in dgrad
    (6, 512, 6, 6, 512, False, True),     # small planes (36 pixels): one partial M tile per image
    (6, 512, 7, 7, 512, False, True),     # 7 x 7 planes (49 pixels, nothing TMA-aligned): NOT folded, the unfused pair runs
    (3, 24, 10, 12, 40, True, False),     # bias + an upstream gradient with non-zero channel sums (S term), odd channel counts
    (2, 8, 5, 5, 8, False, False),        # P = 25: nothing TMA-aligned
]


@pytest.mark.parametrize("case", FOLD_CASES)
@pytest.mark.parametrize("backend", [0, 1])
def test_batchnorm_folded_into_pointwise_vs_oracle(O, case, backend):
    """BatchNorm (no ReLU) -> PointwiseConvLayer as one unit (bn_fold.cu): y, dW, the BatchNorm's dgamma / dbeta and
    its input gradient against the oracle's unfused chain (batch_norm.py:54-174 + pointwise_convolution.py:46-75).  The
    BatchNorm launches its statistics pass only; its backward kernel does not run.  backend 1 (fp32 SIMT GEMMs) pins the
    algebra tightly, backend 0 is the TF32 product path."""
    from dorknet_b200 import api
    from dorknet_b200.array import FoldedBNGrad, ZeroSumGrad, asarray
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    from dorknet_b200.regularisers.l2 import l2
    N, C, H, W, F, bias, zero_sum = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = (rng.standard_normal((N, C, H, W)) * rng.uniform(0.5, 3.0, (1, C, 1, 1)) + rng.uniform(-2, 2, (1, C, 1, 1))).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, (1, C, 1, 1)).astype(np.float32)
    Wt = (rng.standard_normal((F, C)) / np.sqrt(C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32) if bias else None
    dY = rng.standard_normal((N, F, H, W)).astype(np.float32)
    if zero_sum:
        dY -= dY.mean(axis=(0, 2, 3), keepdims=True)
    api.dk_set_gemm_backend(backend)
    try:
        bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
        bn.learned_params["gamma"], bn.learned_params["beta"] = gamma, beta
        pw = PointwiseConvLayer("pw", filter_block_shape=(F, C), with_bias=bias, weight_regulariser=l2(1e-2))
        pw.fold_bn_input = "always"  # (the default folds only when the statistics ride on a depthwise kernel: next test)
        pw.learned_params["weights"] = Wt
        if bias:
            pw.learned_params["bias"] = b
        Xh, cache, rm, rs = O.bn_fwd_train(X, gamma, beta, None, None)
        Yo, pcache = O.pointwise_fwd(Xh, Wt, b, 1)
        Y = pw.forward(bn.forward(X))
        g_tol, w_tol = (GEMM, GEMM_W) if backend == 0 else (FP32_RED, FP32_RED)
        if (H * W) % 4 != 0:
            assert pw._folded_bn is None
            assert_close(Y.get(), Yo, g_tol, "Y (not foldable)")
            return
        assert pw._folded_bn is bn
        assert_close(Y.get(), Yo, g_tol, "Y")
        assert_close(bn.non_learned_params["running_mean"].get(), rm, FP32, "running_mean")
        assert_close(bn.non_learned_params["running_std"].get(), rs, FP32, "running_std")
        dXh_o, pg = O.pointwise_bwd(dY, Wt, pcache, 1, l2_strength=1e-2, with_bias=bias)
        dXo, bg = O.bn_bwd(dXh_o, gamma, cache)
        up = asarray(dY)
        if zero_sum:
            up = ZeroSumGrad(up.t, up.shape)
        back = pw.backward(up)
        assert isinstance(back, FoldedBNGrad) and not back.is_materialised
        dX = bn.backward(back)
        assert not back.is_materialised, "the BatchNorm must take the folded gradient, not trigger the plain dgrad"
        # the BatchNorm-critical sums are differences of TF32 GEMM results: gates relative to the largest entry, with the
        # wgrad tolerance (they ARE wgrad results); fp32 backend: reduction tolerance
        assert_close(pw.grads["weights"].get(), pg["weights"], w_tol, "dW")
        if bias:
            assert_close(pw.grads["bias"].get(), pg["bias"], FP32_RED, "db")
        scale_g = float(np.max(np.abs(bg["gamma"])))
        assert_close(bn.grads["gamma"].get(), bg["gamma"], w_tol, "dgamma", atol=w_tol * scale_g)
        assert_close(bn.grads["beta"].get(), bg["beta"], w_tol, "dbeta", atol=w_tol * max(scale_g, float(np.max(np.abs(bg["beta"])))))
        assert_close(dX.get(), dXo, 2 * g_tol, "dX", atol=1e-6)
        # a late reader of either deferred object gets the reference's tensors
        assert_close(back.get(), dXh_o, g_tol, "dX_hat (late reader)")
        # and the unfused pair agrees too
        pw.fold_bn_input = False
        Y2 = pw.forward(bn.forward(X))
        assert pw._folded_bn is None
        assert_close(Y2.get(), Yo, g_tol, "Y unfused")
    finally:
        api.dk_set_gemm_backend(0)


DW_BN_PW_CASES = [
    # N, C, H, W, F
    (4, 64, 56, 56, 64),     # the first units of ResNet-18-depsep: 14 strips per row, segments cut by warp boundaries
    (8, 128, 28, 28, 128),
    (3, 40, 24, 32, 72),     # odd channel counts, rectangular
    (2, 16, 64, 64, 16),     # several bands per plane
]


@pytest.mark.parametrize("case", DW_BN_PW_CASES)
@pytest.mark.parametrize("backend", [0, 1])
def test_depthwise_batchnorm_pointwise_unit_fused_vs_oracle(O, case, backend):
    """dw3x3 -> BatchNorm -> pointwise, the product default: the depthwise forward kernel also produces the BatchNorm
    statistics (dk_dwconv_fwd_bn), the normalisation is folded into the pointwise GEMMs (bn_fold.cu) and the BatchNorm's
    backward into the dgrad epilogue -- against the oracle's three separate layers, forward, running statistics and every
    gradient down to the depthwise input."""
    from dorknet_b200 import api
    from dorknet_b200.array import LazyDWOutput
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.depthwise_convolution import DepthwiseConvLayer
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    N, C, H, W, F = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = np.maximum(rng.standard_normal((N, C, H, W)) + 0.5, 0).astype(np.float32)  # post-ReLU-like: positive mean
    Wd = (rng.standard_normal((C, 3, 3)) / 3.0 + 0.1).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, (1, C, 1, 1)).astype(np.float32)
    beta = rng.uniform(-0.5, 0.5, (1, C, 1, 1)).astype(np.float32)
    Wp = (rng.standard_normal((F, C)) / np.sqrt(C)).astype(np.float32)
    dY = rng.standard_normal((N, F, H, W)).astype(np.float32)
    dY -= dY.mean(axis=(0, 2, 3), keepdims=True)
    api.dk_set_gemm_backend(backend)
    try:
        dw = DepthwiseConvLayer("dw", filter_block_shape=(C, 3, 3), stride=1, padding=1, with_bias=False)
        dw.learned_params["weights"] = Wd
        bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
        bn.learned_params["gamma"], bn.learned_params["beta"] = gamma, beta
        pw = PointwiseConvLayer("pw", filter_block_shape=(F, C), with_bias=False)
        pw.learned_params["weights"] = Wp
        Do, dcache = O.depthwise_fwd(X, Wd, None, 1, 1)
        Xh, cache, rm, rs = O.bn_fwd_train(Do, gamma, beta, None, None)
        Yo, pcache = O.pointwise_fwd(Xh, Wp, None, 1)
        d_out = dw.forward(X)
        assert isinstance(d_out, LazyDWOutput) and not d_out.is_materialised
        Y = pw.forward(bn.forward(d_out))
        assert pw._folded_bn is bn and d_out.is_materialised
        g_tol, w_tol = (GEMM, GEMM_W) if backend == 0 else (FP32_RED, FP32_RED)
        assert_close(d_out.get(), Do, FP32, "depthwise output")
        assert_close(bn.non_learned_params["running_mean"].get(), rm, FP32, "running_mean")
        assert_close(bn.non_learned_params["running_std"].get(), rs, FP32_RED, "running_std")
        assert_close(Y.get(), Yo, g_tol, "Y")
        # the mean of every output channel is exact although the GEMM multiplied un-centred, truncated activations
        ym, yo_m = Y.get().mean(axis=(0, 2, 3)), Yo.mean(axis=(0, 2, 3))
        assert np.abs(ym - yo_m).max() <= 2e-5 * Yo.std(), "output mean %g vs %g" % (np.abs(ym - yo_m).max(), Yo.std())
        dXh_o, pg = O.pointwise_bwd(dY, Wp, pcache, 1)
        dDo, bg = O.bn_bwd(dXh_o, gamma, cache)
        dXo, dg = O.depthwise_bwd(dDo, Wd, dcache, 1, 1)
        dX = dw.backward(bn.backward(pw.backward(dY)))
        assert_close(pw.grads["weights"].get(), pg["weights"], w_tol, "dW pointwise")
        scale_g = float(np.max(np.abs(bg["gamma"])))
        assert_close(bn.grads["gamma"].get(), bg["gamma"], w_tol, "dgamma", atol=w_tol * scale_g)
        assert_close(bn.grads["beta"].get(), bg["beta"], w_tol, "dbeta", atol=w_tol * scale_g)
        assert_close(dX.get(), dXo, 2 * g_tol, "dX")
        assert_close(dw.grads["weights"].get(), dg["weights"], 2 * w_tol, "dW depthwise")
        # second step: running statistics EMA through the fused path
        X2 = (0.5 * X + 0.25).astype(np.float32)
        Do2, _ = O.depthwise_fwd(X2, Wd, None, 1, 1)
        _, _, rm2, rs2 = O.bn_fwd_train(Do2, gamma, beta, rm, rs)
        pw.forward(bn.forward(dw.forward(X2)))
        assert_close(bn.non_learned_params["running_mean"].get(), rm2, FP32, "running_mean 2")
        assert_close(bn.non_learned_params["running_std"].get(), rs2, FP32_RED, "running_std 2")
        # a lazy depthwise output read by anybody else is the plain forward
        assert_close(dw.forward(X).get(), Do, FP32, "plain deferred forward")
    finally:
        api.dk_set_gemm_backend(0)


@pytest.mark.parametrize("case", [(2, 8, 12, 12, 3, 1, 1, 1), (3, 6, 15, 17, 3, 2, 1, 0), (2, 4, 9, 9, 5, 1, 2, 1)])
def test_depthwise_input_transform_on_load_vs_oracle(O, case):
    """dk_dwconv_fwd / dk_dwconv_bwd with in_scale / in_shift / in_relu (include/dorknet_b200.h: the input is read as
    relu?(x*scale[c] + shift[c]), a producer's deferred BatchNorm(+ReLU) applied on load; zero padding AFTER it) against the
    oracle run on the transformed tensor: forward, dX (gradient with respect to the TRANSFORMED input), dW."""
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty
    N, C, H, W, k, s, p, relu = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wd = rng.standard_normal((C, k, k)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, C).astype(np.float32)
    shift = rng.uniform(-0.5, 0.5, C).astype(np.float32)
    Xt = X * scale[None, :, None, None] + shift[None, :, None, None]
    if relu:
        Xt = np.maximum(Xt, 0)
    Xt = Xt.astype(np.float32)
    Yo, cache = O.depthwise_fwd(Xt, Wd, None, s, p)
    x, w, sc, sh = asarray(X), asarray(Wd), asarray(scale), asarray(shift)
    y = empty(Yo.shape)
    st = runtime.stream()
    api.dk_dwconv_fwd(x.ptr, w.ptr, None, y.ptr, sc.ptr, sh.ptr, relu, N, C, H, W, k, k, s, p, st)
    assert_close(y.get(), Yo, FP32, "Y")
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dXo, g = O.depthwise_bwd(dY, Wd, cache, s, p)
    dy, dx, dw = asarray(dY), empty(X.shape), empty(Wd.shape)
    nb = api.dk_dwconv_ws_bytes(N, C, H, W, k, k, s, p)
    ws, wsn = runtime.scratch(nb)
    api.dk_dwconv_bwd(dy.ptr, x.ptr, w.ptr, dx.ptr, dw.ptr, None, sc.ptr, sh.ptr, relu, None, 0.0, N, C, H, W, k, k, s, p,
                      ws, wsn, st)
    assert_close(dx.get(), dXo, FP32_RED, "dX")
    assert_close(dw.get(), g["weights"], FP32_RED, "dW")


@pytest.mark.parametrize("case", [(6, 512, 7, 7, 512, 1), (4, 256, 14, 14, 512, 2), (3, 64, 56, 56, 128, 2), (5, 24, 7, 7, 40, 1)])
def test_pointwise_packed_operands_match_unpacked(O, case):
    """PointwiseConvLayer keeping ONE re-pitched copy of x (forward + wgrad) and of dY (dgrad + wgrad) through dk_pw_pack /
    dk_pwconv_*_packed gives the results of the entry points that re-pitch per call (same operands and products; the split-K
    plan of a strided wgrad may differ, so equal to fp32 summation order: 1e-5), and both match the oracle."""
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    N, C, H, W, F, s = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    X = rng.standard_normal((N, C, H, W)).astype(np.float32)
    Wt = (rng.standard_normal((F, C)) / np.sqrt(C)).astype(np.float32)
    Yo, cache = O.pointwise_fwd(X, Wt, None, s)
    dY = rng.standard_normal(Yo.shape).astype(np.float32)
    dXo, g = O.pointwise_bwd(dY, Wt, cache, s)
    res = []
    for pack in (True, False):
        pw = PointwiseConvLayer("p", stride=s, filter_block_shape=(F, C), with_bias=False)
        pw.learned_params["weights"] = Wt
        pw.pack_operands = pack
        y = pw.forward(X).get().copy()
        assert (pw._x_pack is not None) == pack
        dx = pw.backward(dY).get().copy()
        res.append((y, dx, pw.grads["weights"].get().copy()))
        assert_close(y, Yo, GEMM, "Y")
        assert_close(dx, dXo, GEMM, "dX")
        assert_close(res[-1][2], g["weights"], GEMM_W, "dW")
    for a, b, what in zip(res[0], res[1], ("Y", "dX", "dW")):
        assert_close(a, b, FP32, what + " packed vs unpacked")
