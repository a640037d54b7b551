"""PointwiseConvLayer (reference: layers/pointwise_convolution.py:6-129)."""
import os

import numpy as np

from .layer import Layer, api, runtime, asarray
from ..array import FoldedBNGrad, LazyBNOutput, LazyReluOutput, LazyStridedGrad, ZeroSumGrad


class PointwiseConvLayer(Layer):
    """1x1 convolution.  The reference transposes to NHWC, calls a GEMM and transposes back
    (pointwise_convolution.py:46-55); here the GEMM runs directly on the NCHW tensor
    (Y[n] = W[F,C] . X[n][C,HW]), so no transposed copy exists in either direction."""
    _h5_attrs = ("with_bias", "num_filters", "num_channels", "stride")
    _h5_optional_attrs = {"stride": 1}  # layers/pointwise_convolution.py:111-115
    _h5_params = ("weights", "bias")  # layers/pointwise_convolution.py:77-130


    def __init__(self, layer_name, stride=1, filter_block_shape=None, with_bias=True,
                 weight_regulariser=None, weight_initialiser="normal"):
        super().__init__(layer_name)
        self.stride = stride
        self.with_bias = with_bias
        self.weight_regulariser = weight_regulariser
        self.weight_initialiser = weight_initialiser
        if filter_block_shape is not None:
            self.num_filters, self.num_channels = filter_block_shape
            if self.weight_initialiser == "glorot_uniform":
                limit = np.sqrt(6.0 / (self.num_channels + self.num_filters))
                weights = np.random.uniform(low=-limit, high=limit, size=filter_block_shape).astype(np.float32)
            elif self.weight_initialiser == "normal":
                weights = 0.01 * np.random.randn(*filter_block_shape).astype(np.float32)
            else:
                raise ValueError("unknown weight_initialiser {!r}".format(weight_initialiser))
            self.learned_params = {"weights": weights}
            self.grads = {"weights": np.zeros_like(weights).astype(np.float32)}
            if with_bias:
                bias = np.zeros(self.num_filters).astype(np.float32)
                self.learned_params.update({"bias": bias})
                self.grads.update({"bias": np.zeros_like(bias)})
        else:
            self.num_filters = None
            self.learned_params = {}
            self.grads = {}
        self._x = None
        self._xgeom = None
        self.lazy_strided_dx = True  # stride > 1: return the input gradient as a LazyStridedGrad
        self.fuse_strided_input = True  # BatchNorm -> ReLU -> stride-s input: normalise only the pixels this layer reads
        # BatchNorm (no ReLU) -> this layer, stride 1: fold the normalisation into the GEMMs (bn_fold.cu).  True: when the
        # BatchNorm's statistics ride on the depthwise kernel before it (always a win); "always": whenever possible; False: never
        self.fold_bn_input = {"0": False, "always": "always"}.get(os.environ.get("DK_FOLD_BN", "1"), True)
        self.trust_zero_sum = os.environ.get("DK_FOLD_ZEROSUM", "1") != "0"  # skip the channel sums of a BatchNorm's input gradient
        self.pack_operands = os.environ.get("DK_PW_PACK", "1") != "0"  # keep re-pitched copies of 7x7 / small strided operands
        self._x_pack = None
        self._folded_bn = None

    def __repr__(self):
        out = "PointwiseConvLayer({}, ".format(self.layer_name)
        if self.num_filters is not None:
            out += "filter_block_shape=({}, {}), ".format(self.num_filters, self.num_channels)
        out += "stride={}, with_bias={}, weight_regulariser={}, is_on_gpu={})".format(
            self.stride, self.with_bias, repr(self.weight_regulariser), self.is_on_gpu)
        return out

    def forward(self, X, test_mode=False):
        self._ensure_gpu()
        X = asarray(X)
        N, C, H, W = X.shape
        if C != self.num_channels:
            raise ValueError("PointwiseConvLayer {}: input has {} channels, filters expect {}".format(
                self.layer_name, C, self.num_channels))
        s = int(self.stride)
        OH, OW = (H - 1) // s + 1, (W - 1) // s + 1  # X[:, :, ::s, ::s]
        self.input_shape = X.shape
        y = self._buf("y", (N, self.num_filters, OH, OW))
        bias = self._param("bias").ptr if self.with_bias else None
        self._folded_bn = None
        if (s == 1 and self.fold_bn_input and not test_mode and isinstance(X, LazyBNOutput) and X.fusable
                and (H * W) % 4 == 0 and X.bn.can_fold(self.fold_bn_input)):
            # BatchNorm -> pointwise with nothing in between: y = (W diag(scale)) . x + W . shift.  The BatchNorm only
            # produces its statistics (inside the depthwise kernel that feeds it); this GEMM reads the BatchNorm's INPUT and
            # the normalised tensor is never produced (unless somebody else reads it: X.consume() leaves a late-reader thunk)
            bn = X.bn
            x_raw, saved, has_resid = bn.fold_statistics()
            X.consume()
            wf, bf = self._buf("w_fold", (self.num_filters, C)), self._buf("b_fold", (self.num_filters,))
            api.dk_bn_fold_fwd(self._param("weights").ptr, bias, saved + 8 * C, saved + 12 * C, saved,
                               saved + 16 * C if has_resid else None, wf.ptr, bf.ptr, self.num_filters, C, runtime.stream())
            self._folded_bn = bn
            self._x, self._xgeom = x_raw, (H, W, 1)
            ws, wsn = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, H, W, self.num_filters, 1))
            api.dk_pwconv_fwd(x_raw.ptr, wf.ptr, bf.ptr, y.ptr, N, C, H, W, self.num_filters, 1, ws, wsn, runtime.stream())
            return y
        if (s > 1 and self.fuse_strided_input and not test_mode and isinstance(X, LazyReluOutput)
                and not X.is_materialised and X.bn._pending is not None):
            # BatchNorm -> ReLU -> X[:, :, ::s, ::s]: only the pixels this layer reads are normalised, into a compact
            # [N, C, OH, OW] operand that both the forward and the wgrad GEMM take through TMA with stride 1
            xc = self._buf("x_sub", (N, C, OH, OW))
            bn = X.bn
            bn.fused_relu_apply_strided(xc, s)
            X.replace_thunk(lambda: bn.apply_saved(X._buf, 1))  # the full-size output, only if somebody else reads it
            self._x, self._xgeom = xc, (OH, OW, 1)
        else:
            self._x, self._xgeom = X, (H, W, s)
        xh, xw, xs = self._xgeom
        ws, wsn = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, xh, xw, self.num_filters, xs))
        self._x_pack = None
        # (the packed entry points are tensor-core only: they need the weight rows, C floats, to be TMA-addressable)
        pb = api.dk_pw_pack_bytes(N, C, xh, xw, xs) if (self.pack_operands and C % 4 == 0) else 0
        if pb:
            # 7x7 planes / small strided planes: TMA needs a re-pitched copy.  It is made once here and kept for the wgrad
            # GEMM of this step (dk_pwconv_fwd / wgrad would each make their own in the workspace)
            xp = self._buf("x_pack", ((pb + 3) // 4,))
            api.dk_pw_pack(self._x.ptr, xp.ptr, N, C, xh, xw, xs, runtime.stream())
            api.dk_pwconv_fwd_packed(xp.ptr, self._param("weights").ptr, bias, y.ptr, N, C, OH, OW, self.num_filters,
                                     ws, wsn, runtime.stream())
            if not test_mode:
                self._x_pack = xp
            return y
        api.dk_pwconv_fwd(self._x.ptr, self._param("weights").ptr, bias, y.ptr, N, C, xh, xw, self.num_filters, xs,
                          ws, wsn, runtime.stream())
        return y

    def backward(self, upstream_dx):
        dY = asarray(upstream_dx)
        N, C, H, W = self.input_shape
        _, F, OH, OW = dY.shape
        s = int(self.stride)
        w = self._param("weights")
        st = runtime.stream()
        ws, wsn = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, max(H, OH * s), max(W, OW * s), F, s))
        dbias = self._grad("bias").ptr if self.with_bias else None
        if self._folded_bn is not None:
            return self._backward_folded(dY, N, C, H, W, F, ws, wsn, st)
        xh, xw, xs = self._xgeom  # (a compact, already subsampled operand has stride 1)
        dyp = None  # re-pitched copy of dY, when this shape needs one
        if runtime.side_enabled():
            # nothing on the backward chain needs dW: inside a captured step the wgrad GEMM (and its reduce / re-pitch
            # kernels) runs on the side stream next to the chain (runtime.side_region), on its own scratch buffer
            dy_ptr, x_ptr, dw_ptr, l2s = dY.ptr, self._x.ptr, self._grad("weights").ptr, self._l2_strength()
            nb = api.dk_pwconv_ws_bytes(N, C, max(H, OH * s), max(W, OW * s), F, s)

            def wgrad():
                ws2, wsn2 = runtime.side_scratch(nb)
                api.dk_pwconv_wgrad(dy_ptr, x_ptr, w.ptr, dw_ptr, dbias, l2s, N, C, xh, xw, F, xs, ws2, wsn2, runtime.stream())
            runtime.side_launch(wgrad)
        else:
            # dY planes that TMA cannot take as they lie (7x7) are re-pitched once for wgrad AND dgrad; x may already have
            # its packed copy from forward
            gb = api.dk_pw_pack_bytes(N, F, OH, OW, 1) if (self.pack_operands and C % 4 == 0) else 0
            if gb:
                dyp = self._buf("dy_pack", ((gb + 3) // 4,))
                api.dk_pw_pack(dY.ptr, dyp.ptr, N, F, OH, OW, 1, st)
            xpk = getattr(self, "_x_pack", None)
            if dyp is not None or xpk is not None:
                if dbias is not None:
                    api.dk_bias_grad(dY.ptr, dbias, N, F, OH * OW, None, 0, st)
                api.dk_pwconv_wgrad_packed(dyp.ptr if dyp is not None else dY.ptr, int(dyp is not None),
                                           xpk.ptr if xpk is not None else self._x.ptr, int(xpk is not None), w.ptr,
                                           self._grad("weights").ptr, None, self._l2_strength(), N, C, xh, xw, F, xs,
                                           ws, wsn, st)
            else:
                api.dk_pwconv_wgrad(dY.ptr, self._x.ptr, w.ptr, self._grad("weights").ptr, dbias, self._l2_strength(),
                                    N, C, xh, xw, F, xs, ws, wsn, st)

        def dgrad(out, stride, ws_, wsn_):
            if dyp is not None:
                api.dk_pwconv_dgrad_packed(dyp.ptr, w.ptr, out.ptr, N, C, OH, OW, F, stride, ws_, wsn_, runtime.stream())
            else:
                api.dk_pwconv_dgrad(dY.ptr, w.ptr, out.ptr, N, C, OH, OW, F, stride, ws_, wsn_, runtime.stream())
        # zero-stuffed dx of shape (OH*s, OW*s): for odd H this is NOT the input shape
        # (pointwise_convolution.py:68-72) -- reproduced on purpose
        dx = self._buf("dx", (N, C, OH * s, OW * s))
        if s > 1 and self.lazy_strided_dx:
            # deferred: a consumer that only needs the non-zero entries (BatchNormLayer.backward) asks for the compact
            # form; anybody else reads the zero-stuffed tensor as before
            def full():
                ws2, wsn2 = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, max(H, OH * s), max(W, OW * s), F, s))
                dgrad(dx, s, ws2, wsn2)

            def compact():
                dxs = self._buf("dx_sub", (N, C, OH, OW))
                ws2, wsn2 = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, OH, OW, F, 1))
                dgrad(dxs, 1, ws2, wsn2)
                return dxs
            return LazyStridedGrad(dx, full, s, compact)
        dgrad(dx, s, ws, wsn)
        return dx

    def _backward_folded(self, dY, N, C, H, W, F, ws, wsn, st):
        """Backward of (BatchNorm -> this layer) as ONE unit (bn_fold.cu): the raw wgrad GEMM dY . x^T, a [F, C]-sized
        kernel that turns it into dW, the BatchNorm's dgamma / dbeta and the coefficients of its input gradient, and the
        dgrad GEMM on the folded weights whose epilogue adds cb*x + cd.  No pass over the activations besides the two
        GEMMs; the BatchNorm's own backward kernel does not run."""
        bn, w = self._folded_bn, self._param("weights")
        saved = bn._bufs["saved"].ptr
        g_raw = self._buf("g_raw", (F, C))
        api.dk_pwconv_wgrad(dY.ptr, self._x.ptr, w.ptr, g_raw.ptr, None, 0.0, N, C, H, W, F, 1, ws, wsn, st)
        s_col = None
        if self.with_bias or not isinstance(dY, ZeroSumGrad) or not self.trust_zero_sum:
            # channel sums of dY: the bias gradient, and the shift term of the folded wgrad.  A BatchNorm's input gradient
            # sums to zero per channel (array.ZeroSumGrad), which is what arrives here in the depthwise-separable unit
            sbuf = self._grad("bias") if self.with_bias else self._buf("s_col", (F,))
            api.dk_bias_grad(dY.ptr, sbuf.ptr, N, F, H * W, None, 0, st)
            s_col = sbuf.ptr
        coef = self._buf("bn_coef", (2, C))
        api.dk_bn_fold_bwd(g_raw.ptr, s_col, w.ptr, bn._param("gamma").ptr, saved, saved + 4 * C, saved + 8 * C,
                           saved + 12 * C, self._l2_strength(), N * H * W, self._grad("weights").ptr,
                           bn._grad("gamma").ptr, bn._grad("beta").ptr, coef.ptr, coef.ptr + 4 * C, F, C, st)
        bn_dx = bn._buf("dx", bn.input_shape)
        api.dk_pwconv_dgrad_affine(dY.ptr, self._bufs["w_fold"].ptr, self._x.ptr, coef.ptr, coef.ptr + 4 * C, bn_dx.ptr,
                                   N, C, H, W, F, ws, wsn, st)
        dx = self._buf("dx", (N, C, H, W))

        def plain():  # somebody other than the folded BatchNorm reads our result: the reference's dX = W^T dY
            ws2, wsn2 = runtime.scratch(api.dk_pwconv_ws_bytes(N, C, H, W, F, 1))
            api.dk_pwconv_dgrad(dY.ptr, w.ptr, dx.ptr, N, C, H, W, F, 1, ws2, wsn2, runtime.stream())
        return FoldedBNGrad(dx, plain, bn, ZeroSumGrad(bn_dx.t, bn_dx.shape))
