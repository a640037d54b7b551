"""DepthwiseConvLayer (reference: layers/depthwise_convolution.py:10-352, layers/im2col.pyx:109-178)."""
import os

import numpy as np

from .layer import Layer, api, runtime, asarray
from ..array import LazyDWOutput


class DepthwiseConvLayer(Layer):
    """filter_block_shape = (num_incoming_channels, num_filter_rows, num_filter_cols)"""
    _h5_attrs = ("stride", "padding", "with_bias", "num_filters", "f_rows", "f_cols")
    _h5_params = ("weights", "bias")  # layers/depthwise_convolution.py:300-353


    def __init__(self, layer_name, filter_block_shape=None, stride=1, padding=1, with_bias=True,
                 weight_regulariser=None, weight_initialiser="normal"):
        super().__init__(layer_name)
        self.stride = stride
        self.padding = padding
        self.with_bias = with_bias
        self.weight_regulariser = weight_regulariser
        self.weight_initialiser = weight_initialiser
        if filter_block_shape is not None:
            self.num_filters, self.f_rows, self.f_cols = filter_block_shape
            if self.weight_initialiser == "glorot_uniform":
                limit = np.sqrt(6.0 / (2 * self.num_filters))
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
                self.grads.update({"bias": np.zeros_like(bias, dtype=np.float32)})
        else:
            self.num_filters = None
            self.learned_params = {}
            self.grads = {}
        self._x = None
        self.defer_forward = os.environ.get("DK_DW_DEFER", "1") != "0"  # training: return a lazy output so that a folded BatchNorm can ride on the forward kernel

    def __repr__(self):
        out = "DepthwiseConvLayer({}, ".format(self.layer_name)
        if self.num_filters is not None:
            out += "filter_block_shape=({}, {}, {}), ".format(self.num_filters, self.f_rows, self.f_cols)
        out += "stride={}, padding={}, with_bias={}, weight_regulariser={})".format(
            self.stride, self.padding, self.with_bias, repr(self.weight_regulariser))
        return out

    def forward(self, X, test_mode=False):
        """depthwise_convolution.py:72-83.  The reference caches the PADDED input for backward; here
        padding is index arithmetic inside the kernels and only a reference to X is kept."""
        self._ensure_gpu()
        X = asarray(X)
        N, C, H, W = X.shape
        if C != self.num_filters:
            raise ValueError("DepthwiseConvLayer {}: input has {} channels, filters expect {}".format(
                self.layer_name, C, self.num_filters))
        s, p = int(self.stride), int(self.padding)
        self.num_row_patches = ((H + 2 * p - self.f_rows) / s) + 1
        self.num_col_patches = ((W + 2 * p - self.f_cols) / s) + 1
        self.input_shape = X.shape
        y = self._buf("y", (N, C, int(self.num_row_patches), int(self.num_col_patches)))
        bias = self._param("bias").ptr if self.with_bias else None

        def plain():
            api.dk_dwconv_fwd(X.ptr, self._param("weights").ptr, bias, y.ptr, None, None, 0,
                              N, C, H, W, self.f_rows, self.f_cols, s, p, runtime.stream())
        if not test_mode:
            self._x = X
            if self.defer_forward and api.dk_dwconv_fwd_bn_ws_bytes(N, C, H, W, self.f_rows, self.f_cols, s, p) > 0:
                # nothing is launched until the consumer is known: a BatchNorm that gets folded into the pointwise layer
                # behind it takes the variant that also produces its statistics (forward_with_bn_statistics)
                return LazyDWOutput(y, plain, self)
        plain()
        return y

    def forward_with_bn_statistics(self, out, gamma, beta, running_mean, running_std, first_batch, momentum, eps, saved):
        """The deferred forward of `out` (a LazyDWOutput of this layer) + the training statistics of the BatchNorm that
        consumes it, in one pass (dk_dwconv_fwd_bn).  `saved`: device pointer to the BatchNorm's [5, C] block
        (mean, invstd, scale, shift, truncation residual)."""
        if out.layer is not self or out.is_materialised:
            raise RuntimeError("DepthwiseConvLayer {}: output already produced".format(self.layer_name))
        N, C, H, W = self.input_shape
        s, p = int(self.stride), int(self.padding)
        bias = self._param("bias").ptr if self.with_bias else None
        x_ptr = self._x.ptr  # (may launch the producer's own deferred kernels: before ours, in stream order)
        ws, wsn = runtime.scratch(api.dk_dwconv_fwd_bn_ws_bytes(N, C, H, W, self.f_rows, self.f_cols, s, p))
        out.launched_by_consumer()
        api.dk_dwconv_fwd_bn(x_ptr, self._param("weights").ptr, bias, out._buf.ptr, N, C, H, W, self.f_rows, self.f_cols, s, p,
                             gamma, beta, running_mean, running_std, int(first_batch), float(momentum), float(eps),
                             saved, saved + 4 * C, saved + 8 * C, saved + 12 * C, saved + 16 * C, ws, wsn, runtime.stream())

    def backward(self, upstream_dx, dx_add=None):
        """depthwise_convolution.py:186-196: fused dX + per-plane dW partials, summed over images.
        `dx_add` (optional) is added into dx: lets a ResidualBlock fold its skip-path gradient in."""
        dY = asarray(upstream_dx)
        N, C, H, W = self.input_shape
        s, p = int(self.stride), int(self.padding)
        dx = self._buf("dx", self.input_shape)
        nbytes = api.dk_dwconv_ws_bytes(N, C, H, W, self.f_rows, self.f_cols, s, p)
        ws, wsn = runtime.scratch(nbytes)
        dbias = self._grad("bias").ptr if self.with_bias else None
        api.dk_dwconv_bwd(dY.ptr, self._x.ptr, self._param("weights").ptr, dx.ptr, self._grad("weights").ptr, dbias,
                          None, None, 0, dx_add.ptr if dx_add is not None else None, self._l2_strength(),
                          N, C, H, W, self.f_rows, self.f_cols, s, p, ws, wsn, runtime.stream())
        return dx
