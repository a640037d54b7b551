"""ConvLayer (reference: layers/convolution.py:12-281, layers/im2col.pyx:16-36,209-234)."""
import numpy as np

from .layer import Layer, api, runtime, asarray
from ..array import LazyDeviceArray


class ConvLayer(Layer):
    """Dense k x k convolution as an implicit GEMM: pad, im2col and the NHWC->NCHW transpose of
    convolution.py:58-87 never materialise; wgrad / dgrad replace the two GEMMs + row2im of :90-126."""
    _h5_attrs = ("with_bias", "num_filters", "filter_chans", "f_rows", "f_cols", "stride", "padding")
    _h5_params = ("weights", "bias")  # layers/convolution.py:226-281


    def __init__(self, layer_name, filter_block_shape=None, stride=1, padding=1,
                 with_bias=True, weight_regulariser=None, weight_initialiser="normal"):
        super().__init__(layer_name)
        self.stride = stride
        self.padding = padding
        self.patches = None  # never materialised here (see im2col_materialise for the debug view)
        self.weight_regulariser = weight_regulariser
        self.weight_initialiser = weight_initialiser
        self.with_bias = with_bias
        # "lazy": dX is computed when first read (a first layer's dX is dropped by the container, so its dgrad
        # never runs); True: always compute; False: return None
        self.needs_input_grad = "lazy"
        if filter_block_shape:
            self.num_filters, self.filter_chans, self.f_rows, self.f_cols = filter_block_shape
            if self.weight_initialiser == "glorot_uniform":
                limit = np.sqrt(6.0 / (self.filter_chans + self.num_filters))
                weights = np.random.uniform(low=-limit, high=limit, size=filter_block_shape).astype(np.float32)
            elif self.weight_initialiser == "normal":
                weights = 0.01 * np.random.randn(*filter_block_shape).astype(np.float32)
            else:
                raise ValueError("unknown weight_initialiser {!r}".format(weight_initialiser))
            self.learned_params = {"weights": weights}
            self.grads = {"weights": np.zeros_like(weights)}
            if with_bias:
                bias = np.zeros(self.num_filters).astype(np.float32)
                self.learned_params.update({"bias": bias})
                self.grads.update({"bias": np.zeros_like(bias)})
        else:
            self.num_filters = None
            self.learned_params = {}
            self.grads = {}
        self._x = None

    def __repr__(self):
        out = "ConvLayer({}, ".format(self.layer_name)
        if self.num_filters is not None:
            out += "filter_block_shape=({},{},{},{}), ".format(self.num_filters, self.filter_chans,
                                                               self.f_rows, self.f_rows)
        out += "stride={}, padding={}, with_bias={}, weight_regulariser={})".format(
            self.stride, self.padding, self.with_bias, self.weight_regulariser)
        return out

    def _geom(self, shape):
        N, C, H, W = shape
        if C != self.filter_chans:
            raise ValueError("ConvLayer {}: input has {} channels, filters expect {}".format(
                self.layer_name, C, self.filter_chans))
        # reference keeps the un-truncated float patch counts (convolution.py:67-68)
        self.num_row_patches = ((H + 2 * self.padding - self.f_rows) / self.stride) + 1
        self.num_col_patches = ((W + 2 * self.padding - self.f_cols) / self.stride) + 1
        return N, C, H, W, int(self.num_row_patches), int(self.num_col_patches)

    def _scratch(self, N, C, H, W):
        return runtime.scratch(api.dk_conv2d_ws_bytes(N, C, H, W, self.num_filters, self.f_rows, self.f_cols,
                                                      self.stride, self.padding))

    def forward(self, X, test_mode=False):
        self._ensure_gpu()
        X = asarray(X)
        self.input_shape = X.shape
        N, C, H, W, OH, OW = self._geom(X.shape)
        y = self._buf("y", (N, self.num_filters, OH, OW))
        bias = self._param("bias").ptr if self.with_bias else None
        ws, wsn = self._scratch(N, C, H, W)
        api.dk_conv2d_fwd(X.ptr, self._param("weights").ptr, bias, y.ptr, N, C, H, W, self.num_filters,
                          self.f_rows, self.f_cols, self.stride, self.padding, ws, wsn, runtime.stream())
        self._x = X
        return y

    def backward(self, upstream_dx):
        dY = asarray(upstream_dx)
        N, C, H, W = self.input_shape
        w = self._param("weights")
        ws, wsn = self._scratch(N, C, H, W)
        st = runtime.stream()
        dbias = self._grad("bias").ptr if self.with_bias else None
        api.dk_conv2d_wgrad(dY.ptr, self._x.ptr, w.ptr, self._grad("weights").ptr, dbias, self._l2_strength(),
                            N, C, H, W, self.num_filters, self.f_rows, self.f_cols, self.stride, self.padding,
                            ws, wsn, st)
        if not self.needs_input_grad:
            return None
        dx = self._buf("dx", self.input_shape)

        def dgrad():
            ws2, wsn2 = self._scratch(N, C, H, W)
            api.dk_conv2d_dgrad(dY.ptr, w.ptr, dx.ptr, N, C, H, W, self.num_filters, self.f_rows, self.f_cols,
                                self.stride, self.padding, ws2, wsn2, runtime.stream())
        if self.needs_input_grad == "lazy":
            return LazyDeviceArray(dx, dgrad)
        dgrad()
        return dx

    def im2col_materialise(self, X):
        """Debug view of the reference's patch matrix (bit-exact index map of im2col_cy)."""
        self._ensure_gpu()
        X = asarray(X)
        N, C, H, W, OH, OW = self._geom(X.shape)
        P = self._buf("patches", (N * OH * OW, C * self.f_rows * self.f_cols))
        api.dk_im2col_materialise(X.ptr, P.ptr, N, C, H, W, self.f_rows, self.f_cols, self.stride, self.padding,
                                  runtime.stream())
        return P
