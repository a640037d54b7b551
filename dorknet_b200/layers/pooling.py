"""GlobalAveragePoolingLayer and MaxPoolLayer (reference: layers/pooling.py, layers/pooling_cy.pyx)."""
import numpy as np

from .layer import Layer, api, runtime, asarray


class GlobalAveragePoolingLayer(Layer):
    """Mean over the spatial dimensions (pooling.py:10-43)."""

    def __init__(self, layer_name):
        super().__init__(layer_name)

    def __repr__(self):
        return "GlobalAveragePoolingLayer({})".format(self.layer_name)

    def forward(self, X, test_mode=False):
        self._ensure_gpu()
        X = asarray(X)
        N, C, H, W = X.shape
        self.spatial_shape = (H, W)
        y = self._buf("y", (N, C))
        api.dk_gap_fwd(X.ptr, y.ptr, N, C, H * W, runtime.stream())
        return y

    def backward(self, upstream_dx):
        dY = asarray(upstream_dx)
        N, C = dY.shape
        H, W = self.spatial_shape
        dx = self._buf("dx", (N, C, H, W))
        api.dk_gap_bwd(dY.ptr, dx.ptr, N, C, H * W, runtime.stream())
        return dx


class MaxPoolLayer(Layer):
    """Square, non-overlapping max pooling (pooling.py:45-77).  Unlike the reference (whose
    __init__ forgets super().__init__, so the layer cannot sit in a FeedForwardNetwork) this one
    is a full Layer; the arithmetic -- first maximum in the row-major window scan wins, int32
    one-hot `max_locations` in input geometry -- is bit-identical to pooling_cy.pyx:36-88."""

    def __init__(self, layer_name, input_shape=None, stride=2):
        super().__init__(layer_name)
        self.stride = stride
        self.max_locations = None

    def __repr__(self):
        return "MaxPoolLayer(stride={})".format(self.stride)

    def forward(self, X, test_mode=False):
        self._ensure_gpu()
        X = asarray(X)
        N, C, H, W = X.shape
        s = int(self.stride)
        if H % s or W % s:
            raise ValueError("MaxPoolLayer {}: H={} and W={} must be divisible by stride {}".format(
                self.layer_name, H, W, s))
        y = self._buf("y", (N, C, H // s, W // s))
        if test_mode:
            api.dk_maxpool_fwd(X.ptr, y.ptr, N, C, H, W, s, runtime.stream())
        else:
            self.max_locations = self._buf("mask", X.shape, np.int32)
            api.dk_maxpool_fwd_train(X.ptr, y.ptr, self.max_locations.ptr, N, C, H, W, s, runtime.stream())
        return y

    def backward(self, upstream_dx):
        dY = asarray(upstream_dx)
        N, C, H, W = self.max_locations.shape
        dx = self._buf("dx", (N, C, H, W))
        api.dk_maxpool_bwd(self.max_locations.ptr, dY.ptr, dx.ptr, N, C, H, W, int(self.stride), runtime.stream())
        return dx
