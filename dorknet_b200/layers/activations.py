"""ReLu (reference: layers/activations.py:6-53, layers/relu_cy.pyx)."""
from .layer import Layer, api, runtime, asarray
from ..array import LazyBNOutput, LazyDeviceArray, LazyReluOutput


class ReLu(Layer):

    def __init__(self, layer_name):
        super().__init__(layer_name)
        self._y = None
        self._fused_bn = None

    def __repr__(self):
        return "ReLu({})".format(self.layer_name)

    def forward(self, X, test_mode=False):
        """out = max(x, 0) (activations.py:14-29).  The reference also stores a float 0/1 mask
        (`positive_locs`); here the mask is implied by the output (y > 0 <=> x > 0), so the forward
        moves 2n bytes instead of 3n and `positive_locs` is materialised only if somebody reads it."""
        self._ensure_gpu()
        X = asarray(X)
        y = self._buf("y", X.shape)
        self._fused_bn = None
        if isinstance(X, LazyBNOutput) and X.fusable and not test_mode:
            # BatchNorm -> ReLU: one pass y = relu(x*scale + shift); backward: the BatchNorm masks dY itself.  The pass
            # is launched when the consumer reads the result (a strided pointwise consumer asks for less: LazyReluOutput)
            bn = X.bn
            X.consume()
            self._fused_bn = bn
            bn.expect_fused_relu(y)
            out = LazyReluOutput(y, lambda: bn.fused_relu_apply(y), bn, self)
            self._y = out
            return out
        api.dk_relu_fwd(X.ptr, y.ptr, None, X.size, runtime.stream())
        if not test_mode:
            self._y = y
        return y

    @property
    def positive_locs(self):
        if self._y is None:
            return None
        m = self._buf("mask", self._y.shape)
        tmp = self._buf("mask_tmp", self._y.shape)
        api.dk_relu_fwd(self._y.ptr, tmp.ptr, m.ptr, self._y.size, runtime.stream())
        return m

    def backward(self, upstream_dx):
        """dY * mask (activations.py:44-47)"""
        upstream_dx = asarray(upstream_dx)
        if self._fused_bn is not None:
            if isinstance(self._y, LazyDeviceArray) and self._fused_bn._pending is not None:
                self._y.materialise()  # (nobody read the output: the deferred BatchNorm pass still has to run)
            return upstream_dx  # the fused BatchNorm's backward applies the (x_hat*gamma+beta > 0) mask
        dx = self._buf("dx", upstream_dx.shape)
        api.dk_relu_bwd(upstream_dx.ptr, self._y.ptr, dx.ptr, upstream_dx.size, runtime.stream())
        return dx
