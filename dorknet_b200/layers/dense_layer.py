"""DenseLayer (reference: layers/dense_layer.py:6-116)."""
import numpy as np

from .layer import Layer, api, runtime, asarray


class DenseLayer(Layer):
    _h5_attrs = ("incoming_chans", "output_dim", "with_bias")
    _h5_params = ("weights", "bias")  # layers/dense_layer.py:69-117


    def __init__(self, layer_name, incoming_chans=None, output_dim=None, with_bias=True,
                 weight_regulariser=None, weight_initialiser="normal"):
        super().__init__(layer_name)
        self.incoming_chans = incoming_chans
        self.output_dim = output_dim
        self.with_bias = with_bias
        self.weight_regulariser = weight_regulariser
        self.downstream_X = None
        self.weight_initialiser = weight_initialiser
        if incoming_chans is not None and output_dim is not None:
            if self.weight_initialiser == "glorot_uniform":
                limit = np.sqrt(6.0 / (self.incoming_chans + self.output_dim))
                weights = np.random.uniform(low=-limit, high=limit,
                                            size=(self.incoming_chans, self.output_dim)).astype(np.float32)
            elif self.weight_initialiser == "normal":
                weights = 0.01 * np.random.randn(self.incoming_chans, self.output_dim).astype(np.float32)
            else:
                raise ValueError("unknown weight_initialiser {!r}".format(weight_initialiser))
            self.learned_params = {"weights": weights}
            self.grads = {"weights": np.zeros_like(weights)}
            if with_bias:
                bias = np.zeros(output_dim).astype(np.float32)
                self.learned_params.update({"bias": bias})
                self.grads.update({"bias": np.zeros_like(bias, dtype=np.float32)})
        else:
            self.learned_params = {}
            self.grads = {}

    def __repr__(self):
        return "DenseLayer({}, incoming_chans={}, output_dim={}, weight_regulariser={})".format(
            self.layer_name, self.incoming_chans, self.output_dim, repr(self.weight_regulariser))

    def forward(self, X, test_mode=False):
        """X @ W (+ b), W is [in, out] (dense_layer.py:46-55)"""
        self._ensure_gpu()
        X = asarray(X)
        B, D = X.shape
        if D != self.incoming_chans:
            raise ValueError("DenseLayer {}: input has {} features, weights expect {}".format(
                self.layer_name, D, self.incoming_chans))
        if not test_mode:
            self.downstream_X = X
        y = self._buf("y", (B, self.output_dim))
        bias = self._param("bias").ptr if self.with_bias else None
        ws, wsn = runtime.scratch(api.dk_dense_ws_bytes(B, D, self.output_dim))
        api.dk_dense_fwd(X.ptr, self._param("weights").ptr, bias, y.ptr, B, D, self.output_dim, ws, wsn,
                         runtime.stream())
        return y

    def backward(self, upstream_dx):
        """dense_layer.py:57-67"""
        dY = asarray(upstream_dx)
        B, D = self.downstream_X.shape
        dx = self._buf("dx", (B, D))
        dbias = self._grad("bias").ptr if self.with_bias else None
        ws, wsn = runtime.scratch(api.dk_dense_ws_bytes(B, D, self.output_dim))
        api.dk_dense_bwd(dY.ptr, self.downstream_X.ptr, self._param("weights").ptr, dx.ptr,
                         self._grad("weights").ptr, dbias, self._l2_strength(), B, D, self.output_dim, ws, wsn,
                         runtime.stream())
        return dx
