"""float64 evaluation of a Dorknet network (TEST INFRASTRUCTURE: the "fp64 restatement" SURVEY A.12 asks tolerances to
be stated against).  Walks the layer objects of ANY class namespace (the reference's own classes from oracle/_ref, or
the product's) by class name and public attributes, reads their fp32 parameters, and computes forward + backward in
numpy float64 with the semantics of SURVEY Appendix A -- every function cites the reference lines it restates.

Used by tests/golden/make_golden_r18.py to measure how far the REFERENCE'S OWN fp32 gradients are from the exact
ones (they are not close: sequential fp32 sums + cancellation, see tests/net_parity.py), so that the GPU path can be
held to "at least as close to the exact answer as the reference is".
"""
import numpy as np

F8 = np.float64

# Arithmetic model of the tensor-core path: tcgen05 kind::tf32 reads fp32 operands and ignores the 13 low mantissa
# bits (round toward zero); products are exact, accumulation here is float64.  Off by default (exact evaluation).
TF32_OPERANDS = False


def _t(a):
    if not TF32_OPERANDS:
        return a
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32).astype(F8)


def _pad(X, p):
    return np.pad(X, ((0, 0), (0, 0), (p, p), (p, p))) if p else X


def _out(H, k, s, p):
    return int((H + 2 * p - k) / s + 1)  # layers/im2col.pyx:18-21 (float patch count, floored)


class Node:
    """one layer's cached forward state"""


def conv_fwd(L, X):  # layers/convolution.py:58-87, im2col.pyx:16-36
    W = np.asarray(L.learned_params["weights"], F8)
    F, C, kh, kw = W.shape
    s, p = int(L.stride), int(L.padding)
    Xp = _pad(X, p)
    N, _, Hp, Wp = Xp.shape
    OH, OW = _out(X.shape[2], kh, s, p), _out(X.shape[3], kw, s, p)
    Y = np.zeros((N, F, OH, OW), F8)
    for i in range(kh):
        for j in range(kw):
            patch = Xp[:, :, i:i + s * (OH - 1) + 1:s, j:j + s * (OW - 1) + 1:s]
            Y += np.einsum("nchw,fc->nfhw", _t(patch), _t(W[:, :, i, j]), optimize=True)
    if "bias" in L.learned_params:
        Y += np.asarray(L.learned_params["bias"], F8)[None, :, None, None]
    return Y, (Xp, W, s, p, OH, OW, X.shape)


def conv_bwd(L, dY, cache, grads):  # convolution.py:90-126, im2col.pyx:209-234
    Xp, W, s, p, OH, OW, xshape = cache
    F, C, kh, kw = W.shape
    dW = np.zeros_like(W)
    dXp = np.zeros_like(Xp)
    for i in range(kh):
        for j in range(kw):
            sl = (slice(None), slice(None), slice(i, i + s * (OH - 1) + 1, s), slice(j, j + s * (OW - 1) + 1, s))
            dW[:, :, i, j] = np.einsum("nfhw,nchw->fc", _t(dY), _t(Xp[sl]), optimize=True)
            dXp[sl] += np.einsum("nfhw,fc->nchw", _t(dY), _t(W[:, :, i, j]), optimize=True)
    if L.weight_regulariser is not None:
        dW += float(L.weight_regulariser.strength) * W
    grads[(L.layer_name, "weights")] = dW
    if "bias" in L.learned_params:
        grads[(L.layer_name, "bias")] = dY.sum((0, 2, 3))
    H, Wd = xshape[2], xshape[3]
    return dXp[:, :, p:p + H, p:p + Wd]


def pw_fwd(L, X):  # layers/pointwise_convolution.py:46-55
    W = np.asarray(L.learned_params["weights"], F8)
    s = int(L.stride)
    Xs = X[:, :, ::s, ::s]
    Y = np.einsum("nchw,fc->nfhw", _t(Xs), _t(W), optimize=True)
    if "bias" in L.learned_params:
        Y += np.asarray(L.learned_params["bias"], F8)[None, :, None, None]
    return Y, (Xs, W, s)


def pw_bwd(L, dY, cache, grads):  # pointwise_convolution.py:57-75 (zero-stuffed dX of shape OH*s)
    Xs, W, s = cache
    dW = np.einsum("nfhw,nchw->fc", _t(dY), _t(Xs), optimize=True)
    if L.weight_regulariser is not None:
        dW += float(L.weight_regulariser.strength) * W
    grads[(L.layer_name, "weights")] = dW
    if "bias" in L.learned_params:
        grads[(L.layer_name, "bias")] = dY.sum((0, 2, 3))
    dXs = np.einsum("nfhw,fc->nchw", _t(dY), _t(W), optimize=True)
    if s == 1:
        return dXs
    N, C, OH, OW = dXs.shape
    dX = np.zeros((N, C, OH * s, OW * s), F8)
    dX[:, :, ::s, ::s] = dXs
    return dX


def dw_fwd(L, X):  # layers/depthwise_convolution.py:72-83, im2col.pyx:109-139
    W = np.asarray(L.learned_params["weights"], F8)
    C, kh, kw = W.shape
    s, p = int(L.stride), int(L.padding)
    Xp = _pad(X, p)
    OH, OW = _out(X.shape[2], kh, s, p), _out(X.shape[3], kw, s, p)
    Y = np.zeros((X.shape[0], C, OH, OW), F8)
    for i in range(kh):
        for j in range(kw):
            Y += Xp[:, :, i:i + s * (OH - 1) + 1:s, j:j + s * (OW - 1) + 1:s] * W[None, :, i, j, None, None]
    if "bias" in L.learned_params:
        Y += np.asarray(L.learned_params["bias"], F8)[None, :, None, None]
    return Y, (Xp, W, s, p, OH, OW, X.shape)


def dw_bwd(L, dY, cache, grads):  # depthwise_convolution.py:186-196, im2col.pyx:143-178
    Xp, W, s, p, OH, OW, xshape = cache
    C, kh, kw = W.shape
    dW = np.zeros_like(W)
    dXp = np.zeros_like(Xp)
    for i in range(kh):
        for j in range(kw):
            sl = (slice(None), slice(None), slice(i, i + s * (OH - 1) + 1, s), slice(j, j + s * (OW - 1) + 1, s))
            dW[:, i, j] = (dY * Xp[sl]).sum((0, 2, 3))
            dXp[sl] += dY * W[None, :, i, j, None, None]
    if L.weight_regulariser is not None:
        dW += float(L.weight_regulariser.strength) * W
    grads[(L.layer_name, "weights")] = dW
    if "bias" in L.learned_params:
        grads[(L.layer_name, "bias")] = dY.sum((0, 2, 3))
    return dXp[:, :, p:p + xshape[2], p:p + xshape[3]]


def bn_fwd(L, X):  # layers/batch_norm.py:54-100 (training mode; biased variance, eps inside the sqrt)
    ax = (0, 2, 3) if X.ndim == 4 else (0,)
    shp = (1, -1, 1, 1) if X.ndim == 4 else (1, -1)
    g = np.asarray(L.learned_params["gamma"], F8).reshape(shp)
    b = np.asarray(L.learned_params["beta"], F8).reshape(shp)
    mu = X.mean(ax).reshape(shp)
    var = X.var(ax).reshape(shp)
    std = np.sqrt(var + 1e-5)
    xhat = (X - mu) / std
    return g * xhat + b, (xhat, std, g, ax, shp, mu)


def bn_bwd(L, dY, cache, grads):  # batch_norm.py:118-174
    xhat, std, g, ax, shp, _ = cache
    dg = (dY * xhat).sum(ax)
    db = dY.sum(ax)
    pshape = np.asarray(L.learned_params["gamma"]).shape
    grads[(L.layer_name, "gamma")] = dg.reshape(pshape)
    grads[(L.layer_name, "beta")] = db.reshape(pshape)
    m = dY.size / dg.size
    return (g / std) * (dY - db.reshape(shp) / m - xhat * dg.reshape(shp) / m)


class Evaluator:
    def __init__(self, net):
        self.net = net
        self.cache = {}
        self.grads = {}
        self.stats = {}

    # -- forward ----------------------------------------------------------------------------------------
    def _fwd(self, L, X):
        t = type(L).__name__
        if t == "ConvLayer":
            Y, c = conv_fwd(L, X)
        elif t == "PointwiseConvLayer":
            Y, c = pw_fwd(L, X)
        elif t == "DepthwiseConvLayer":
            Y, c = dw_fwd(L, X)
        elif t == "BatchNormLayer":
            Y, c = bn_fwd(L, X)
            self.stats[L.layer_name] = (c[5].reshape(-1), c[1].reshape(-1))
        elif t == "ReLu":  # layers/activations.py:14-29
            Y, c = np.maximum(X, 0.0), (X > 0)
        elif t == "GlobalAveragePoolingLayer":  # layers/pooling.py:23-36
            Y, c = X.mean((2, 3)), X.shape
        elif t == "DenseLayer":  # layers/dense_layer.py:46-51
            W = np.asarray(L.learned_params["weights"], F8)
            Y = _t(X) @ _t(W)
            if "bias" in L.learned_params:
                Y = Y + np.asarray(L.learned_params["bias"], F8)
            c = (X, W)
        elif t == "ResidualBlock":  # layers/residual_block.py:65-75
            Xt = X
            for l in L.layer_list:
                Xt = self._fwd(l, Xt)
            sk = self._fwd(L.skip_projection, X) if L.skip_projection is not None else X
            Y = self._fwd(L.post_skip_activation, Xt + sk)
            c = None
        else:
            raise NotImplementedError(t)
        self.cache[id(L)] = c
        return Y

    def _reg(self, L):
        """layers/layer.py:42-46, residual_block.py:78-84 (skip projections are NOT in the loss)"""
        if type(L).__name__ == "ResidualBlock":
            return sum(self._reg(l) for l in L.layer_list)
        r = getattr(L, "weight_regulariser", None)
        if r is not None and L.learned_params and "weights" in L.learned_params:
            return 0.5 * float(r.strength) * float(np.sum(np.asarray(L.learned_params["weights"], F8) ** 2))
        return 0.0

    def forward(self, X, Y1h):
        X = np.asarray(X, F8)
        reg = 0.0
        for L in self.net.layers:
            X = self._fwd(L, X)
            reg += self._reg(L)
        e = np.exp(X)  # layers/losses.py:13-27 (no max subtraction)
        p = e / e.sum(1, keepdims=True)
        self.p, self.y = p, np.asarray(Y1h, F8)
        loss = float(np.mean(-np.log((p * self.y).sum(1)))) + reg
        return loss, p

    # -- backward ---------------------------------------------------------------------------------------
    def _bwd(self, L, dY):
        t = type(L).__name__
        c = self.cache[id(L)]
        if t == "ConvLayer":
            return conv_bwd(L, dY, c, self.grads)
        if t == "PointwiseConvLayer":
            return pw_bwd(L, dY, c, self.grads)
        if t == "DepthwiseConvLayer":
            return dw_bwd(L, dY, c, self.grads)
        if t == "BatchNormLayer":
            return bn_bwd(L, dY, c, self.grads)
        if t == "ReLu":
            return dY * c
        if t == "GlobalAveragePoolingLayer":
            N, C, H, W = c
            return np.broadcast_to(dY[:, :, None, None] / (H * W), c).copy()
        if t == "DenseLayer":  # dense_layer.py:54-67
            X, W = c
            dW = _t(X).T @ _t(dY)
            if L.weight_regulariser is not None:
                dW += float(L.weight_regulariser.strength) * W
            self.grads[(L.layer_name, "weights")] = dW
            if "bias" in L.learned_params:
                self.grads[(L.layer_name, "bias")] = dY.sum(0)
            return _t(dY) @ _t(W).T
        if t == "ResidualBlock":  # residual_block.py:86-97
            joined = self._bwd(L.post_skip_activation, dY)
            dx = joined
            for l in L.layer_list[::-1]:
                dx = self._bwd(l, dx)
            sk = self._bwd(L.skip_projection, joined) if L.skip_projection is not None else joined
            return dx + sk
        raise NotImplementedError(t)

    def backward(self):
        d = (self.p - self.y) / self.p.shape[0]  # losses.py:29-34
        for L in self.net.layers[::-1]:
            d = self._bwd(L, d)
        return self.grads
