"""CPU oracle: NumPy/C restatement of the reference's CNN training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in dorknet_b200/ imports this module; it is the checker
the CUDA path is compared with (tests/, __graft_entry__.smoke()) and the "port" CPU baseline
bench.py can time.  The product path never routes through it.

Parity status: PINNED -- tests/test_oracle.py checks every function below against the
reference's own CPU implementation built into oracle/_ref (oracle/build_ref.py) and against
the golden vectors committed in tests/golden/ (made from the live reference by
tests/golden/make_golden.py).

Each function cites the reference file:line (relative to /root/reference) it restates.
All arrays are float32 NCHW unless `dtype=np.float64` is passed (the fp64 tie-breaker).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    """Compile oracle/dk_oracle.c -> oracle/_build/libdk_oracle.so (gcc, OpenMP)."""
    out_dir = os.path.join(HERE, "_build")
    so = os.path.join(out_dir, "libdk_oracle.so")
    src = os.path.join(HERE, "dk_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([gcc, "-O3", "-fopenmp", "-ffast-math", "-shared", "-fPIC", src, "-o", so])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ----------------------------------------------------------------------------- shapes
def out_hw(H, W, kh, kw, stride, pad):
    """layers/im2col.pyx:18-21: OH = int((Hp - kh)/s + 1) with Hp = H + 2 pad."""
    return (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1


def pad_nchw(X, pad):
    """layers/convolution.py:144-151 (pad_input): symmetric zero padding of H and W."""
    if pad == 0:
        return X
    return np.pad(X, ((0, 0), (0, 0), (pad, pad), (pad, pad)), "constant")


# ----------------------------------------------------------------------------- im2col / col2im
def im2col(Xp, kh, kw, stride):
    """layers/im2col.pyx:16-36.  Returns P[N*OH*OW, C*kh*kw] (bit-exact index map)."""
    Xp = _f32c(Xp)
    N, C, Hp, Wp = Xp.shape
    OH, OW = (Hp - kh) // stride + 1, (Wp - kw) // stride + 1
    P = np.empty((N * OH * OW, C * kh * kw), np.float32)
    _lib().dk_oracle_im2col(_p(Xp), N, C, Hp, Wp, kh, kw, stride, _p(P))
    return P


def row2im(rows, N, C, Hp, Wp, kh, kw, stride, pad):
    """layers/im2col.pyx:209-234: scatter-add rows back to the padded image, crop pad."""
    rows = _f32c(rows)
    dX = np.empty((N, C, Hp - 2 * pad, Wp - 2 * pad), np.float32)
    scratch = np.empty((N, C, Hp, Wp), np.float32)
    _lib().dk_oracle_row2im(_p(rows), N, C, Hp, Wp, kh, kw, stride, pad, _p(dX), _p(scratch))
    return dX


# ----------------------------------------------------------------------------- ConvLayer
def conv_fwd(X, W, b, stride, pad):
    """layers/convolution.py:58-87: pad, im2col, P @ Wflat.T (+b), NHWC rows -> NCHW."""
    F, C, kh, kw = W.shape
    Xp = pad_nchw(X, pad)
    N, _, Hp, Wp = Xp.shape
    OH, OW = (Hp - kh) // stride + 1, (Wp - kw) // stride + 1
    P = im2col(Xp, kh, kw, stride)
    out = P @ W.reshape(F, -1).T
    if b is not None:
        out = out + b.reshape(1, -1)
    Y = np.ascontiguousarray(out.reshape(N, OH, OW, F).transpose(0, 3, 1, 2))
    return Y, {"P": P, "in_shape": X.shape, "padded": (Hp, Wp)}


def conv_bwd(dY, W, cache, stride, pad, l2_strength=0.0, with_bias=False):
    """layers/convolution.py:90-126: db = sum(dY,(0,2,3)); dW = dY_flat.T @ P (+ s*W);
    dX = crop(col2im(dY_flat @ Wflat))."""
    F, C, kh, kw = W.shape
    N = cache["in_shape"][0]
    Hp, Wp = cache["padded"]
    g = dY.transpose(0, 2, 3, 1).reshape(-1, F)
    grads = {}
    if with_bias:
        grads["bias"] = dY.sum(axis=(0, 2, 3))
    dW = (g.T @ cache["P"]).reshape(W.shape)
    if l2_strength:
        dW = dW + l2_strength * W
    grads["weights"] = dW
    rows = g @ W.reshape(F, -1)
    dX = row2im(rows, N, C, Hp, Wp, kh, kw, stride, pad)
    return dX, grads


# ----------------------------------------------------------------------------- PointwiseConvLayer
def pointwise_fwd(X, W, b, stride):
    """layers/pointwise_convolution.py:46-55: subsample X[:,:,::s,::s]; Xnhwc @ W.T (+b)."""
    Xs = X[:, :, ::stride, ::stride] if stride > 1 else X
    N, C, H, Wd = Xs.shape
    rows = Xs.transpose(0, 2, 3, 1).reshape(-1, C)
    out = rows @ W.T
    if b is not None:
        out = out + b.reshape(1, -1)
    Y = np.ascontiguousarray(out.reshape(N, H, Wd, W.shape[0]).transpose(0, 3, 1, 2))
    return Y, {"rows": rows}


def pointwise_bwd(dY, W, cache, stride, l2_strength=0.0, with_bias=False):
    """layers/pointwise_convolution.py:57-75: dW = dY_flat.T @ rows (+ s*W); dX = dY_flat @ W;
    stride > 1 -> zero-stuffed [N, C, OH*s, OW*s] (NOT the original H, W)."""
    N, F, OH, OW = dY.shape
    C = W.shape[1]
    g = dY.transpose(0, 2, 3, 1).reshape(-1, F)
    grads = {}
    if with_bias:
        grads["bias"] = dY.sum(axis=(0, 2, 3))
    dW = (g.T @ cache["rows"]).reshape(W.shape)
    if l2_strength:
        dW = dW + l2_strength * W
    grads["weights"] = dW
    dXs = (g @ W).reshape(N, OH, OW, C).transpose(0, 3, 1, 2)
    if stride > 1:
        dX = np.zeros((N, C, OH * stride, OW * stride), dY.dtype)
        dX[:, :, ::stride, ::stride] = dXs
    else:
        dX = np.ascontiguousarray(dXs)
    return dX, grads


# ----------------------------------------------------------------------------- DepthwiseConvLayer
def depthwise_fwd(X, W, b, stride, pad):
    """layers/depthwise_convolution.py:72-83 + layers/im2col.pyx:109-139."""
    Xp = _f32c(pad_nchw(X, pad))
    W = _f32c(W)
    N, C, Hp, Wp = Xp.shape
    kh, kw = W.shape[1:]
    OH, OW = (Hp - kh) // stride + 1, (Wp - kw) // stride + 1
    Y = np.empty((N, C, OH, OW), np.float32)
    _lib().dk_oracle_depthwise_fwd(_p(Xp), _p(W), N, C, Hp, Wp, kh, kw, stride, _p(Y))
    if b is not None:
        Y += b[None, :, None, None]
    return Y, {"Xp": Xp}


def depthwise_bwd(dY, W, cache, stride, pad, l2_strength=0.0, with_bias=False):
    """layers/depthwise_convolution.py:186-196 + layers/im2col.pyx:143-178: per-image dW
    partials summed over N; dX scattered into the padded buffer then cropped."""
    Xp = cache["Xp"]
    dY = _f32c(dY)
    W = _f32c(W)
    N, C, Hp, Wp = Xp.shape
    kh, kw = W.shape[1:]
    dX = np.empty((N, C, Hp - 2 * pad, Wp - 2 * pad), np.float32)
    dWn = np.empty((N, C, kh, kw), np.float32)
    scratch = np.empty((N, C, Hp, Wp), np.float32)
    _lib().dk_oracle_depthwise_bwd(_p(dY), _p(Xp), _p(W), N, C, Hp, Wp, kh, kw, stride, pad,
                                   _p(dX), _p(dWn), _p(scratch))
    grads = {}
    if with_bias:
        grads["bias"] = dY.sum(axis=(0, 2, 3))
    dW = dWn.sum(axis=0)
    if l2_strength:
        dW = dW + l2_strength * W
    grads["weights"] = dW
    return dX, grads


# ----------------------------------------------------------------------------- BatchNormLayer
def bn_stats(X):
    """layers/batch_norm_stats_cy.pyx:17-46 (4-D) / layers/batch_norm.py:67-68 (2-D)."""
    if X.ndim == 4 and X.dtype == np.float32:
        X = _f32c(X)
        N, C, H, W = X.shape
        mean = np.empty(C, np.float32)
        var = np.empty(C, np.float32)
        _lib().dk_oracle_bn_stats(_p(X), N, C, H, W, _p(mean), _p(var))
        return mean, var
    ax = (0, 2, 3) if X.ndim == 4 else 0
    return X.mean(axis=ax), X.var(axis=ax)


def bn_fwd_train(X, gamma, beta, running_mean, running_std, momentum=0.95, eps=1e-5):
    """layers/batch_norm.py:64-100: std = sqrt(var + eps); X_hat = (X - mean)/std;
    running_mean / running_STD EMA (first batch assigns); out = gamma*X_hat + beta.
    gamma/beta/running_* have shape (1,C,1,1) for 4-D inputs, (C,) for 2-D."""
    mean, var = bn_stats(X)
    std = np.sqrt(var + X.dtype.type(eps))
    if X.ndim == 4:
        mean = mean[None, :, None, None]
        std = std[None, :, None, None]
    X_demean = X - mean
    X_hat = X_demean / std
    m = X.dtype.type(momentum)
    new_rm = mean if running_mean is None else m * running_mean + (1 - m) * mean
    new_rs = std if running_std is None else m * running_std + (1 - m) * std
    Y = gamma * X_hat + beta
    return Y, {"X_demean": X_demean, "X_hat": X_hat, "std": std, "mean": mean, "in_shape": X.shape}, new_rm, new_rs


def bn_fwd_test(X, gamma, beta, running_mean, running_std):
    """layers/batch_norm.py:112-115."""
    return gamma * ((X - running_mean) / running_std) + beta


def bn_bwd(dY, gamma, cache):
    """layers/batch_norm.py:118-174: dgamma = sum(dY*X_hat), dbeta = sum(dY),
    dX = gamma/std * (dY - mean(dY) - X_demean * sum(dY*X_demean) / (N_eff * std^2))."""
    four = dY.ndim == 4
    ax = (0, 2, 3) if four else 0
    shp = cache["in_shape"]
    n_eff = float(shp[0] * shp[2] * shp[3]) if four else float(shp[0])
    ex = (lambda v: v[None, :, None, None]) if four else (lambda v: v)
    dgamma = ex((dY * cache["X_hat"]).sum(axis=ax))
    dbeta = ex(dY.sum(axis=ax))
    std_recip = 1.0 / cache["std"]
    up_mean = ex(dY.mean(axis=ax))
    dot_sum = ex((dY * cache["X_demean"]).sum(axis=ax))
    other = (1.0 / n_eff) * (cache["X_demean"] * std_recip ** 2)
    dX = (gamma * std_recip) * (dY - up_mean - other * dot_sum)
    return dX.astype(dY.dtype), {"gamma": dgamma, "beta": dbeta}


# ----------------------------------------------------------------------------- ReLu
def relu_fwd(X, want_mask=True):
    """layers/relu_cy.pyx:11-107 / layers/activations.py:14-29."""
    X = _f32c(X)
    Y = np.empty_like(X)
    mask = np.empty_like(X) if want_mask else None
    _lib().dk_oracle_relu_fwd(_p(X), ctypes.c_size_t(X.size), _p(Y), _p(mask) if want_mask else None)
    return Y, mask


def relu_bwd(dY, mask):
    """layers/activations.py:44-47."""
    return dY * mask


# ----------------------------------------------------------------------------- pooling
def gap_fwd(X):
    """layers/pooling.py:23-27."""
    return X.mean(axis=(2, 3))


def gap_bwd(dY, H, W):
    """layers/pooling.py:29-36."""
    scaled = (1.0 / float(H * W)) * dY[:, :, None, None]
    return scaled * np.ones((dY.shape[0], dY.shape[1], H, W), dY.dtype)


def maxpool_fwd(X, stride, train=True):
    """layers/pooling_cy.pyx:10-69: returns (Y, int32 one-hot mask in input geometry | None)."""
    X = _f32c(X)
    N, C, H, W = X.shape
    Y = np.empty((N, C, H // stride, W // stride), np.float32)
    mask = np.empty((N, C, H, W), np.int32) if train else None
    _lib().dk_oracle_pool(_p(X), N, C, H, W, stride, _p(Y), _p(mask) if train else None)
    return Y, mask


def maxpool_bwd(mask, dY, stride):
    """layers/pooling_cy.pyx:72-88."""
    mask = np.ascontiguousarray(mask, dtype=np.int32)
    dY = _f32c(dY)
    N, C, H, W = mask.shape
    dX = np.empty((N, C, H, W), np.float32)
    _lib().dk_oracle_pool_bwd(_p(mask), _p(dY), N, C, H, W, stride, _p(dX))
    return dX


# ----------------------------------------------------------------------------- Dense / loss / l2
def dense_fwd(X, W, b):
    """layers/dense_layer.py:46-55: X @ W (W is [in, out]) (+ b)."""
    out = X @ W
    return out + b[None, :] if b is not None else out


def dense_bwd(dY, X, W, l2_strength=0.0, with_bias=True):
    """layers/dense_layer.py:57-67."""
    grads = {}
    if with_bias:
        grads["bias"] = dY.sum(axis=0)
    dW = X.T @ dY
    if l2_strength:
        dW = dW + l2_strength * W
    grads["weights"] = dW
    return dY @ W.T, grads


def softmax_xent_fwd(X, y_one_hot):
    """layers/losses.py:13-27: softmax WITHOUT max subtraction; loss = mean(-log(sum_j p_j y_j))."""
    e = np.exp(X)
    p = (1.0 / e.sum(axis=1)).reshape(-1, 1) * e
    loss = (1.0 / float(X.shape[0])) * np.sum(-np.log((p * y_one_hot).sum(axis=1)))
    return loss, p


def softmax_xent_bwd(p, y_one_hot):
    """layers/losses.py:29-34."""
    return (1.0 / float(p.shape[0])) * (p - y_one_hot)


def l2_fwd(W, strength):
    """regularisers/l2.py:12-14."""
    return 0.5 * strength * np.sum(np.power(W, 2))


# ----------------------------------------------------------------------------- optimisers
def sgd_update(w, g, lr):
    """optimisers/SGD.py:20-24."""
    return w + (-lr * g)


def sgdm_update(w, g, v, lr, momentum):
    """optimisers/SGDMomentum.py:31-39: v' = -lr*g + m*v ; w' = w + v'."""
    v2 = -lr * g + momentum * v
    return w + v2, v2


def rmsprop_update(w, g, c, lr, decay):
    """optimisers/RMSProp.py:28-36: c' = d*c + (1-d)*g^2 ; w' = w - lr*g/sqrt(c' + 1e-5)."""
    c2 = decay * c + (1 - decay) * np.power(g, 2)
    return w + (-lr * g / np.sqrt(c2 + 1e-5)), c2


# ----------------------------------------------------------------------------- data (next row, §8f-1)
def input_u8_nhwc(img, img_b=None, lam=0.0, sub=128.0):
    """Decoded uint8 NHWC batch -> the fp32 NCHW network input: data_loading/image_preprocessor.py:36-37
    (`im.astype(np.float32).transpose(2,0,1); im -= 128.0` per image), and with a second batch the loader's mixup
    X = lam*X_b + (1-lam)*X_a (data_loading/image_data_loader.py:100-105).  What dk_input_u8_nhwc does on the device."""
    Xa = np.ascontiguousarray(np.asarray(img).astype(np.float32).transpose(0, 3, 1, 2)) - np.float32(sub)
    if img_b is None:
        return Xa
    Xb = np.ascontiguousarray(np.asarray(img_b).astype(np.float32).transpose(0, 3, 1, 2)) - np.float32(sub)
    lam = np.float32(lam)
    return lam * Xb + (1 - lam) * Xa


def mixup(Xa, Xb, ya, yb, lam):
    """data_loading/image_data_loader.py:100-112: X = lam*X_b + (1-lam)*X_a (same for labels)."""
    return lam * Xb + (1 - lam) * Xa, lam * yb + (1 - lam) * ya
