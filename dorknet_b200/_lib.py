"""ctypes binding of libdorknet_b200.so (the C ABI declared in include/dorknet_b200.h).

The product path has NO fallback: if the shared library is missing, or there is no B200, every
call raises.  Nothing here (or anywhere in dorknet_b200/) imports oracle/.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DK_LIB_PATH") or os.path.join(HERE, "lib", "libdorknet_b200.so")

DK_OK, DK_ERR_INVALID, DK_ERR_CUDA, DK_ERR_WORKSPACE, DK_ERR_UNSUPPORTED = 0, 1, 2, 3, 4

P, I, F, L, Z = c_void_p, c_int, c_float, c_int64, c_size_t

# name -> (restype, argtypes).  Functions returning c_int are status codes and get an error check.
PROTOS = {
    "dk_version": (I, []),
    "dk_last_error": (c_char_p, []),
    "dk_init": (I, [I]),
    "dk_destroy": (I, []),
    "dk_sm_count": (I, []),
    "dk_kernel_launches": (ctypes.c_ulonglong, []),
    "dk_gemm_call_counts": (None, [P, P]),
    "dk_tc_debug_set": (I, [I, I]),
    "dk_dw_debug_set": (I, [I]),
    "dk_set_gemm_backend": (I, [I]),
    "dk_get_gemm_backend": (I, []),
    "dk_relu_fwd": (I, [P, P, P, L, P]),
    "dk_relu_bwd": (I, [P, P, P, L, P]),
    "dk_add_relu_fwd": (I, [P, P, P, L, P]),
    "dk_add": (I, [P, P, P, L, P]),
    "dk_bn_ws_bytes": (Z, [I]),
    "dk_bn_stats": (I, [P, P, P, I, I, I, P, Z, P]),
    "dk_bn_fwd_train": (I, [P, P, P, P, P, P, I, F, F, P, P, P, P, I, I, I, I, P, Z, P]),
    "dk_bn_fwd_train_add": (I, [P, P, P, P, P, P, P, I, F, F, P, P, P, P, I, I, I, I, P, Z, P]),
    "dk_bn_fwd_infer": (I, [P, P, P, P, P, P, I, I, I, I, P]),
    "dk_bn_apply": (I, [P, P, P, P, I, I, I, I, P]),
    "dk_bn_apply_strided": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "dk_bn_bwd": (I, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, P, Z, P]),
    "dk_bn_bwd_join": (I, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, P, Z, P]),
    "dk_bn_bwd_strided": (I, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, P, Z, P]),
    "dk_dwconv_ws_bytes": (Z, [I, I, I, I, I, I, I, I]),
    "dk_dwconv_fwd": (I, [P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "dk_dwconv_bwd": (I, [P, P, P, P, P, P, P, P, I, P, F, I, I, I, I, I, I, I, I, P, Z, P]),
    "dk_conv2d_ws_bytes": (Z, [I, I, I, I, I, I, I, I, I]),
    "dk_conv2d_fwd": (I, [P, P, P, P, I, I, I, I, I, I, I, I, I, P, Z, P]),
    "dk_conv2d_dgrad": (I, [P, P, P, I, I, I, I, I, I, I, I, I, P, Z, P]),
    "dk_conv2d_wgrad": (I, [P, P, P, P, P, F, I, I, I, I, I, I, I, I, I, P, Z, P]),
    "dk_im2col_materialise": (I, [P, P, I, I, I, I, I, I, I, I, P]),
    "dk_pwconv_ws_bytes": (Z, [I, I, I, I, I, I]),
    "dk_pwconv_fwd": (I, [P, P, P, P, I, I, I, I, I, I, P, Z, P]),
    "dk_pwconv_dgrad": (I, [P, P, P, I, I, I, I, I, I, P, Z, P]),
    "dk_pwconv_wgrad": (I, [P, P, P, P, P, F, I, I, I, I, I, I, P, Z, P]),
    "dk_pw_pack_bytes": (Z, [I, I, I, I, I]),
    "dk_pw_pack": (I, [P, P, I, I, I, I, I, P]),
    "dk_pwconv_fwd_packed": (I, [P, P, P, P, I, I, I, I, I, P, Z, P]),
    "dk_pwconv_dgrad_packed": (I, [P, P, P, I, I, I, I, I, I, P, Z, P]),
    "dk_pwconv_wgrad_packed": (I, [P, I, P, I, P, P, P, F, I, I, I, I, I, I, P, Z, P]),
    "dk_bn_fold_fwd": (I, [P, P, P, P, P, P, P, P, I, I, P]),
    "dk_dwconv_fwd_bn_ws_bytes": (Z, [I, I, I, I, I, I, I, I]),
    "dk_dwconv_fwd_bn": (I, [P, P, P, P, I, I, I, I, I, I, I, I, P, P, P, P, I, F, F, P, P, P, P, P, P, Z, P]),
    "dk_bn_fold_bwd": (I, [P, P, P, P, P, P, P, P, F, L, P, P, P, P, P, I, I, P]),
    "dk_pwconv_dgrad_affine": (I, [P, P, P, P, P, P, I, I, I, I, I, P, Z, P]),
    "dk_dense_ws_bytes": (Z, [I, I, I]),
    "dk_dense_fwd": (I, [P, P, P, P, I, I, I, P, Z, P]),
    "dk_dense_bwd": (I, [P, P, P, P, P, P, F, I, I, I, P, Z, P]),
    "dk_bias_grad": (I, [P, P, I, I, I, P, Z, P]),
    "dk_gap_fwd": (I, [P, P, I, I, I, P]),
    "dk_gap_bwd": (I, [P, P, I, I, I, P]),
    "dk_maxpool_fwd": (I, [P, P, I, I, I, I, I, P]),
    "dk_maxpool_fwd_train": (I, [P, P, P, I, I, I, I, I, P]),
    "dk_maxpool_bwd": (I, [P, P, P, I, I, I, I, I, P]),
    "dk_softmax_xent_fwd": (I, [P, P, P, P, I, I, P]),
    "dk_softmax_xent_bwd": (I, [P, P, P, I, I, P]),
    "dk_sumsq": (I, [P, P, F, L, P]),
    "dk_sumsq_multi": (I, [P, I, P]),
    "dk_opt_sgd_multi": (I, [P, I, L, F, F, P, P]),
    "dk_opt_sgdm_multi": (I, [P, I, L, F, F, F, P, P]),
    "dk_opt_rmsprop_multi": (I, [P, I, L, F, F, F, P, P]),
    "dk_mixup": (I, [P, P, P, F, L, P]),
    "dk_input_u8_nhwc": (I, [P, P, P, F, F, I, I, I, I, P]),
    "dk_p2p_alloc": (I, [Z, P, P]),
    "dk_p2p_open": (I, [P, P]),
    "dk_p2p_close": (I, [P]),
    "dk_p2p_free": (I, [P]),
    "dk_p2p_wait_done": (I, [P, P]),
    "dk_opt_multi_p2p": (I, [I, P, I, L, P, P, P]),
}

# value-returning (not status) functions
_NO_CHECK = {"dk_version", "dk_last_error", "dk_sm_count", "dk_kernel_launches", "dk_gemm_call_counts", "dk_get_gemm_backend", "dk_bn_ws_bytes",
             "dk_dwconv_ws_bytes", "dk_dwconv_fwd_bn_ws_bytes", "dk_pw_pack_bytes", "dk_conv2d_ws_bytes", "dk_pwconv_ws_bytes", "dk_dense_ws_bytes"}


class OptTensor(ctypes.Structure):
    """mirror of dk_opt_tensor"""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("state", c_void_p), ("n", c_int64)]


class P2PCtx(ctypes.Structure):
    """mirror of dk_p2p_ctx"""
    _fields_ = [("world", c_int), ("rank", c_int), ("grad_delta", c_int64 * 8), ("ready", c_void_p * 8),
                ("done", c_void_p * 8), ("epoch", c_void_p), ("reduced", c_void_p * 8), ("grad_base", c_void_p),
                ("nfloats", c_int64), ("slice", c_int64)]


class SumsqTask(ctypes.Structure):
    """mirror of dk_sumsq_task"""
    _fields_ = [("w", c_void_p), ("out", c_void_p), ("n", c_int64), ("scale", c_float), ("pad_", c_int)]


class DorknetError(RuntimeError):
    pass


_cdll = None
_launches = 0  # kernels-launching C-ABI calls made by this process (bench.py reports it)
_timer = None  # optional {name: callback(name, args) -> context} installed by set_call_timer()


def set_call_timer(timer):
    """Install (or clear, with None) a per-call timing hook: `timer` maps C-ABI function names to a
    callable(name, args) returning an object with .stop().  Only the named calls pay for it; bench.py
    uses it to bracket the dominant kernel family with CUDA events inside the timed region."""
    global _timer
    _timer = timer


def load():
    """Load the shared library (no GPU needed for this); raise loudly if it was never built."""
    global _cdll
    if _cdll is None:
        if not os.path.exists(LIB_PATH):
            raise DorknetError(
                "libdorknet_b200.so not found at %s: build it with `python -m dorknet_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        _cdll = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOS.items():
            fn = getattr(_cdll, name)
            fn.restype = res
            fn.argtypes = args
    return _cdll


def last_error():
    return load().dk_last_error().decode("utf-8", "replace")


def check(rc, name):
    if rc == DK_OK:
        return
    msg = "%s failed (code %d): %s" % (name, rc, last_error())
    if rc == DK_ERR_INVALID:
        raise ValueError(msg)
    raise DorknetError(msg)


def launch_count():
    """C-ABI calls that launch kernels, made by this process."""
    return _launches


def kernel_launches():
    """CUDA kernels actually launched by the library (counted at every launch site)."""
    return int(load().dk_kernel_launches())


def gemm_call_counts():
    tc, simt = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    load().dk_gemm_call_counts(ctypes.byref(tc), ctypes.byref(simt))
    return int(tc.value), int(simt.value)


class _Api:
    """Attribute access returns a checked callable: api.dk_relu_fwd(...)"""

    def __getattr__(self, name):
        if name not in PROTOS:
            raise AttributeError(name)
        fn = getattr(load(), name)
        if name in _NO_CHECK:
            wrapped = fn
        else:
            def wrapped(*a, _fn=fn, _name=name):
                global _launches
                _launches += 1
                if _timer is not None and _name in _timer:
                    tok = _timer[_name](_name, a)
                    rc = _fn(*a)
                    tok.stop()
                else:
                    rc = _fn(*a)
                if rc != DK_OK:
                    check(rc, _name)
        setattr(self, name, wrapped)
        return wrapped


api = _Api()
