"""Helpers shared by the -m gpu parity tests (they call the CUDA path through the C ABI via the
Python mirror and compare with the oracle / the golden vectors)."""
import numpy as np


def err(a, b):
    """normalised max-abs error max|a-b| / max|b|"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    return float(d / max(np.max(np.abs(b)) if b.size else 1.0, 1e-30))


def assert_close(a, b, tol, what="", atol=0.0):
    """max|a-b| <= tol*max|b| + atol.  `atol` is for quantities that are zero in exact arithmetic
    (e.g. dbeta of a BatchNorm that feeds another BatchNorm), where both sides are rounding noise."""
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a64.shape == b64.shape, (a64.shape, b64.shape)
    d = float(np.max(np.abs(a64 - b64))) if a64.size else 0.0
    ref = float(np.max(np.abs(b64))) if b64.size else 1.0
    assert d <= tol * max(ref, 1e-30) + atol, "%s: max-abs error %.3e > %.1e * %.3e + %.1e" % (what, d, tol, ref, atol)


# Tolerances, stated once:
FP32 = 1e-5      # fp32 bandwidth kernels vs oracle / golden
FP32_RED = 2e-5  # long fp32 reductions (BN grads, dW of depthwise)
GEMM = 2e-3      # TF32 tensor-core GEMMs, fwd / dgrad   (fp32 SIMT backend is held to 1e-5 separately)
GEMM_W = 5e-3    # TF32 wgrad (reduction over N*OH*OW)
SIMT = 2e-5
