"""Stub of `cupy` for running the reference's CPU path (TEST INFRASTRUCTURE ONLY).

The reference imports cupy at module level everywhere (/root/reference/layers/layer.py:1)
but its CPU path only calls cp.get_array_module, cp.dot on NumPy arrays
(/root/reference/layers/convolution.py:83, pointwise_convolution.py:51) and cp.asnumpy
(/root/reference/network/feed_forward_network.py:83).  Everything resolves to NumPy.
"""
import numpy as _np


def get_array_module(*args):
    return _np


def asnumpy(a):
    return _np.asarray(a)


def asarray(a, dtype=None):
    return _np.asarray(a, dtype=dtype)


def __getattr__(name):
    return getattr(_np, name)
