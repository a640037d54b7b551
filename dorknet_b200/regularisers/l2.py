"""l2 weight regulariser (reference: regularisers/l2.py:4-17)."""
import numpy as np

from .. import runtime
from .._lib import api
from ..array import DeviceArray, DeviceScalar, alloc_scalar_slot


class l2:
    def __init__(self, strength=0.005):
        self.type = "l2"
        self.strength = strength
        self._slots = {}  # one device float per regularised tensor

    def __repr__(self):
        return "l2(strength={})".format(self.strength)

    def forward(self, X):
        """0.5 * strength * sum(X^2) (l2.py:12-14) as a lazily-read device scalar."""
        if not isinstance(X, DeviceArray):
            return 0.5 * self.strength * np.sum(np.power(X, 2))
        slot = self._slots.get(X.ptr)
        if slot is None:
            slot = self._slots[X.ptr] = alloc_scalar_slot()
        api.dk_sumsq(X.ptr, slot[0].ptr, 1.0, X.size, runtime.stream())
        return DeviceScalar([(slot[1], 0.5 * float(self.strength))])

    def backward(self, X):
        """strength * X (l2.py:16-17).  The layers fold this term into their wgrad kernels; this
        method exists for API parity and host arrays."""
        return self.strength * np.asarray(X)
