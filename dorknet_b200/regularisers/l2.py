"""l2 weight regulariser (reference: regularisers/l2.py:4-17)."""
import ctypes

import numpy as np

from .. import runtime
from .._lib import SumsqTask, api
from ..array import DeviceArray, DeviceScalar, alloc_scalar_slot

# l2.forward() calls of the current step that have not been launched yet: (w DeviceArray, slot DeviceArray)
_pending = []
_tables = {}  # tuple of (w.ptr, slot.ptr, n) -> device task table


def flush_pending():
    """Launch ONE kernel computing sum(w^2) for every l2.forward() issued since the last flush.  The loss layer
    calls this (its forward runs after every layer's regulariser_forward, feed_forward_network.py:52-61);
    DeviceScalar reads call it too, so a term is never read before it was computed."""
    global _pending
    if not _pending:
        return
    import torch
    tasks, _pending = _pending, []
    key = tuple((w.ptr, slot.ptr, w.size) for w, slot in tasks)
    tab = _tables.get(key)
    if tab is None:
        arr = (SumsqTask * len(tasks))()
        for i, (w, slot) in enumerate(tasks):
            arr[i].w, arr[i].out, arr[i].n, arr[i].scale = w.ptr, slot.ptr, w.size, 1.0
        raw = np.frombuffer(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)), dtype=np.uint8).copy()
        tab = torch.from_numpy(raw).to(runtime.device())
        if len(_tables) > 64:
            _tables.clear()
        _tables[key] = tab
    api.dk_sumsq_multi(tab.data_ptr(), len(tasks), runtime.stream())


class l2:
    def __init__(self, strength=0.005):
        self.type = "l2"
        self.strength = strength
        self._slots = {}  # one device float per regularised tensor

    def __repr__(self):
        return "l2(strength={})".format(self.strength)

    def forward(self, X):
        """0.5 * strength * sum(X^2) (l2.py:12-14) as a lazily-read device scalar; the reduction itself is
        batched with the other layers' terms (flush_pending)."""
        if not isinstance(X, DeviceArray):
            return 0.5 * self.strength * np.sum(np.power(X, 2))
        slot = self._slots.get(X.ptr)
        if slot is None:
            slot = self._slots[X.ptr] = alloc_scalar_slot()
        _pending.append((X, slot[0]))
        return DeviceScalar([(slot[1], 0.5 * float(self.strength))])

    def backward(self, X):
        """strength * X (l2.py:16-17).  The layers fold this term into their wgrad kernels; this
        method exists for API parity and host arrays."""
        return self.strength * np.asarray(X)
