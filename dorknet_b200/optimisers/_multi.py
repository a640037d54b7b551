"""Shared machinery of the fused multi-tensor optimisers.

The reference updates each tensor with a handful of NumPy/CuPy elementwise ops in a Python loop
(optimisers/SGDMomentum.py:31-39: ~134 tensors x 3-4 kernels for ResNet-18-depsep).  Here the whole
update set is ONE launch: a device table of (param, grad, state, n) descriptors is built once and the
kernel grid covers (chunk, tensor).
"""
import ctypes

import numpy as np

from .. import runtime
from .._lib import OptTensor
from ..array import DeviceArray, asarray, zeros


def collect_layers(network, descend, include_skip=False):
    """The reference's traversal, quirks included.

    descend=True  -> SGDMomentum.py:7-14: top-level layers with params + every layer in a composite's
                     layer_list with params (skip_projection is never visited).
    descend=False -> SGD.py:6-11 / RMSProp.py:9-15: the inner test re-checks the *block*, whose
                     learned_params is None, so nothing inside a ResidualBlock is ever updated.
    include_skip  -> opt-in fix (not reference behaviour): also update skip projections."""
    out = []
    for layer in network.layers:
        if layer.learned_params is not None:
            out.append(layer)
        if hasattr(layer, "layer_list"):
            for l in layer.layer_list:
                if descend and l.learned_params is not None:
                    out.append(l)
            if include_skip and descend and getattr(layer, "skip_projection", None) is not None:
                out.append(layer.skip_projection)
    return out


class MultiTensorOptimiser:
    needs_state = True

    def __init__(self, network, learning_rate):
        self.network = network
        self.learning_rate = learning_rate
        self.grad_scale = 1.0  # set to 1/world_size by the data-parallel wrapper
        self._table = None
        self._sig = None
        self._hyper = None       # device {lr, momentum/decay, grad_scale}
        self._hyper_host = None  # pinned staging
        self._hyper_vals = None
        self._p2p = None         # data_parallel.P2PExchange: gradients are summed over the peers inside the update kernel

    def attach_p2p(self, exchange):
        self._p2p = exchange

    def _update(self, kind, plain):
        """One launch: `plain()` = the single-GPU kernel, or the peer-memory variant when a P2PExchange is attached."""
        from .._lib import api
        tab, n, max_n = self._args()
        if not n:
            return
        if self._p2p is not None:
            api.dk_opt_multi_p2p(kind, tab, n, max_n, self.push_hyper(), self._p2p.ctx_ptr, runtime.stream())
            # nobody may overwrite gradients a peer is still reading over NVLink: the "all peers have read mine" wait
            # belongs to the exchange, not to the caller's loop -- enqueued here, right behind the kernel, it orders
            # every later backward of this stream whatever loop drives the step
            self._p2p.wait_done()
        else:
            plain(tab, n, max_n)

    # -- state checkpoint (the reference keeps this state in memory only) -------------------
    def hyper_parameters(self):
        out = {"learning_rate": float(self.learning_rate)}
        for k in ("momentum", "decay_rate"):
            if hasattr(self, k):
                out[k] = float(getattr(self, k))
        return out

    def state_dict(self):
        """{(layer_name, param): host ndarray} of the per-tensor state (velocities / running squared gradients);
        empty for plain SGD.  Tensors that were never updated yet are reported as zeros, their state's value."""
        out = {}
        if not self.needs_state:
            return out
        for layer in self.learnable_layers:
            for k in layer.learned_params.keys():
                st = self.grad_cache.get(layer, {}).get(k)
                if isinstance(st, DeviceArray):
                    st = st.get()
                elif st is None:
                    st = np.zeros(np.shape(layer.learned_params[k]), np.float32)
                out[(layer.layer_name, k)] = np.asarray(st, np.float32)
        return out

    def load_state_dict(self, state):
        """Inverse of state_dict(); every tensor of the update set must be present with its shape."""
        if not self.needs_state:
            return
        for layer in self.learnable_layers:
            for k in layer.learned_params.keys():
                key = (layer.layer_name, k)
                if key not in state:
                    raise KeyError("optimiser state has no entry for {}/{}".format(*key))
                v = np.ascontiguousarray(state[key], np.float32)
                if v.shape != tuple(np.shape(layer.learned_params[k])):
                    raise ValueError("optimiser state {}/{}: shape {} != parameter shape {}".format(
                        key[0], key[1], v.shape, tuple(np.shape(layer.learned_params[k]))))
                cur = self.grad_cache.setdefault(layer, {}).get(k)
                if isinstance(cur, DeviceArray) and cur.shape == v.shape:
                    cur.set(v)  # in place: the device table (and a captured CUDA graph) keeps pointing at it
                else:
                    self.grad_cache[layer][k] = v  # uploaded by _build_table
                    self._sig = None

    def set_learning_rate(self, new_lr):
        self.learning_rate = new_lr

    def multiply_learning_rate(self, multiplier):
        self.learning_rate *= multiplier

    def _build_table(self):
        """(Re)build the device descriptor table when any param/grad buffer changed identity."""
        import torch
        entries = []
        for layer in self.learnable_layers:
            layer._ensure_gpu()
            for k in layer.learned_params.keys():
                entries.append((layer, k, layer._param(k), layer._grad(k)))
        sig = tuple((p.ptr, g.ptr, p.size) for _, _, p, g in entries)
        if sig == self._sig:
            return
        n = len(entries)
        arr = (OptTensor * max(n, 1))()
        self._max_n = 0
        for i, (layer, k, p, g) in enumerate(entries):
            st = None
            if self.needs_state:
                st = self.grad_cache.setdefault(layer, {}).get(k)
                if not isinstance(st, DeviceArray) or st.shape != p.shape:
                    st = asarray(st) if st is not None and np.shape(st) == p.shape else zeros(p.shape)
                    self.grad_cache[layer][k] = st
            arr[i].param, arr[i].grad = p.ptr, g.ptr
            arr[i].state = st.ptr if st is not None else None
            arr[i].n = p.size
            self._max_n = max(self._max_n, p.size)
        raw = np.frombuffer(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)), dtype=np.uint8).copy()
        self._table = torch.from_numpy(raw).to(runtime.device())
        self._n = n
        self._sig = sig

    def _args(self):
        self._build_table()
        return self._table.data_ptr(), self._n, self._max_n

    def _second_hyper(self):
        return 0.0

    def push_hyper(self):
        """Mirror (lr, momentum/decay, grad_scale) into device memory; returns the device pointer.  The kernel
        reads them from there, so a CUDA-graph replay of update_weights() follows set_learning_rate()."""
        import torch
        vals = (float(self.learning_rate), float(self._second_hyper()), float(self.grad_scale))
        if self._hyper is None:
            self._hyper = torch.zeros(4, dtype=torch.float32, device=runtime.device())
            self._hyper_host = []
        if vals != self._hyper_vals:
            # a FRESH pinned staging buffer per change: an earlier non_blocking copy may still be queued, and rewriting
            # the buffer it reads from would let a learning-rate change land one step early.  Changes are rare (per
            # epoch); the last few buffers are kept alive until their copies have certainly drained.
            host = torch.tensor([vals[0], vals[1], vals[2], 0.0], dtype=torch.float32).pin_memory()
            self._hyper_host.append(host)
            del self._hyper_host[:-8]
            self._hyper.copy_(host, non_blocking=True)
            self._hyper_vals = vals
        return self._hyper.data_ptr()
