"""Layer base class: the API contract of the reference (layers/layer.py:3-46) on device buffers."""
import numpy as np

from .. import runtime
from .._lib import api
from ..array import DeviceArray, DeviceScalar, asarray, empty, zeros


class Layer:
    """Same attributes and methods as the reference's Layer (layers/layer.py:5-46).

    Differences that follow from having no CPU path: parameters are created on the host with the
    reference's initialisers (so seeded construction matches), moved to HBM by `to_gpu()` -- which
    `forward` calls implicitly the first time -- and every output / gradient / cache buffer is
    allocated once per input shape and reused on later steps (outputs are overwritten by the next
    `forward` of the same layer).  `backward` writes `grads[k]` in place.
    """

    def __init__(self, layer_name, *args, **kwargs):
        self.layer_name = layer_name
        self.is_on_gpu = False
        self.learned_params = None
        self.non_learned_params = None
        self.grads = None
        self.weight_regulariser = None
        self._bufs = {}
        self._ws = None

    def __repr__(self):
        return "Layer of type {} didn't implement __repr__".format(self.__class__.__name__)

    # -- device placement (layers/layer.py:18-34) --------------------------------------------
    def to_gpu(self):
        if self.is_on_gpu:
            print("Layer {} is already on GPU, ignoring request".format(self.layer_name))
            return
        runtime.ensure_init()
        if self.learned_params is not None:
            for k, v in self.learned_params.items():
                self.learned_params[k] = asarray(v)
        if self.non_learned_params is not None:
            for k, v in self.non_learned_params.items():
                if v is not None:
                    self.non_learned_params[k] = asarray(v)
        if self.grads is not None:
            for k, v in self.grads.items():
                self.grads[k] = asarray(v)
        self.is_on_gpu = True

    def _ensure_gpu(self):
        if not self.is_on_gpu:
            self.to_gpu()

    def _param(self, key):
        """learned_params[key] as a DeviceArray (re-uploads if the user assigned a NumPy array)."""
        v = self.learned_params[key]
        if not isinstance(v, DeviceArray):
            v = asarray(v)
            self.learned_params[key] = v
        return v

    def _grad(self, key):
        v = self.grads.get(key)
        p = self._param(key)
        if not isinstance(v, DeviceArray) or v.shape != p.shape:
            v = zeros(p.shape)
            self.grads[key] = v
        return v

    def _buf(self, key, shape, dtype=np.float32):
        """Persistent buffer keyed by name; reallocated only when the shape changes."""
        shape = tuple(int(s) for s in shape)
        b = self._bufs.get(key)
        if b is None or b.shape != shape or b.dtype != np.dtype(dtype):
            b = empty(shape, dtype)
            self._bufs[key] = b
        return b

    def _zeroed_ws(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = runtime.zeroed_workspace(nbytes)
        return self._ws.data_ptr(), self._ws.numel()

    def _l2_strength(self):
        r = self.weight_regulariser
        return float(r.strength) if r is not None else 0.0

    def forward(self, X, *args, test_mode=False, **kwargs):
        pass

    def backward(self, upstream_dx, *args, **kwargs):
        pass

    def regulariser_forward(self):
        """layers/layer.py:42-46"""
        out = 0
        if self.weight_regulariser:
            out += self.weight_regulariser.forward(self._param("weights"))
        return out

    # -- checkpoints: the reference's HDF5 layout (dorknet_b200/checkpoint.py) ------------------
    _h5_attrs = ()  # constructor arguments stored as attrs of <name>/layer_info
    _h5_optional_attrs = {}  # attr -> default when an old file lacks it
    _h5_params = ()  # learned_params keys (+ grads/<key>)
    _h5_state = ()  # non_learned_params keys

    def save_to_h5(self, open_f, save_grads=True):
        from ..checkpoint import save_layer
        save_layer(self, open_f, save_grads=save_grads)

    def load_from_h5(self, open_f, load_grads=True):
        from ..checkpoint import load_layer
        load_layer(self, open_f, load_grads=load_grads)


__all__ = ["Layer", "api", "runtime", "DeviceArray", "DeviceScalar", "asarray", "empty", "zeros", "np"]
