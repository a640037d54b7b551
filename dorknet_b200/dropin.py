"""Make the B200 layers importable under the reference's own module names.

The reference's container and example scripts do `from layers.convolution import ConvLayer`,
`from regularisers.l2 import l2`, `from optimisers.SGDMomentum import SGDMomentum`, `import cupy as
cp` (network/feed_forward_network.py:1-13, examples/*.py).  install() binds those names to this
package (and `cupy` to a three-function shim over DeviceArray), so the reference's
network/feed_forward_network.py and training loops run UNCHANGED on top of the sm_100a kernels.
"""
import importlib
import sys
import types

_LAYER_MODULES = ["layer", "convolution", "depthwise_convolution", "pointwise_convolution", "batch_norm",
                  "pooling", "activations", "dense_layer", "residual_block", "losses"]


def _cupy_shim():
    import numpy as np
    from . import array as A
    m = types.ModuleType("cupy")
    m.__doc__ = "dorknet_b200 shim: the subset of CuPy the reference's container and loops touch"
    m.asarray = A.asarray
    m.asnumpy = A.asnumpy
    m.ndarray = A.DeviceArray
    m.float32 = np.float32

    def get_array_module(*arrays):
        return m if any(isinstance(a, A.DeviceArray) for a in arrays) else np

    def argmax(a, axis=None):
        return np.argmax(A.asnumpy(a), axis=axis)

    def _sum(a, *args, **kw):
        return np.sum(A.asnumpy(a), *args, **kw)

    m.get_array_module = get_array_module
    m.argmax = argmax
    m.sum = _sum
    return m


def install(shim_cupy=True, stub_h5py=True):
    """Idempotent (stub_h5py: bind `h5py` to dorknet_b200.minih5 when the real package is missing).  Refuses to shadow a real, already-imported `layers` package."""
    pkg = importlib.import_module("dorknet_b200.layers")
    existing = sys.modules.get("layers")
    if existing is not None and existing is not pkg:
        raise RuntimeError("a different `layers` package is already imported; cannot install the drop-in")
    sys.modules["layers"] = pkg
    for name in _LAYER_MODULES:
        sys.modules["layers." + name] = importlib.import_module("dorknet_b200.layers." + name)
    for top, subs in (("regularisers", ["l2"]), ("optimisers", ["SGD", "SGDMomentum", "RMSProp"])):
        sys.modules[top] = importlib.import_module("dorknet_b200." + top)
        for sname in subs:
            sys.modules[top + "." + sname] = importlib.import_module("dorknet_b200.%s.%s" % (top, sname))
    if shim_cupy and "cupy" not in sys.modules:
        try:
            importlib.import_module("cupy")
        except ImportError:
            sys.modules["cupy"] = _cupy_shim()
    if stub_h5py and "h5py" not in sys.modules:
        try:
            importlib.import_module("h5py")
        except ImportError:
            # the reference's `import h5py` (network/feed_forward_network.py:2, every layer module) gets the pure-Python
            # HDF5 subset its checkpoints use: save_weights_to_h5 / load_network_from_json_and_h5 run unchanged
            sys.modules["h5py"] = importlib.import_module("dorknet_b200.minih5")


def accelerate(network, optimiser=None, warmup=2):
    """Full speed behind an unchanged loop: after `warmup` steps network.forward / network.backward /
    optimiser.update_weights of THESE INSTANCES replay CUDA graphs captured from the very same calls (one line added to a
    reference training script, right after the network and optimiser are built).  See dorknet_b200.graph.AutoGraph."""
    from .graph import AutoGraph
    return AutoGraph(network, optimiser, warmup=warmup)
