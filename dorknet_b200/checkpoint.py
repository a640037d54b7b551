"""HDF5 checkpoints in the reference's on-disk layout.

The reference writes one group per layer (network/feed_forward_network.py:90-95, each layer's save_to_h5, e.g.
layers/convolution.py:226-258, layers/batch_norm.py:176-211, layers/residual_block.py:99-114):

    <layer>/layer_info            empty float32 dataset; attrs: "type" + the constructor arguments
    <layer>/weights | bias        learned parameters (attrs weight_regulariser_type / _strength as byte strings)
    <layer>/gamma | beta | running_mean | running_std         (BatchNormLayer)
    <layer>/grads/<param>         gradients (save_grads=True)

and rebuilds a network from the JSON structure file + the "type" attrs (feed_forward_network.py:106-139).  Here the
layout is data: every layer class lists its attrs / parameter names (`_h5_attrs`, `_h5_params`, `_h5_state`) and the
two functions below do the work, copying device buffers to the host with `.get()` (a deferred BatchNorm is flushed by
reading its running statistics).  `h5` is h5py when it is installed, dorknet_b200.minih5 otherwise -- both write
files the other (and the reference) reads.

Optimiser state (the reference keeps `velocities` / `running_sq_grads` only in memory, so a
resumed run restarts its momentum): `save_optimiser_state` / `load_optimiser_state` put it under
`__optimiser__/<layer>/<param>`, a group name no layer can have in the JSON structure file, so reference readers
ignore it.
"""
import numpy as np

from .array import DeviceArray, asnumpy


def h5_module():
    try:
        import h5py  # the real one, or minih5 installed under that name by dropin.install()
        return h5py
    except ImportError:
        from . import minih5
        return minih5


def _host(v):
    if isinstance(v, DeviceArray):
        return v.get()
    return np.asarray(asnumpy(v))


def _py(v):
    """attribute value as read back by h5py -> plain Python (np.int64 -> int, np.bool_ -> bool, bytes stay bytes)"""
    if isinstance(v, np.generic):
        return v.item()
    return v


def save_layer(layer, f, save_grads=True):
    info = f.create_dataset(layer.layer_name + "/layer_info", dtype=np.float32)
    info.attrs["type"] = type(layer).__name__
    for key in layer._h5_attrs:
        info.attrs[key] = getattr(layer, key)
    first = True
    for key in layer._h5_params:
        if key == "bias" and not getattr(layer, "with_bias", True):
            continue
        arr = _host(layer.learned_params[key])
        d = f.create_dataset(layer.layer_name + "/" + key, arr.shape, dtype=arr.dtype)
        d[:] = arr
        reg = getattr(layer, "weight_regulariser", None)
        if first and key == "weights" and reg is not None:
            d.attrs["weight_regulariser_type"] = np.bytes_(reg.type)
            d.attrs["weight_regulariser_strength"] = np.bytes_(reg.strength)
        first = False
        if save_grads:
            g = _host(layer.grads[key]).astype(arr.dtype, copy=False)
            gd = f.create_dataset(layer.layer_name + "/grads/" + key, g.shape, dtype=arr.dtype)
            gd[:] = g
    for key in layer._h5_state:
        v = layer.non_learned_params[key]
        if v is None:
            raise ValueError("{} {}: no {} yet (save after at least one training batch)".format(
                type(layer).__name__, layer.layer_name, key))
        arr = _host(v)
        d = f.create_dataset(layer.layer_name + "/" + key, arr.shape, dtype=arr.dtype)
        d[:] = arr


def load_layer(layer, f, load_grads=True):
    info = f[layer.layer_name + "/layer_info"].attrs
    for key in layer._h5_attrs:
        if key in layer._h5_optional_attrs:
            v = info.get(key, None)
            setattr(layer, key, _py(v) if v else layer._h5_optional_attrs[key])
        else:
            setattr(layer, key, _py(info[key]))
    if layer._h5_params and layer.learned_params is None:
        layer.learned_params, layer.grads = {}, {}
    for key in layer._h5_params:
        if key == "bias" and not getattr(layer, "with_bias", True):
            continue
        d = f[layer.layer_name + "/" + key]
        if key == "weights":
            rtype = d.attrs.get("weight_regulariser_type", None)
            if rtype:
                strength = d.attrs["weight_regulariser_strength"]
                if bytes(rtype) == b"l2":
                    from .regularisers.l2 import l2
                    layer.weight_regulariser = l2(strength=float(strength))
        layer.learned_params[key] = np.ascontiguousarray(d[:], np.float32)
        if load_grads:
            layer.grads[key] = np.ascontiguousarray(f[layer.layer_name + "/grads/" + key][:], np.float32)
        else:
            layer.grads[key] = np.zeros_like(layer.learned_params[key])
    for key in layer._h5_state:
        layer.non_learned_params[key] = np.ascontiguousarray(f[layer.layer_name + "/" + key][:], np.float32)
    layer.is_on_gpu = False  # parameters are host arrays again: the next forward uploads them
    layer._bufs = {}


def layer_registry():
    """type attr -> class, the set feed_forward_network.py:117-136 / residual_block.py:119-131 can rebuild"""
    from .layers.activations import ReLu
    from .layers.batch_norm import BatchNormLayer
    from .layers.convolution import ConvLayer
    from .layers.dense_layer import DenseLayer
    from .layers.depthwise_convolution import DepthwiseConvLayer
    from .layers.losses import SoftmaxWithCrossEntropy
    from .layers.pointwise_convolution import PointwiseConvLayer
    from .layers.pooling import GlobalAveragePoolingLayer
    from .layers.residual_block import ResidualBlock
    return {c.__name__: c for c in (ReLu, BatchNormLayer, ConvLayer, DenseLayer, DepthwiseConvLayer,
                                    SoftmaxWithCrossEntropy, PointwiseConvLayer, GlobalAveragePoolingLayer,
                                    ResidualBlock)}


def make_layer(type_name, layer_name):
    reg = layer_registry()
    type_name = _py(type_name)
    if isinstance(type_name, bytes):
        type_name = type_name.decode()
    if type_name not in reg:
        raise ValueError("checkpoint names a layer type this build cannot rebuild: {!r}".format(type_name))
    return reg[type_name](layer_name)


# ---- optimiser state (§8f-4) ----------------------------------------------------------------------------------------
OPT_GROUP = "__optimiser__"


def save_optimiser_state(opt, f):
    """`opt`: an optimiser of dorknet_b200.optimisers (state_dict() -> {(layer_name, param): ndarray})."""
    info = f.create_dataset(OPT_GROUP + "/info", dtype=np.float32)
    info.attrs["type"] = type(opt).__name__
    for k, v in opt.hyper_parameters().items():
        info.attrs[k] = v
    for (lname, pname), arr in opt.state_dict().items():
        d = f.create_dataset("{}/{}/{}".format(OPT_GROUP, lname, pname), arr.shape, dtype=arr.dtype)
        d[:] = arr


def load_optimiser_state(opt, f):
    if OPT_GROUP not in f:
        raise KeyError("checkpoint holds no optimiser state (written by the reference, or without one)")
    info = f[OPT_GROUP + "/info"].attrs
    if _py(info["type"]) != type(opt).__name__:
        raise ValueError("checkpoint optimiser state is {}'s, not {}'s".format(_py(info["type"]), type(opt).__name__))
    state = {}
    grp = f[OPT_GROUP]
    for lname in grp.keys():
        if lname == "info":
            continue
        for pname in grp[lname].keys():
            state[(lname, pname)] = np.asarray(grp[lname][pname][:], np.float32)
    opt.load_state_dict(state)
