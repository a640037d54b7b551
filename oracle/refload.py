"""Loader for the built reference binaries in oracle/_ref (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  It makes the UNMODIFIED reference CPU path importable under its own
module names (layers.*, network.*, optimisers.*, regularisers.*; top-level im2col, ...),
as built by oracle/build_ref.py from /root/reference.  Because the product package mirrors
those module names, never call this in a process that also ran dorknet_b200.dropin.install().
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")


def available():
    return os.path.exists(os.path.join(REF, ".built"))


def load_reference():
    """Return a SimpleNamespace of the reference's public classes (CPU path)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists")
    for name in ("layers", "network", "optimisers", "regularisers"):
        m = sys.modules.get(name)
        if m is not None and REF not in "".join(map(str, getattr(m, "__path__", []))):
            raise RuntimeError("module %r is already bound to something that is not the reference" % name)
    for p in (REF, STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # NumPy 2 removed numpy.lib.function_base; the reference has an unused
    # `from numpy.lib.function_base import select` (/root/reference/layers/depthwise_convolution.py:6).
    if "numpy.lib.function_base" not in sys.modules:
        import numpy as np
        shim = types.ModuleType("numpy.lib.function_base")
        shim.select = np.select
        sys.modules["numpy.lib.function_base"] = shim
    ns = types.SimpleNamespace()
    imp = importlib.import_module
    ns.im2col = imp("im2col")
    ns.pooling_cy = imp("pooling_cy")
    ns.relu_cy = imp("relu_cy")
    ns.batch_norm_stats_cy = imp("batch_norm_stats_cy")
    ns.ConvLayer = imp("layers.convolution").ConvLayer
    ns.DepthwiseConvLayer = imp("layers.depthwise_convolution").DepthwiseConvLayer
    ns.PointwiseConvLayer = imp("layers.pointwise_convolution").PointwiseConvLayer
    ns.BatchNormLayer = imp("layers.batch_norm").BatchNormLayer
    ns.ReLu = imp("layers.activations").ReLu
    ns.GlobalAveragePoolingLayer = imp("layers.pooling").GlobalAveragePoolingLayer
    ns.MaxPoolLayer = imp("layers.pooling").MaxPoolLayer
    ns.DenseLayer = imp("layers.dense_layer").DenseLayer
    ns.ResidualBlock = imp("layers.residual_block").ResidualBlock
    ns.SoftmaxWithCrossEntropy = imp("layers.losses").SoftmaxWithCrossEntropy
    ns.FeedForwardNetwork = imp("network.feed_forward_network").FeedForwardNetwork
    ns.SGD = imp("optimisers.SGD").SGD
    ns.SGDMomentum = imp("optimisers.SGDMomentum").SGDMomentum
    ns.RMSProp = imp("optimisers.RMSProp").RMSProp
    ns.l2 = imp("regularisers.l2").l2
    return ns
