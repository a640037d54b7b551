"""Inference path: BatchNorm folded into the preceding convolution / pointwise / depthwise / dense
weights.

The reference's test mode (`network.forward(..., test_mode=True)`, feed_forward_network.py:47-88) still runs every
BatchNormLayer as its own pass, `gamma * (X - running_mean) / running_std + beta` (batch_norm.py:101-115) -- 34 extra
activation-sized read+write passes in ResNet-18-depsep.  In test mode that is an affine map per channel, and an affine
map after a linear layer is the same linear layer with other weights:

    scale[f] = gamma[f] / running_std[f]            shift[f] = beta[f] - running_mean[f] * scale[f]
    W'[f, ...] = W[f, ...] * scale[f]               b'[f]    = b[f] * scale[f] + shift[f]

`fold_batchnorm(net)` returns a NEW network of the same layer classes (parameters copied, the trained network is not
touched) in which every `linear -> BatchNorm` pair is one layer with bias; BatchNorms that do not follow a linear layer
are kept.  `terminal_layer_name` early exit (CAM demo, examples/*_CAM.py:77-80) keeps working for every layer that
survives; asking for a folded-away BatchNorm's name raises.
"""
import numpy as np

from .array import asnumpy


def _host(v):
    return np.array(asnumpy(v), dtype=np.float32, copy=True)


def _bn_affine(bn):
    rm, rs = bn.non_learned_params["running_mean"], bn.non_learned_params["running_std"]
    if rm is None:
        raise ValueError("BatchNormLayer {}: no running statistics (train or load a checkpoint first)".format(bn.layer_name))
    g, b = _host(bn.learned_params["gamma"]).reshape(-1), _host(bn.learned_params["beta"]).reshape(-1)
    rm, rs = _host(rm).reshape(-1), _host(rs).reshape(-1)
    scale = g / rs
    return scale, b - rm * scale


def _clone_linear(layer, bn):
    """A copy of `layer` (Conv / Depthwise / Pointwise / Dense) whose weights absorb `bn` (None: plain copy)."""
    from .layers.dense_layer import DenseLayer
    cls = type(layer)
    new = cls(layer.layer_name)
    for k in cls._h5_attrs:
        setattr(new, k, getattr(layer, k))
    new.weight_regulariser = None
    W = _host(layer.learned_params["weights"])
    bias = _host(layer.learned_params["bias"]) if layer.with_bias else None
    if bn is not None:
        scale, shift = _bn_affine(bn)
        if cls is DenseLayer:  # W is [in, out] (dense_layer.py:19-26)
            if scale.shape[0] != W.shape[1]:
                raise ValueError("BatchNorm {} does not match {}".format(bn.layer_name, layer.layer_name))
            W = W * scale[np.newaxis, :]
        else:
            if scale.shape[0] != W.shape[0]:
                raise ValueError("BatchNorm {} does not match {}".format(bn.layer_name, layer.layer_name))
            W = W * scale.reshape((-1,) + (1,) * (W.ndim - 1))
        bias = shift if bias is None else bias * scale + shift
        new.with_bias = True
    new.learned_params = {"weights": np.ascontiguousarray(W, np.float32)}
    new.grads = {"weights": np.zeros_like(W)}
    if bias is not None:
        new.learned_params["bias"] = np.ascontiguousarray(bias, np.float32)
        new.grads["bias"] = np.zeros_like(new.learned_params["bias"])
    return new


def _clone_bn(bn):
    from .layers.batch_norm import BatchNormLayer
    new = BatchNormLayer(bn.layer_name, input_dimension=bn.input_dimension, incoming_chans=bn.incoming_chans,
                         run_momentum=bn.run_momentum)
    new.eps = bn.eps
    new.learned_params = {k: _host(v) for k, v in bn.learned_params.items()}
    new.grads = {k: np.zeros_like(v) for k, v in new.learned_params.items()}
    new.non_learned_params["running_mean"] = _host(bn.non_learned_params["running_mean"])
    new.non_learned_params["running_std"] = _host(bn.non_learned_params["running_std"])
    return new


def _fold_list(layers, folded_names):
    from .layers.activations import ReLu
    from .layers.batch_norm import BatchNormLayer
    from .layers.convolution import ConvLayer
    from .layers.dense_layer import DenseLayer
    from .layers.depthwise_convolution import DepthwiseConvLayer
    from .layers.pointwise_convolution import PointwiseConvLayer
    from .layers.pooling import GlobalAveragePoolingLayer, MaxPoolLayer
    from .layers.residual_block import ResidualBlock
    linear = (ConvLayer, DepthwiseConvLayer, PointwiseConvLayer, DenseLayer)
    out, i = [], 0
    while i < len(layers):
        l = layers[i]
        nxt = layers[i + 1] if i + 1 < len(layers) else None
        if isinstance(l, linear):
            dims_match = isinstance(nxt, BatchNormLayer) and nxt.input_dimension == (2 if isinstance(l, DenseLayer) else 4)
            if dims_match:
                out.append(_clone_linear(l, nxt))
                folded_names.append(nxt.layer_name)
                i += 2
                continue
            out.append(_clone_linear(l, None))
        elif isinstance(l, BatchNormLayer):
            out.append(_clone_bn(l))
        elif isinstance(l, ResidualBlock):
            skip = _clone_linear(l.skip_projection, None) if l.skip_projection is not None else None
            act = l.post_skip_activation
            out.append(ResidualBlock(l.layer_name, layer_list=_fold_list(l.layer_list, folded_names), skip_projection=skip,
                                     post_skip_activation=type(act)(act.layer_name) if act is not None else None))
        elif isinstance(l, MaxPoolLayer):
            out.append(MaxPoolLayer(l.layer_name, stride=l.stride))
        elif isinstance(l, (ReLu, GlobalAveragePoolingLayer)):
            out.append(type(l)(l.layer_name))
        else:
            raise TypeError("fold_batchnorm: no rule for layer {!r}".format(l))
        i += 1
    return out


class FoldedNetwork:
    """What fold_batchnorm returns: `.forward(X, y_one_hot=None, test_mode=True, terminal_layer_name=None)` and
    `.test(...)` of the container (feed_forward_network.py:47-88), inference only."""

    def __init__(self, net, folded_names):
        self._net = net
        self.name = net.name
        self.layers = net.layers
        self.loss_layer = net.loss_layer
        self.folded_batchnorms = tuple(folded_names)

    def forward(self, X, y_one_hot=None, test_mode=True, terminal_layer_name=None):
        if not test_mode:
            raise ValueError("a BatchNorm-folded network is inference only: train the original network")
        if terminal_layer_name in self.folded_batchnorms:
            raise KeyError("layer {} was folded into the layer before it; stop at that layer instead".format(
                terminal_layer_name))
        return self._net.forward(X, y_one_hot, test_mode=True, terminal_layer_name=terminal_layer_name)

    def test(self, data_loader, batch_size, test_set_size):
        return self._net.test(data_loader, batch_size, test_set_size)

    def to_gpu(self):
        self._net.to_gpu()


def fold_batchnorm(net):
    """`net`: a trained dorknet_b200 FeedForwardNetwork (or the reference's container holding dorknet_b200 layers)."""
    from .layers.losses import SoftmaxWithCrossEntropy
    from .network.feed_forward_network import FeedForwardNetwork
    folded = []
    new = FeedForwardNetwork(net.name)
    for l in _fold_list(list(net.layers), folded):
        new.add_layer(l)
    if net.loss_layer is not None:
        new.set_loss_layer(SoftmaxWithCrossEntropy(net.loss_layer.layer_name))
    return FoldedNetwork(new, folded)
