"""ResidualBlock (reference: layers/residual_block.py:12-151)."""
import numpy as np

from .layer import Layer, api, runtime, asarray
from .activations import ReLu
from .depthwise_convolution import DepthwiseConvLayer
from ..array import LazyBNOutput
from .batch_norm import BatchNormLayer
from .pointwise_convolution import PointwiseConvLayer


class ResidualBlock(Layer):
    """out = post_skip_activation(layer_list(X) + skip_projection(X)); skip_projection=None is the
    identity.  Same constructor and traversal as the reference; the join `X_tmp + skippee` followed
    by ReLU (residual_block.py:75) is one kernel, and in backward the join `dx + joined_dx`
    (:93-95) is folded into the first branch layer's backward when that layer supports it."""

    def __init__(self, layer_name, layer_list=None, skip_projection=None, post_skip_activation=None):
        super().__init__(layer_name)
        self.layer_list = layer_list
        self.skip_projection = skip_projection
        self.post_skip_activation = post_skip_activation
        self.fuse_join = True  # fold `X_tmp + skippee` + ReLU into the branch's last BatchNorm when it is still deferred
        if layer_list is None:
            self.layer_list = []

    def __repr__(self):
        return "ResidualBlock({}, layer_list={}, skip_projection={}, post_skip_activation={})".format(
            self.layer_name, self.layer_list, self.skip_projection, self.post_skip_activation)

    def to_gpu(self):
        if self.is_on_gpu:
            print("Layer already on GPU, ignoring request")
            return
        for layer in self.layer_list:
            layer.to_gpu()
        if self.skip_projection is not None:
            self.skip_projection.to_gpu()
        if self.post_skip_activation is not None:
            self.post_skip_activation.to_gpu()
        self.is_on_gpu = True

    def forward(self, X, test_mode=False):
        self._ensure_gpu()
        X = asarray(X)
        X_tmp = self.layer_list[0].forward(X, test_mode=test_mode)
        for layer in self.layer_list[1:]:
            X_tmp = layer.forward(X_tmp, test_mode=test_mode)
        skippee = self.skip_projection.forward(X, test_mode=test_mode) if self.skip_projection is not None else X
        act = self.post_skip_activation
        if type(act) is ReLu:
            if X_tmp.shape != skippee.shape:
                raise ValueError("operands could not be broadcast together with shapes {} {}".format(
                    X_tmp.shape, skippee.shape))
            y = act._buf("y", X_tmp.shape)
            if (isinstance(X_tmp, LazyBNOutput) and X_tmp.fusable and not test_mode and self.fuse_join):
                # the branch ends in a BatchNorm whose normalisation pass has not run: it adds the skip and applies
                # the ReLU itself (one pass over the activation instead of three)
                X_tmp.bn.fused_add_relu_apply(y, skippee)  # (a lazy skip output materialises on .ptr)
                X_tmp.consume()
            else:
                api.dk_add_relu_fwd(X_tmp.ptr, skippee.ptr, y.ptr, y.size, runtime.stream())
            act._y = y  # the reference calls the activation without test_mode, so it always records (:75)
            return y
        return act.forward(X_tmp + skippee)

    def regulariser_forward(self):
        """Sums over layer_list only -- the skip projection's l2 term is not in the loss although its
        gradient carries strength*W (residual_block.py:78-84)."""
        regularisation = 0
        for l in self.layer_list:
            if hasattr(l, "regulariser_forward"):
                regularisation += l.regulariser_forward()
        return regularisation

    def backward(self, upstream_dx):
        act, last = self.post_skip_activation, self.layer_list[-1]
        if (self.fuse_join and type(act) is ReLu and act._fused_bn is None and type(last) is BatchNormLayer
                and last.input_dimension == 4 and not last._relu_fused and act._y is not None
                and tuple(act._y.shape) == tuple(last.input_shape)):
            # the block's ReLU backward and the last BatchNorm's backward in one kernel; joined_dx is still produced
            # (the skip path needs it)
            upstream_dx = asarray(upstream_dx)
            joined_dx = act._buf("dx", upstream_dx.shape)
            dx = last.backward_join(upstream_dx, act._y, joined_dx)
        else:
            joined_dx = act.backward(upstream_dx)
            dx = last.backward(joined_dx)
        first = self.layer_list[0]
        for l in self.layer_list[-2:0:-1]:
            dx = l.backward(dx)
        if len(self.layer_list) == 1:
            first_dx_in = None  # already consumed above
        else:
            first_dx_in = dx
        skip_dx = self.skip_projection.backward(joined_dx) if self.skip_projection is not None else joined_dx
        if first_dx_in is None:
            return dx + skip_dx
        if type(first) is DepthwiseConvLayer and tuple(first.input_shape) == tuple(skip_dx.shape):
            return first.backward(first_dx_in, dx_add=skip_dx)
        return first.backward(first_dx_in) + skip_dx

    # -- checkpoints (residual_block.py:99-152): the block's layer_info lists its members by type and name, the
    #    members are ordinary top-level groups of the file
    def save_to_h5(self, open_f, save_grads=True):
        info = open_f.create_dataset(self.layer_name + "/layer_info", dtype=np.float32)
        info.attrs["type"] = type(self).__name__
        info.attrs["layer_type_list"] = [type(l).__name__ for l in self.layer_list]
        info.attrs["layer_name_list"] = [l.layer_name for l in self.layer_list]
        info.attrs["post_skip_activation_type"] = type(self.post_skip_activation).__name__
        info.attrs["post_skip_activation_name"] = self.post_skip_activation.layer_name
        members = list(self.layer_list)
        if self.skip_projection is not None:
            info.attrs["skip_projection_type"] = type(self.skip_projection).__name__
            info.attrs["skip_projection_name"] = self.skip_projection.layer_name
            members.append(self.skip_projection)
        members.append(self.post_skip_activation)
        for l in members:
            l.save_to_h5(open_f, save_grads=save_grads)

    def load_from_h5(self, open_f, load_grads=True):
        from ..checkpoint import make_layer
        info = open_f[self.layer_name + "/layer_info"].attrs
        self.layer_list = [make_layer(t, str(n) if not isinstance(n, bytes) else n.decode())
                           for t, n in zip(info["layer_type_list"], info["layer_name_list"])]
        for l in self.layer_list:
            l.load_from_h5(open_f, load_grads=load_grads)
        self.skip_projection = None
        if info.get("skip_projection_type", None):
            self.skip_projection = make_layer(info["skip_projection_type"], info["skip_projection_name"])
            if type(self.skip_projection) is not PointwiseConvLayer:
                print("ResidualBlock: Unrecognised skip_projection type {}".format(info["skip_projection_type"]))
            self.skip_projection.load_from_h5(open_f, load_grads=load_grads)
        self.post_skip_activation = make_layer(info["post_skip_activation_type"], info["post_skip_activation_name"])
        self.post_skip_activation.load_from_h5(open_f, load_grads=load_grads)
        self.is_on_gpu = False
