"""The BASELINE.json workloads as network builders.

Every builder takes a namespace `M` of layer classes (FeedForwardNetwork, ConvLayer, ...): pass
`dorknet_b200.workloads.ours()` for the sm_100a path, or the reference's own classes (tests and the
benchmark's reference arm do, from their side) to build the very same network on the reference's CPU path.
Layer names follow the reference example so parameters can be copied across by name.
"""
import types

import numpy as np


def ours():
    from .layers.convolution import ConvLayer
    from .layers.depthwise_convolution import DepthwiseConvLayer
    from .layers.pointwise_convolution import PointwiseConvLayer
    from .layers.batch_norm import BatchNormLayer
    from .layers.activations import ReLu
    from .layers.pooling import GlobalAveragePoolingLayer, MaxPoolLayer
    from .layers.dense_layer import DenseLayer
    from .layers.residual_block import ResidualBlock
    from .layers.losses import SoftmaxWithCrossEntropy
    from .network.feed_forward_network import FeedForwardNetwork
    from .optimisers.SGD import SGD
    from .optimisers.SGDMomentum import SGDMomentum
    from .optimisers.RMSProp import RMSProp
    from .regularisers.l2 import l2
    return types.SimpleNamespace(**{k: v for k, v in locals().items()})


def depthwise_sep_unit(M, name, cin, cout, stride=1, l2_strength=1e-4, final_relu=True):
    """dw3x3 - BN - pw - BN [- ReLU] with the reference's default flags
    (examples/imagenet_dogs_225_resnet_18_depsep.py:34-70: batch_norm_depthwise=True,
    relu_depthwise=False, no bias, l2 on the pointwise weights only)."""
    layers = [
        M.DepthwiseConvLayer(name + "_dw", filter_block_shape=(cin, 3, 3), stride=stride, padding=1,
                             with_bias=False, weight_regulariser=None),
        M.BatchNormLayer(name + "_dw_bn", input_dimension=4, incoming_chans=cin),
        M.PointwiseConvLayer(name + "_pw", filter_block_shape=(cout, cin), with_bias=False,
                             weight_regulariser=M.l2(strength=l2_strength)),
        M.BatchNormLayer(name + "_pw_bn", input_dimension=4, incoming_chans=cout),
    ]
    if final_relu:
        layers.append(M.ReLu(name + "pw_relu"))
    return layers


def build_resnet18_depsep(M, classes=120, conv0_padding=1, seed=0):
    """ResNet-18 depthwise-separable of examples/imagenet_dogs_225_resnet_18_depsep.py:32-160.

    conv0_padding=1 is the reference network and needs 225x225 inputs (at 224 the reference's
    pw0.backward yields 112 != 111 rows); conv0_padding=2 is the 224x224 variant."""
    np.random.seed(seed)
    net = M.FeedForwardNetwork("ResNet18_depsep")
    net.add_layer(M.ConvLayer("conv0", filter_block_shape=(64, 3, 5, 5), with_bias=False, stride=2,
                              padding=conv0_padding, weight_regulariser=M.l2(0.0001)))
    net.add_layer(M.BatchNormLayer("conv0_bn", input_dimension=4, incoming_chans=64))
    net.add_layer(M.ReLu("conv0_relu"))
    net.add_layer(M.PointwiseConvLayer("pw0", filter_block_shape=(64, 64), with_bias=False, stride=2,
                                       weight_regulariser=M.l2(0.0001)))
    net.add_layer(M.BatchNormLayer("pw0_bn", input_dimension=4, incoming_chans=64))
    net.add_layer(M.ReLu("pw0_relu"))

    def res_block(name, cout, cin, downsample=False):
        ll = depthwise_sep_unit(M, name + "_dw1", cin, cout, stride=2 if downsample else 1, final_relu=True)
        ll += depthwise_sep_unit(M, name + "_dw2", cout, cout, stride=1, final_relu=False)
        skip = None
        if downsample:
            skip = M.PointwiseConvLayer(name + "_pw_skip", filter_block_shape=(cout, cin), stride=2,
                                        with_bias=False, weight_regulariser=M.l2(strength=0.0001))
        net.add_layer(M.ResidualBlock(name, layer_list=ll, skip_projection=skip,
                                      post_skip_activation=M.ReLu(name + "_relu2")))

    res_block("res1", 64, 64)
    res_block("res2", 64, 64)
    res_block("res3", 128, 64, downsample=True)
    res_block("res4", 128, 128)
    res_block("res5", 256, 128, downsample=True)
    res_block("res6", 256, 256)
    res_block("res7", 512, 256, downsample=True)
    res_block("res8", 512, 512)
    net.add_layer(M.GlobalAveragePoolingLayer("global_pool1"))
    net.add_layer(M.DenseLayer("dense1", incoming_chans=512, output_dim=classes, weight_regulariser=M.l2(0.0001)))
    net.set_loss_layer(M.SoftmaxWithCrossEntropy("softmax1"))
    return net


def build_mnist_convnet(M, seed=0):
    """MNISTNet of examples/MNIST_basic_convnet.py:15-69 (cfg1)."""
    np.random.seed(seed)
    net = M.FeedForwardNetwork("MNISTNet")

    def block(i, shape, stride, pad, reg):
        net.add_layer(M.ConvLayer("conv%d" % i, filter_block_shape=shape, stride=stride, padding=pad,
                                  with_bias=False, weight_regulariser=M.l2(reg)))
        net.add_layer(M.BatchNormLayer("bn%d" % i, input_dimension=4, incoming_chans=shape[0]))
        net.add_layer(M.ReLu("relu%d" % i))

    block(1, (32, 1, 3, 3), 1, 1, 1e-4)
    block(2, (32, 32, 3, 3), 1, 1, 1e-4)
    block(3, (64, 32, 4, 4), 2, 1, 1e-4)
    block(4, (64, 64, 3, 3), 1, 1, 1e-4)
    block(5, (128, 64, 4, 4), 2, 1, 1e-4)
    net.add_layer(M.GlobalAveragePoolingLayer("global_pool"))
    net.add_layer(M.DenseLayer("dense1", incoming_chans=128, output_dim=10, weight_regulariser=M.l2(5e-4)))
    net.set_loss_layer(M.SoftmaxWithCrossEntropy("softmax"))
    return net


MOBILENET_PLAN = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2)] + [(512, 1)] * 5 + [(1024, 2), (1024, 1)]


def build_mobilenet_depsep(M, classes=120, seed=0):
    """cfg5: the MobileNet-style flat depsep stack frozen in SURVEY.md §8(d) (the reference only ships the
    depthwise_sep_layer(..., add_layers=True) helper for it)."""
    np.random.seed(seed)
    net = M.FeedForwardNetwork("MobileNet_depsep")
    net.add_layer(M.ConvLayer("conv0", filter_block_shape=(32, 3, 3, 3), with_bias=False, stride=2, padding=1,
                              weight_regulariser=M.l2(0.0001)))
    net.add_layer(M.BatchNormLayer("conv0_bn", input_dimension=4, incoming_chans=32))
    net.add_layer(M.ReLu("conv0_relu"))
    cin = 32
    for i, (cout, stride) in enumerate(MOBILENET_PLAN):
        for layer in depthwise_sep_unit(M, "ds%d" % (i + 1), cin, cout, stride=stride, final_relu=True):
            net.add_layer(layer)
        cin = cout
    net.add_layer(M.GlobalAveragePoolingLayer("global_pool1"))
    net.add_layer(M.DenseLayer("dense1", incoming_chans=cin, output_dim=classes, weight_regulariser=M.l2(0.0001)))
    net.set_loss_layer(M.SoftmaxWithCrossEntropy("softmax1"))
    return net


def iter_param_layers(net, include_skip=True):
    for l in net.layers:
        if getattr(l, "learned_params", None):
            yield l
        if hasattr(l, "layer_list"):
            for m in l.layer_list:
                if getattr(m, "learned_params", None):
                    yield m
            if include_skip and getattr(l, "skip_projection", None) is not None:
                yield l.skip_projection


def synthetic_batch_u8(batch, channels, size, classes, seed=0, mixup=False):
    """The synthetic batch as a data loader would hold it: decoded images, uint8 NHWC in [0, 255] (cv2 order), one-hot
    labels, and for mixup a second batch + lam ~ U(0, 0.3) (data_loading/image_data_loader.py:100-112).
    Returns dict(img, y, Y[, img_b, y_b, Y_b, lam]); Y is the (blended) label matrix the loss sees."""
    rng = np.random.default_rng(seed)
    out = {"img": rng.integers(0, 256, size=(batch, size, size, channels), dtype=np.uint8)}
    y = rng.integers(0, classes, size=batch)
    Y = np.zeros((batch, classes), np.float32)
    Y[np.arange(batch), y] = 1.0
    out["y"] = y
    if mixup:
        out["img_b"] = rng.integers(0, 256, size=(batch, size, size, channels), dtype=np.uint8)
        yb = rng.integers(0, classes, size=batch)
        Yb = np.zeros_like(Y)
        Yb[np.arange(batch), yb] = 1.0
        lam = np.float32(rng.uniform(0.0, 0.3))
        out["lam"] = float(lam)
        Y = lam * Yb + (1 - lam) * Y
    out["Y"] = Y.astype(np.float32)
    return out


def synthetic_batch(batch, channels, size, classes, seed=0, mixup=False):
    """images in [0,255] - 128 as fp32 NCHW (image_preprocessor.py:36-37 on the uint8 images of
    synthetic_batch_u8), one-hot labels; with mixup two batches are blended with lam ~ U(0, 0.3) as
    data_loading/image_data_loader.py:100-112 does."""
    raw = synthetic_batch_u8(batch, channels, size, classes, seed=seed, mixup=mixup)
    X = raw["img"].transpose(0, 3, 1, 2).astype(np.float32) - np.float32(128.0)
    if mixup:
        Xb = raw["img_b"].transpose(0, 3, 1, 2).astype(np.float32) - np.float32(128.0)
        lam = np.float32(raw["lam"])
        X = (np.float32(1.0) - lam) * X + lam * Xb
    return np.ascontiguousarray(X, np.float32), raw["y"], raw["Y"]


# ---- algorithmic bytes / flops (unfused per-layer minimum) ------------------------
def algorithmic_cost(net, input_shape):
    """Walk the layer list with shapes only.  Returns a list of dict(name, kind, fwd_bytes, bwd_bytes,
    fwd_flops, bwd_flops) -- the denominators bench.py reports rooflines against."""
    rows = []

    def visit(layer, shape):
        kind = type(layer).__name__
        n = int(np.prod(shape))
        r = dict(name=layer.layer_name, kind=kind, fwd_bytes=0, bwd_bytes=0, fwd_flops=0, bwd_flops=0, in_shape=shape)
        if kind == "ConvLayer":
            N, C, H, W = shape
            F, k, s, p = layer.num_filters, layer.f_rows, layer.stride, layer.padding
            OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - layer.f_cols) // s + 1
            out = (N, F, OH, OW)
            no = int(np.prod(out))
            fl = 2 * N * OH * OW * F * C * k * layer.f_cols
            r.update(fwd_bytes=4 * (n + no), bwd_bytes=4 * (2 * no + 2 * n), fwd_flops=fl, bwd_flops=2 * fl)
        elif kind == "PointwiseConvLayer":
            N, C, H, W = shape
            F, s = layer.num_filters, layer.stride
            OH, OW = (H - 1) // s + 1, (W - 1) // s + 1
            out = (N, F, OH, OW)
            no = int(np.prod(out))
            used = N * C * OH * OW
            fl = 2 * N * OH * OW * F * C
            r.update(fwd_bytes=4 * (used + no), bwd_bytes=4 * (2 * no + used + N * C * OH * s * OW * s),
                     fwd_flops=fl, bwd_flops=2 * fl)
        elif kind == "DepthwiseConvLayer":
            N, C, H, W = shape
            k, s, p = layer.f_rows, layer.stride, layer.padding
            OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - layer.f_cols) // s + 1
            out = (N, C, OH, OW)
            no = int(np.prod(out))
            fl = 2 * no * k * layer.f_cols
            r.update(fwd_bytes=4 * (n + no), bwd_bytes=4 * (no + 2 * n), fwd_flops=fl, bwd_flops=2 * fl)
        elif kind == "BatchNormLayer":
            out = shape
            r.update(fwd_bytes=4 * 3 * n, bwd_bytes=4 * 5 * n)
        elif kind == "ReLu":
            out = shape
            r.update(fwd_bytes=4 * 2 * n, bwd_bytes=4 * 3 * n)
        elif kind == "GlobalAveragePoolingLayer":
            out = shape[:2]
            r.update(fwd_bytes=4 * (n + int(np.prod(out))), bwd_bytes=4 * (n + int(np.prod(out))))
        elif kind == "DenseLayer":
            B, D = shape
            out = (B, layer.output_dim)
            fl = 2 * B * D * layer.output_dim
            r.update(fwd_bytes=4 * (n + B * layer.output_dim), bwd_bytes=4 * (2 * B * layer.output_dim + 2 * n),
                     fwd_flops=fl, bwd_flops=2 * fl)
        elif kind == "ResidualBlock":
            s2 = shape
            for l in layer.layer_list:
                s2 = visit(l, s2)
            if layer.skip_projection is not None:
                visit(layer.skip_projection, shape)
            no = int(np.prod(s2))
            rows.append(dict(name=layer.layer_name + "_add", kind="ResidualAdd", fwd_bytes=4 * 3 * no,
                             bwd_bytes=4 * 3 * no, fwd_flops=0, bwd_flops=0, in_shape=s2))
            visit(layer.post_skip_activation, s2)
            return s2
        else:
            out = shape
        rows.append(r)
        return tuple(out)

    shape = tuple(input_shape)
    for layer in net.layers:
        shape = visit(layer, shape)
    return rows


def total_cost(rows):
    return dict(bytes=sum(r["fwd_bytes"] + r["bwd_bytes"] for r in rows),
                flops=sum(r["fwd_flops"] + r["bwd_flops"] for r in rows))
