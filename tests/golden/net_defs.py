"""Network definitions shared by make_golden.py (reference classes) and the GPU tests (our classes)."""
import numpy as np


def build_small_net(M, seed):
    """A miniature of the ResNet-18-depsep pattern (examples/imagenet_dogs_225_resnet_18_depsep.py):
    conv s2 - BN - ReLU - pw s2 - BN - ReLU - ResidualBlock(identity) - ResidualBlock(downsample,
    pw-s2 skip) - GAP - dense - softmax.  `M` is a namespace of layer classes (reference or ours)."""
    np.random.seed(seed)
    net = M.FeedForwardNetwork("mini")
    net.add_layer(M.ConvLayer("conv0", filter_block_shape=(8, 3, 5, 5), with_bias=False, stride=2, padding=1,
                              weight_regulariser=M.l2(1e-3)))
    net.add_layer(M.BatchNormLayer("conv0_bn", input_dimension=4, incoming_chans=8))
    net.add_layer(M.ReLu("conv0_relu"))
    net.add_layer(M.PointwiseConvLayer("pw0", filter_block_shape=(8, 8), with_bias=False, stride=2,
                                       weight_regulariser=M.l2(1e-3)))
    net.add_layer(M.BatchNormLayer("pw0_bn", input_dimension=4, incoming_chans=8))
    net.add_layer(M.ReLu("pw0_relu"))

    def unit(nm, cin, cout, stride, final_relu):
        ll = [M.DepthwiseConvLayer(nm + "_dw", filter_block_shape=(cin, 3, 3), stride=stride, padding=1, with_bias=False),
              M.BatchNormLayer(nm + "_dw_bn", input_dimension=4, incoming_chans=cin),
              M.PointwiseConvLayer(nm + "_pw", filter_block_shape=(cout, cin), with_bias=False,
                                   weight_regulariser=M.l2(1e-3)),
              M.BatchNormLayer(nm + "_pw_bn", input_dimension=4, incoming_chans=cout)]
        if final_relu:
            ll.append(M.ReLu(nm + "_relu"))
        return ll

    net.add_layer(M.ResidualBlock("res1", layer_list=unit("res1_a", 8, 8, 1, True) + unit("res1_b", 8, 8, 1, False),
                                  skip_projection=None, post_skip_activation=M.ReLu("res1_relu2")))
    net.add_layer(M.ResidualBlock("res2", layer_list=unit("res2_a", 8, 16, 2, True) + unit("res2_b", 16, 16, 1, False),
                                  skip_projection=M.PointwiseConvLayer("res2_skip", filter_block_shape=(16, 8), stride=2,
                                                                       with_bias=False, weight_regulariser=M.l2(1e-3)),
                                  post_skip_activation=M.ReLu("res2_relu2")))
    net.add_layer(M.GlobalAveragePoolingLayer("gap"))
    net.add_layer(M.DenseLayer("dense1", incoming_chans=16, output_dim=5, weight_regulariser=M.l2(1e-3)))
    net.set_loss_layer(M.SoftmaxWithCrossEntropy("softmax1"))
    return net


def iter_param_layers(net):
    for l in net.layers:
        if getattr(l, "learned_params", None):
            yield l
        if hasattr(l, "layer_list"):
            for m in l.layer_list:
                if getattr(m, "learned_params", None):
                    yield m
            if l.skip_projection is not None:
                yield l.skip_projection
