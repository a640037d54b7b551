"""CPU: host-side logic of the Python mirror that needs no GPU."""
import subprocess
import sys
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeLayer:
    def __init__(self, name, params=True):
        self.layer_name = name
        self.learned_params = {"weights": np.zeros(2, np.float32)} if params else None
        self.grads = {"weights": np.zeros(2, np.float32)} if params else None


class FakeBlock:
    def __init__(self, name, inner, skip):
        self.layer_name = name
        self.learned_params = None
        self.layer_list = inner
        self.skip_projection = skip


class FakeNet:
    def __init__(self, layers):
        self.layers = layers


def test_optimiser_update_sets_reproduce_reference_quirks():
    """SGDMomentum descends into layer_list but never sees skip_projection;
    SGD / RMSProp never update anything inside a ResidualBlock."""
    from dorknet_b200.optimisers._multi import collect_layers
    a, b, c, skip, relu = FakeLayer("a"), FakeLayer("b"), FakeLayer("c"), FakeLayer("skip"), FakeLayer("r", False)
    net = FakeNet([a, relu, FakeBlock("blk", [b, relu, c], skip)])
    assert [l.layer_name for l in collect_layers(net, descend=True)] == ["a", "b", "c"]
    assert [l.layer_name for l in collect_layers(net, descend=False)] == ["a"]
    assert [l.layer_name for l in collect_layers(net, descend=True, include_skip=True)] == ["a", "b", "c", "skip"]


def test_sgd_and_rmsprop_fixed_traversal_is_opt_in_and_state_round_trips():
    """the reference update sets stay the default; fixed_traversal / include_skip_projections widen them;
    state_dict / load_state_dict carry velocities by (layer name, parameter)."""
    from dorknet_b200.optimisers.SGD import SGD
    from dorknet_b200.optimisers.RMSProp import RMSProp
    from dorknet_b200.optimisers.SGDMomentum import SGDMomentum
    a, b, c, skip, relu = FakeLayer("a"), FakeLayer("b"), FakeLayer("c"), FakeLayer("skip"), FakeLayer("r", False)
    net = FakeNet([a, relu, FakeBlock("blk", [b, relu, c], skip)])
    names = lambda o: [l.layer_name for l in o.learnable_layers]  # noqa: E731
    assert names(SGD(net, 0.1)) == ["a"] and names(RMSProp(net, 0.1, 0.9)) == ["a"]
    assert names(SGD(net, 0.1, fixed_traversal=True)) == ["a", "b", "c"]
    assert names(RMSProp(net, 0.1, 0.9, fixed_traversal=True)) == ["a", "b", "c"]
    assert names(SGD(net, 0.1, include_skip_projections=True)) == ["a", "b", "c", "skip"]
    assert names(RMSProp(net, 0.1, 0.9, fixed_traversal=True, include_skip_projections=True)) == ["a", "b", "c", "skip"]
    opt = SGDMomentum(net, 0.1, 0.9)
    assert opt.hyper_parameters() == {"learning_rate": 0.1, "momentum": 0.9}
    sd = opt.state_dict()
    assert sorted(sd) == [("a", "weights"), ("b", "weights"), ("c", "weights")] and all(v.shape == (2,) for v in sd.values())
    sd[("b", "weights")] = np.array([1.0, 2.0], np.float32)
    opt.load_state_dict(sd)
    np.testing.assert_array_equal(opt.state_dict()[("b", "weights")], [1.0, 2.0])
    with pytest.raises(KeyError):
        opt.load_state_dict({("a", "weights"): np.zeros(2, np.float32)})
    with pytest.raises(ValueError):
        opt.load_state_dict({k: np.zeros(3, np.float32) for k in sd})
    assert SGD(net, 0.1).state_dict() == {}


def test_constructors_match_reference_signatures_and_init():
    from dorknet_b200.layers.convolution import ConvLayer
    from dorknet_b200.layers.depthwise_convolution import DepthwiseConvLayer
    from dorknet_b200.layers.pointwise_convolution import PointwiseConvLayer
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.dense_layer import DenseLayer
    np.random.seed(3)
    c = ConvLayer("c", filter_block_shape=(4, 3, 5, 5), stride=2, padding=1, with_bias=True)
    np.random.seed(3)
    ref_w = 0.01 * np.random.randn(4, 3, 5, 5).astype(np.float32)  # convolution.py:26-27
    assert np.array_equal(c.learned_params["weights"], ref_w)
    assert c.learned_params["bias"].shape == (4,) and c.grads["weights"].shape == (4, 3, 5, 5)
    d = DepthwiseConvLayer("d", filter_block_shape=(6, 3, 3), weight_initialiser="glorot_uniform", with_bias=False)
    assert np.abs(d.learned_params["weights"]).max() <= np.sqrt(6.0 / 12)  # depthwise_convolution.py:26
    p = PointwiseConvLayer("p", stride=2, filter_block_shape=(8, 6))
    assert p.learned_params["weights"].shape == (8, 6)
    bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=6)
    assert bn.learned_params["gamma"].shape == (1, 6, 1, 1) and bn.non_learned_params["running_std"] is None
    with pytest.raises(ValueError):
        BatchNormLayer("bad", input_dimension=3, incoming_chans=2)
    dl = DenseLayer("fc", incoming_chans=5, output_dim=3)
    assert dl.learned_params["weights"].shape == (5, 3)
    assert "ConvLayer(c, filter_block_shape=(4,3,5,5)" in repr(c)


def test_device_scalar_is_lazy_linear_algebra():
    from dorknet_b200.array import DeviceScalar

    class Slot:
        def __init__(self, v):
            self.v = v

        def get(self):
            return np.array([self.v], np.float32)

    a, b = DeviceScalar([(Slot(2.0), 1.0)]), DeviceScalar([(Slot(3.0), 0.5)])
    total = 0
    total += a
    total += sum([b, 0, b])
    assert isinstance(total, DeviceScalar) and float(total) == pytest.approx(2.0 + 1.5 + 1.5)
    assert float(0.1 * total + 1) == pytest.approx(0.5 + 1)


def test_dropin_binds_reference_module_names_in_a_clean_process():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from dorknet_b200 import dropin; dropin.install()\n"
        "from layers.convolution import ConvLayer\n"
        "from layers.residual_block import ResidualBlock\n"
        "from regularisers.l2 import l2\n"
        "from optimisers.SGDMomentum import SGDMomentum\n"
        "import cupy as cp, numpy as np\n"
        "import dorknet_b200.layers.convolution as m\n"
        "assert ConvLayer is m.ConvLayer and cp.get_array_module(np.zeros(1)) is np\n"
        "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/network/feed_forward_network.py"),
                    reason="reference tree not present")
def test_reference_container_imports_unchanged_on_top_of_our_layers():
    """The reference's network/feed_forward_network.py, byte for byte, resolves its imports to us."""
    code = (
        "import sys, importlib.util; sys.path.insert(0, %r)\n"
        "from dorknet_b200 import dropin; dropin.install()\n"
        "spec = importlib.util.spec_from_file_location('network.feed_forward_network',"
        " '/root/reference/network/feed_forward_network.py')\n"
        "mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)\n"
        "import dorknet_b200.layers.dense_layer as d\n"
        "assert mod.DenseLayer is d.DenseLayer\n"
        "net = mod.FeedForwardNetwork('x'); net.add_layer(d.DenseLayer('fc', 4, 2)); print(repr(net)); print('ok')\n"
        % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
