"""-m gpu: host-side semantics the reference's own loops rely on (round-1 advisor findings)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tiny_net():
    from dorknet_b200 import workloads
    M = workloads.ours()
    np.random.seed(3)
    net = M.FeedForwardNetwork("t")
    net.add_layer(M.PointwiseConvLayer("pw", filter_block_shape=(8, 4), with_bias=False, weight_regulariser=M.l2(1e-2)))
    net.add_layer(M.BatchNormLayer("bn", input_dimension=4, incoming_chans=8))
    net.add_layer(M.ReLu("relu"))
    net.add_layer(M.GlobalAveragePoolingLayer("gap"))
    net.add_layer(M.DenseLayer("fc", incoming_chans=8, output_dim=3, weight_regulariser=M.l2(1e-2)))
    net.set_loss_layer(M.SoftmaxWithCrossEntropy("sm"))
    return M, net


def _batch(seed):
    g = np.random.default_rng(seed)
    return g.standard_normal((6, 4, 8, 8)).astype(np.float32), np.eye(3, dtype=np.float32)[g.integers(0, 3, 6)]


@pytest.mark.parametrize("graphed", [False, True])
def test_a_held_loss_keeps_the_value_of_its_step(graphed):
    """examples/imagenet_dogs_225_resnet_18_depsep.py:222-226 holds `loss` across steps (running average): the value of
    step i must not change when step i+1 rewrites the device slots, and the running average must not grow a term list."""
    from dorknet_b200.graph import GraphedTrainStep
    M, net = _tiny_net()
    opt = M.SGDMomentum(net, 0.1, 0.9)
    step = GraphedTrainStep(net, opt, warmup=1, enabled=graphed)
    X0, Y0 = _batch(0)
    held, vals = [], []
    running = None
    for i in range(6):
        loss = step(X0, Y0)  # same batch every step: the loss falls as the weights train
        held.append(loss)
        vals.append(float(loss))
        running = loss if running is None else 0.9 * running + 0.1 * loss
        assert len(running.terms) <= len(loss.terms) + 1
    assert vals[0] != vals[-1]
    assert [float(h) for h in held] == vals  # every held loss still reads ITS step's value
    expect = vals[0]
    for v in vals[1:]:
        expect = 0.9 * expect + 0.1 * v
    assert float(running) == pytest.approx(expect, rel=1e-6)


def test_consumed_batchnorm_output_is_still_readable():
    """BN -> ResidualBlock([ReLu, ...], identity skip) (pre-activation layout): the ReLu fuses with the deferred BatchNorm,
    and the block's identity skip must still read the plain BatchNorm output."""
    from dorknet_b200 import workloads
    M = workloads.ours()
    g = np.random.default_rng(5)
    X = g.standard_normal((4, 8, 6, 6)).astype(np.float32)
    bn = M.BatchNormLayer("bn", input_dimension=4, incoming_chans=8)
    pw = M.PointwiseConvLayer("pw", filter_block_shape=(8, 8), with_bias=False)
    blk = M.ResidualBlock("res", layer_list=[M.ReLu("r0"), pw], skip_projection=None, post_skip_activation=M.ReLu("r1"))
    y_bn = bn.forward(X)
    out = blk.forward(y_bn).get()
    mu, var = X.mean((0, 2, 3), keepdims=True), X.var((0, 2, 3), keepdims=True)
    yb = (X - mu) / np.sqrt(var + 1e-5)
    W = np.asarray(pw.learned_params["weights"].get() if hasattr(pw.learned_params["weights"], "get") else pw.learned_params["weights"])
    branch = np.einsum("nchw,fc->nfhw", np.maximum(yb, 0), W)
    expect = np.maximum(branch + yb, 0)
    assert np.max(np.abs(out - expect)) <= 2e-3 * np.max(np.abs(expect))
    assert np.max(np.abs(y_bn.get() - yb)) <= 1e-5 * np.max(np.abs(yb))
    # and backward runs through the fused pair
    dx = bn.backward(blk.backward(np.ones_like(out)))
    assert dx.shape == X.shape and np.isfinite(dx.get()).all()
