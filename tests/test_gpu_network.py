"""-m gpu: network-level parity at the BASELINE configurations against golden vectors of the live reference
(tests/golden/make_golden_r18.py), on the product default backend -- no relaxed "chaotic" band -- and the same step
driven by the REFERENCE'S OWN UNCHANGED container.  The checks live in tests/net_parity.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("backend,fold", [(0, None), (0, False), (1, False), (1, None)])
def test_resnet18_depsep_225_batch8_training_step(backend, fold):
    """cfg3's network (examples/imagenet_dogs_225_resnet_18_depsep.py:32-160) at 225x225, batch 8: loss, scores, all
    137 parameter gradients, BatchNorm running statistics, test-mode scores after the update.  fold None = the product
    default (depthwise BatchNorms folded into the pointwise GEMMs where their statistics ride on the depthwise kernel),
    False = every BatchNorm as its own layer, the reference's operation order (the tight fp32 gates apply to that one)."""
    import net_parity
    net_parity.run("r18", "ours", backend, fold=fold)


@pytest.mark.parametrize("backend", [0, 1])
def test_mnist_convnet_training_step(backend):
    """cfg1's network (examples/MNIST_basic_convnet.py:15-69), batch 16: 3x3 and 4x4-stride-2 ConvLayers."""
    import net_parity
    net_parity.run("mnist", "ours", backend)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "network")),
                    reason="oracle/_ref (the compiled reference) did not travel")
@pytest.mark.parametrize("net", ["r18", "mnist"])
def test_reference_container_unchanged_drives_cuda_layers(net):
    """a16: network/feed_forward_network.py of the reference, compiled unmodified into oracle/_ref, runs forward /
    backward / test-mode forward on top of the CUDA layers (dorknet_b200.dropin) and meets the same gates.  Own
    process: the reference's module names (`layers`, `network`, ...) are bound to the product there."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "net_parity.py"), "--net", net, "--container", "ref"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "net_parity ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "network.feed_forward_network" in out.stdout
    if net == "mnist":  # dropin.accelerate: CUDA-graph replay behind the reference container's unchanged methods
        assert "autograph ok" in out.stdout and "(ref container)" in out.stdout, out.stdout[-2000:]


def test_p2p_gradient_exchange_two_gpus():
    """dk_opt_multi_p2p (gradient exchange fused into the optimiser kernel over NVLink peer memory) against NCCL's
    all-reduce on 2 GPUs: tests/p2p_check.py under torchrun.  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(ROOT, "tests", "p2p_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("which", ["mini", "mnist"])
def test_batchnorm_folded_inference_matches_test_mode(which):
    """fold_batchnorm(net) scores like net.forward(test_mode=True) (batch_norm.py:101-115 as an affine
    map absorbed by the preceding layer's weights); every linear -> BatchNorm pair disappears; terminal_layer_name
    early exit (feed_forward_network.py:55-56) still works."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import dorknet_b200.workloads as wl
    from dorknet_b200.inference import fold_batchnorm
    from net_defs import build_small_net
    M = wl.ours()
    rng = np.random.default_rng(11)
    if which == "mini":
        net, shape, classes = build_small_net(M, seed=5), (6, 3, 33, 33), 5
    else:
        net, shape, classes = wl.build_mnist_convnet(M, seed=5), (16, 1, 28, 28), 10
    X = rng.standard_normal(shape).astype(np.float32)
    Y = np.zeros((shape[0], classes), np.float32)
    Y[np.arange(shape[0]), rng.integers(0, classes, shape[0])] = 1
    opt = M.SGDMomentum(net, 0.05, 0.9)
    for _ in range(3):  # running statistics and gamma / beta away from their initial values
        net.forward(X, Y)
        net.backward()
        opt.update_weights()
    Xt = rng.standard_normal(shape).astype(np.float32)
    _, want = net.forward(Xt, None, test_mode=True)
    want = want.get()
    folded = fold_batchnorm(net)
    n_bn = sum(1 for l in wl.iter_param_layers(net) if type(l).__name__ == "BatchNormLayer")
    assert len(folded.folded_batchnorms) == n_bn and n_bn > 0
    assert not any(type(l).__name__ == "BatchNormLayer" for l in wl.iter_param_layers(folded))
    _, got = folded.forward(Xt)
    got = got.get()
    assert np.abs(got - want).max() <= 2e-3 * np.abs(want).max() + 1e-6
    assert (np.argmax(got, 1) == np.argmax(want, 1)).mean() >= 0.9
    # early exit at a surviving layer: same activations as the unfolded network one BatchNorm later
    first = net.layers[0].layer_name
    _, a = folded.forward(Xt, terminal_layer_name=first)
    _, b = net.forward(Xt, None, test_mode=True, terminal_layer_name=net.layers[1].layer_name)
    a, b = a.get(), b.get()
    assert a.shape == b.shape and np.abs(a - b).max() <= 2e-3 * np.abs(b).max() + 1e-6
    with pytest.raises(KeyError):
        folded.forward(Xt, terminal_layer_name=net.layers[1].layer_name)
    with pytest.raises(ValueError):
        folded.forward(Xt, Y, test_mode=False)


def test_autograph_behind_an_unchanged_loop_is_bit_identical():
    """dropin.accelerate (graph.AutoGraph): forward / backward / update_weights replayed from CUDA graphs after two eager
    steps, a test-mode forward in between and a smaller last batch: bit-identical to the eager twin."""
    import dorknet_b200.workloads as wl
    import net_parity
    assert net_parity.autograph_check(wl.ours()) >= 3
