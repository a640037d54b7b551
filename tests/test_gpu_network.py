"""-m gpu: network-level parity at the BASELINE configurations against golden vectors of the live reference
(tests/golden/make_golden_r18.py), on the product default backend -- no relaxed "chaotic" band -- and the same step
driven by the REFERENCE'S OWN UNCHANGED container.  The checks live in tests/net_parity.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("backend", [0, 1])
def test_resnet18_depsep_225_batch8_training_step(backend):
    """cfg3's network (examples/imagenet_dogs_225_resnet_18_depsep.py:32-160) at 225x225, batch 8: loss, scores, all
    137 parameter gradients, BatchNorm running statistics, test-mode scores after the update."""
    import net_parity
    net_parity.run("r18", "ours", backend)


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
