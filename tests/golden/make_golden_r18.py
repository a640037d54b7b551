#!/usr/bin/env python3
"""Golden vectors of ONE training step of the reference's real ResNet-18-depsep (cfg3, BASELINE.json) at 225x225,
batch 8, produced by the LIVE, UNMODIFIED reference CPU path (oracle/_ref, built from /root/reference).

    python tests/golden/make_golden_r18.py          # -> tests/golden/r18_b8.npz  (~6 MB: every gradient)
    python tests/golden/make_golden_r18.py --mnist  # -> tests/golden/mnist_b16.npz (cfg1 network, one step)

Network: dorknet_b200.workloads.build_resnet18_depsep built on the REFERENCE's classes (the layer list of
examples/imagenet_dogs_225_resnet_18_depsep.py:32-160), weights drawn by the reference initialisers from
np.random.seed(0) -- the test rebuilds the same net on our classes with the same seed (same draw order) and checks
the per-tensor checksums stored here before comparing anything.  Input: workloads.synthetic_batch(8, 3, 225, 120,
seed=7) (uint8 images - 128).  Stored: loss, scores, every parameter gradient, BatchNorm running statistics after the
step, test-mode scores after one SGDMomentum update, checksums of the initial weights -- all from the reference --
and, next to them, the SAME step evaluated in float64 (tests/golden/fp64_net.py): loss64, scores64, grad64/*, and
referr/* = the reference's own normalised max-abs error against that exact evaluation.

Why the float64 evaluation is stored.  The reference's fp32 gradients are NOT accurate at these sizes: against the
exact (float64) gradients of the very same weights and inputs its own results are off by up to 9.5 % (res7_pw_skip,
res7_dw2_pw), 2-5 % on most of res3-res8, 2e-3 / 6e-3 on the MNIST net's conv1 / conv2, while its forward pass
agrees to 2e-7.  The loss is in the per-channel fp32 reductions whose terms cancel (dbeta = sum(dY) and mean(dY) of
BatchNormLayer.dx, layers/batch_norm.py:127,171): on the MNIST net bn2/beta is off by 3.5e-3 where bn2/gamma of the same
layer is exact to 4e-6, and everything upstream of that layer inherits the error.  A GPU gradient can therefore only
be compared with the reference's up to the reference's own error, and is held instead to
the exact gradient, at least as tightly as the reference meets it (tests/net_parity.py).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fp64_net  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _workloads():
    """dorknet_b200/workloads.py loaded as a plain file: it only needs numpy, and importing the package would bind
    the product's `layers` names next to the reference's."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dk_workloads", os.path.join(ROOT, "dorknet_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def run(net, W, R, X, Y, lr, out):
    for l in W.iter_param_layers(net):
        for k, v in l.learned_params.items():
            out["initsum/%s/%s" % (l.layer_name, k)] = np.float64(np.sum(np.abs(np.asarray(v, np.float64))))
    t0 = time.time()
    ev = fp64_net.Evaluator(net)
    loss64, p64 = ev.forward(X, Y)
    g64 = ev.backward()
    print("float64 evaluation: %.1f s" % (time.time() - t0))
    out["loss64"] = np.float64(loss64)
    out["scores64"] = f32(p64)
    for name, (mu, std) in ev.stats.items():
        out["mean64/%s" % name] = f32(mu)
        out["std64/%s" % name] = f32(std)
    opt = R.SGDMomentum(net, lr, 0.9)
    t0 = time.time()
    loss, scores = net.forward(X, Y)
    net.backward()
    print("reference forward+backward: %.1f s" % (time.time() - t0))
    out["loss"] = np.float64(loss)
    out["scores"] = f32(scores)
    for l in W.iter_param_layers(net):
        for k, v in l.grads.items():
            out["grad/%s/%s" % (l.layer_name, k)] = f32(v).copy()
            exact = g64[(l.layer_name, k)].reshape(np.shape(v))
            out["grad64/%s/%s" % (l.layer_name, k)] = f32(exact)
            out["referr/%s/%s" % (l.layer_name, k)] = np.float64(
                np.max(np.abs(np.asarray(v, np.float64) - exact)) / max(float(np.max(np.abs(exact))), 1e-30))
        nl = getattr(l, "non_learned_params", None)
        if nl and nl.get("running_mean") is not None:
            out["rm/%s" % l.layer_name] = f32(nl["running_mean"]).reshape(-1)
            out["rs/%s" % l.layer_name] = f32(nl["running_std"]).reshape(-1)
    opt.update_weights()
    _, st = net.forward(X, None, test_mode=True)
    out["scores_test"] = f32(st)


def main():
    R = load_reference()
    W = _workloads()
    out = {}
    if "--mnist" in sys.argv:
        net = W.build_mnist_convnet(R, seed=0)
        g = np.random.default_rng(11)
        X = f32(g.uniform(0, 1, (16, 1, 28, 28)))
        y = g.integers(0, 10, 16)
        Y = np.eye(10, dtype=np.float32)[y]
        out["X"], out["Y"] = X, Y
        run(net, W, R, X, Y, 0.01, out)
        name = "mnist_b16"
    else:
        net = W.build_resnet18_depsep(R, classes=120, conv0_padding=1, seed=0)
        X, _, Y = W.synthetic_batch(8, 3, 225, 120, seed=7)
        run(net, W, R, X, Y, 0.05 * 8 / 200.0, out)
        name = "r18_b8"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "%.1f MB" % (os.path.getsize(os.path.join(OUT, name + ".npz")) / 1e6))


if __name__ == "__main__":
    main()
