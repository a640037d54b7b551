#!/usr/bin/env python3
"""Golden vectors of ONE training step of the reference's real ResNet-18-depsep (cfg3, BASELINE.json) at 225x225,
batch 8, produced by the LIVE, UNMODIFIED reference CPU path (oracle/_ref, built from /root/reference).

    python tests/golden/make_golden_r18.py          # -> tests/golden/r18_b8.npz  (~6 MB: every gradient)
    python tests/golden/make_golden_r18.py --mnist  # -> tests/golden/mnist_b16.npz (cfg1 network, one step)

Network: dorknet_b200.workloads.build_resnet18_depsep built on the REFERENCE's classes (the layer list of
examples/imagenet_dogs_225_resnet_18_depsep.py:32-160), weights drawn by the reference initialisers from
np.random.seed(0) -- the test rebuilds the same net on our classes with the same seed (same draw order) and checks
the per-tensor checksums stored here before comparing anything.  Input: workloads.synthetic_batch(8, 3, 225, 120,
seed=7) (uint8 images - 128).  Stored: loss, l2-free data loss, scores, every parameter gradient, BatchNorm running
statistics after the step, test-mode scores after one SGDMomentum update, checksums of the initial weights.
"""
import os
import sys
import time

import numpy as np

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
