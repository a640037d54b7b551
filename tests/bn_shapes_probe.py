"""Diagnostic (B200): fused vs split BatchNorm kernels at the ResNet-18-depsep layer shapes (forward + backward, with and
without the fused ReLU) -- relative differences of y, dx, dgamma, dbeta."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dorknet_b200 import api
    from dorknet_b200.layers.batch_norm import BatchNormLayer
    from dorknet_b200.layers.activations import ReLu
    rng = np.random.default_rng(1)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for (C, H) in ((64, 112), (64, 56), (128, 28), (256, 14), (512, 7)):
        X = (rng.standard_normal((N, C, H, H)) * 0.1 + 0.05).astype(np.float32)
        dY = (rng.standard_normal((N, C, H, H)) * 1e-3).astype(np.float32)
        for relu in (0, 1):
            out = {}
            for fused in (0, 1):
                api.dk_tc_debug_set(9, fused)
                bn = BatchNormLayer("bn", input_dimension=4, incoming_chans=C)
                act = ReLu("r")
                y = bn.forward(X)
                if relu:
                    y = act.forward(y)
                yh = y.get().copy()
                g = act.backward(dY) if relu else dY
                dx = bn.backward(g).get().copy()
                out[fused] = (yh, dx, bn.grads["gamma"].get().copy(), bn.grads["beta"].get().copy())
            rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))  # noqa: E731
            print("N=%d C=%d H=%d relu=%d: y %.2e  dx %.2e  dgamma %.2e  dbeta %.2e" % (
                N, C, H, relu, rel(out[1][0], out[0][0]), rel(out[1][1], out[0][1]), rel(out[1][2], out[0][2]),
                rel(out[1][3], out[0][3])), flush=True)
    api.dk_tc_debug_set(9, 1)


if __name__ == "__main__":
    main()
