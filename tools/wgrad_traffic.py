"""One eager dk_pwconv_wgrad call per pointwise shape of ResNet-18-depsep (batch 64), for an `ncu --set full` capture of
the family's DRAM traffic:
    ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|splitk_reduce|repack_planes" \
        -o gpurun_out/r02_wgrad_full python tools/wgrad_traffic.py
Prints the call list (shape, multiplicity per training step, algorithmic bytes) so that the kernels of the report can be
attributed to calls in launch order."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# C, F, H (input), stride, calls per step
SHAPES = [(64, 64, 56, 1, 4), (128, 128, 28, 1, 3), (256, 256, 14, 1, 3), (512, 512, 7, 1, 3), (64, 64, 112, 2, 1),
          (64, 128, 28, 1, 1), (128, 256, 14, 1, 1), (256, 512, 7, 1, 1), (64, 128, 56, 2, 1), (128, 256, 28, 2, 1),
          (256, 512, 14, 2, 1)]


def main():
    import torch
    from dorknet_b200 import api, runtime
    from dorknet_b200.array import asarray, empty
    runtime.ensure_init()
    N = 64
    rng = np.random.default_rng(0)
    calls = []
    for (C, F, H, s, mult) in SHAPES:
        OH = (H - 1) // s + 1
        x = asarray(rng.standard_normal((N, C, H, H)).astype(np.float32))
        dy = asarray(rng.standard_normal((N, F, OH, OH)).astype(np.float32))
        w = asarray((rng.standard_normal((F, C)) / 8).astype(np.float32))
        dw = empty((F, C))
        ws, wsn = runtime.scratch(max(api.dk_pwconv_ws_bytes(N, C, H, H, F, s), 1 << 20))
        # flush L2 with a 256 MB write so that the capture sees cold operands like a training step does
        flush = torch.empty(64 << 20, dtype=torch.float32, device=runtime.device())
        flush.fill_(1.0)
        torch.cuda.synchronize()
        k0 = api.kernel_launches() if hasattr(api, "kernel_launches") else 0
        api.dk_pwconv_wgrad(dy.ptr, x.ptr, w.ptr, dw.ptr, None, 1e-4, N, C, H, H, F, s, ws, wsn, runtime.stream())
        torch.cuda.synchronize()
        calls.append({"shape": "N=%d C=%d F=%d H=%d stride=%d" % (N, C, F, H, s), "calls_per_step": mult,
                      "algorithmic_bytes": 4 * N * OH * OH * (C + F)})
        del flush
    print(json.dumps(calls))


if __name__ == "__main__":
    main()
