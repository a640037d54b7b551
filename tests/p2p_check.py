"""torchrun check (2+ B200s of one box): the peer-memory gradient exchange fused into the optimiser kernel
(dk_opt_multi_p2p) against NCCL -- (w_old - w_new) / lr must equal the all-reduced mean gradient, and the replicas
must stay bit-identical.   python -m torch.distributed.run --nproc-per-node 2 tests/p2p_check.py"""
import os
import sys

import numpy as np

LR = 0.02
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from dorknet_b200 import runtime, workloads as W
    from dorknet_b200.data_parallel import DataParallel, init_process_group
    from dorknet_b200.graph import GraphedTrainStep
    rank, world = init_process_group()
    runtime.ensure_init()
    M = W.ours()
    net = W.build_resnet18_depsep(M, classes=10, conv0_padding=1, seed=0)
    opt = M.SGDMomentum(net, LR, 0.0)  # momentum 0: w += -lr * mean gradient
    net.to_gpu()
    dp = DataParallel(net, opt)
    assert dp.mode == "p2p", "peer-memory mode did not come up: " + dp.mode
    dp.broadcast_parameters(0)
    step = GraphedTrainStep(net, opt, dp, warmup=1)
    worst = 0.0
    for it in range(4):  # 1 eager step, then captured graphs
        X, _, Y = W.synthetic_batch(8, 3, 65, 10, seed=100 * rank + it)
        before = [l.learned_params[k].t.clone() for l, k in dp.entries]
        step(X, Y)
        torch.cuda.synchronize()
        g = dp.flat.clone()  # this rank's gradients of the step (the optimiser does not modify them)
        dist.all_reduce(g)
        g /= world
        for (l, k), b, off, n in zip(dp.entries, before, dp.offsets, dp.sizes):
            if l not in opt.learnable_layers:
                continue
            ref = g[off:off + n]
            want = b.reshape(-1) - LR * ref  # w - lr * mean gradient
            got = l.learned_params[k].t.reshape(-1)
            # (NCCL adds the ranks in another order than rank 0, 1, ...: rounding of the sum, relative to its size)
            err = float((got - want).abs().max() / (b.abs().max() + LR * ref.abs().max() + 1e-12))
            worst = max(worst, err)
        dist.barrier()
    chk = torch.stack([l.learned_params[k].t.double().sum() for l, k in dp.entries]).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(float(c) == float(allc[0]) for c in allc)
    if rank == 0:
        print("p2p_check: world %d, worst |w_new - (w - lr * nccl mean gradient)| / max|w| = %.3e, replicas identical: %s, graphs %d"
              % (world, worst, same, step.num_graphs), flush=True)
    assert worst < 1e-5 and same
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
