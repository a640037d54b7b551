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
    from dorknet_b200.array import asarray
    step = GraphedTrainStep(net, opt, dp, warmup=1)
    worst = 0.0
    nf = int(dp.p2p.nfloats)
    nf4 = (nf + 3) // 4 * 4
    per = (nf4 + world - 1) // world
    sl = (per + 3) // 4 * 4  # floats per rank slice (data_parallel.P2PExchange)

    def check(before, mean_grad):
        w = 0.0
        for (l, k), b, off, n in zip(dp.entries, before, dp.offsets, dp.sizes):
            if l not in opt.learnable_layers:
                continue
            ref = mean_grad[off:off + n]
            want = b.reshape(-1) - LR * ref  # w - lr * mean gradient
            got = l.learned_params[k].t.reshape(-1)
            # (NCCL adds the ranks in another order than rank 0, 1, ...: rounding of the sum, relative to its size)
            w = max(w, float((got - want).abs().max() / (b.abs().max() + LR * ref.abs().max() + 1e-12)))
        return w

    for it in range(5):
        X, _, Y = W.synthetic_batch(8, 3, 65, 10, seed=100 * rank + it)
        before = [l.learned_params[k].t.clone() for l, k in dp.entries]
        if it < 2:
            # phases apart: the gradients of this rank are read BEFORE the exchange (the reduce-scatter kernel writes the
            # sum over all ranks into slice `rank` of this buffer, in place), NCCL reduces the copy
            net.forward(asarray(X), asarray(Y))
            net.backward()
            torch.cuda.synchronize()
            g = dp.flat.clone()
            dist.all_reduce(g)
            g /= world
            dp.step()
            torch.cuda.synchronize()
            worst = max(worst, check(before, g))
            # and the in-place reduce-scatter itself: my slice now holds world * the NCCL mean
            lo, hi = rank * sl, min((rank + 1) * sl, nf)
            if hi > lo:
                mine = dp.flat[lo:hi]
                e = float((mine - world * g[lo:hi]).abs().max() / (world * g[lo:hi].abs().max() + 1e-12))
                worst = max(worst, e)
        else:
            # the whole step as one CUDA graph (handshakes, reduce-scatter, update all captured): the update must equal
            # -lr * (gathered reduced slices) / world
            step(X, Y)
            torch.cuda.synchronize()
            pad = torch.zeros(world * sl, device=dp.flat.device)
            lo, hi = rank * sl, min((rank + 1) * sl, nf)
            mine = torch.zeros(sl, device=dp.flat.device)
            if hi > lo:
                mine[:hi - lo] = dp.flat[lo:hi]
            dist.all_gather_into_tensor(pad, mine)
            worst = max(worst, check(before, pad[:nf] / world))
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
