"""torchrun diagnostic: NCCL all-reduce latency/bandwidth at the gradient-buffer size, plus the transport NCCL picked."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    for n in (1508344, 16 * 1024 * 1024, 256 * 1024 * 1024):
        t = torch.ones(n, dtype=torch.float32, device="cuda")
        for _ in range(5):
            dist.all_reduce(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if rank == 0:
            print("all_reduce %9d floats: %.3f ms  (%.1f GB/s algbw)" % (n, ms, 4 * n / ms / 1e6), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
