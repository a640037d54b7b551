"""How long does a CUDA graph of N dependent tiny kernels take on this GPU?  (per-node launch gap: the part of the 217-launch
training step that no kernel optimisation removes).  Measurement tool for DESIGN.md."""
import torch
x = torch.zeros(32, device="cuda")
for n in (1, 50, 217, 434):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            x.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                x.add_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    print("graph of %3d tiny kernels: %.1f us per replay, %.2f us per node" % (n, us, us / n))
