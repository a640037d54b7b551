"""cuBLAS TF32 / bf16 matmul throughput on this box (a denominator for the TF32 tensor-pipe numbers in DESIGN.md; not a product path)."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True
for dt, n in ((torch.float32, 8192), (torch.bfloat16, 8192)):
    a = torch.randn(n, n, device="cuda", dtype=dt)
    b = torch.randn(n, n, device="cuda", dtype=dt)
    for _ in range(3):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s %d^3: %.3f ms  %.1f TFLOP/s" % (dt, n, ms, 2 * n ** 3 / ms / 1e9))
