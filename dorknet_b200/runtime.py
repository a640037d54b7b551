"""Per-process device context: one process drives one B200 (one rank per GPU under torchrun).

torch is used for plumbing only: the caching allocator, the current CUDA stream handle, and
torch.distributed/NCCL for the data-parallel gradient exchange.
"""
import os

_state = {"init": False, "device": None, "scratch": None}


def _torch():
    import torch
    return torch


def device():
    ensure_init()
    return _state["device"]


def ensure_init():
    """Bind to the local GPU and initialise the native library.  Fails loudly without a B200."""
    if _state["init"]:
        return
    torch = _torch()
    from ._lib import api
    if not torch.cuda.is_available():
        raise RuntimeError("dorknet_b200 needs a CUDA device (sm_100a): there is no CPU fallback on this path")
    idx = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
    if idx >= torch.cuda.device_count():
        idx = torch.cuda.current_device()
    torch.cuda.set_device(idx)
    api.dk_init(idx)
    _state["device"] = torch.device("cuda", idx)
    _state["init"] = True


def stream():
    """cudaStream_t (as int) of torch's current stream: every kernel is enqueued there, so CUDA-graph
    capture, events and NCCL ordering all see the same stream."""
    return _torch().cuda.current_stream().cuda_stream


def synchronize():
    _torch().cuda.synchronize()


def scratch(nbytes):
    """Shared GEMM scratch (split-K partials, permuted filters): grows on demand, never zeroed.
    Stream-ordered reuse is safe because all layers run on one stream.  Returns (ptr, nbytes)."""
    torch = _torch()
    ensure_init()
    cur = _state["scratch"]
    if cur is None or cur.numel() < nbytes:
        size = max(int(nbytes), 1 << 20)
        size = (size + (1 << 20) - 1) & ~((1 << 20) - 1)
        _state["scratch"] = torch.empty(size, dtype=torch.uint8, device=_state["device"])
        cur = _state["scratch"]
    return cur.data_ptr(), cur.numel()


def zeroed_workspace(nbytes):
    """A private, zero-initialised workspace (arrival counters must start at 0; kernels reset them)."""
    torch = _torch()
    ensure_init()
    return torch.zeros(max(int(nbytes), 4), dtype=torch.uint8, device=_state["device"])
