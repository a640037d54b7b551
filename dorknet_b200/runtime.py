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


# ---- a second stream for work that is off the backward critical path ------------------------------------------------
# Backward is one dependent chain of small kernels (dgrad -> BatchNorm backward -> depthwise backward -> ...), but the
# weight gradient of a pointwise layer hangs off that chain: it reads dY and the saved input, and nothing needs dW before
# the optimiser.  Inside side_region() -- GraphedTrainStep / AutoGraph wrap network.backward() in it, so the fork and the
# join are part of the captured CUDA graph -- such launches go to a side stream (own scratch buffer) and fill the SMs the
# chain's partial waves and kernel tails leave idle; leaving the region joins the side stream back.
# MEASURED (B200, ResNet-18-depsep batch 64, 13 pointwise wgrads = 0.44 ms of eager kernel time moved off the chain):
# 3.318 ms per step with it, 3.307 ms without -- the chain's kernels and the wgrad GEMMs want the same SMs (200 KB of
# shared memory per GEMM CTA) and simply interleave.  Off by default (DK_ASYNC_WGRAD=1 turns it on); parity-green either way.
_side = {"stream": None, "enabled": False, "dirty": False, "scratch": None}


class side_region:
    def __enter__(self):
        self._was = _side["enabled"]
        _side["enabled"] = os.environ.get("DK_ASYNC_WGRAD", "0") == "1"
        return self

    def __exit__(self, *exc):
        side_join()
        _side["enabled"] = self._was
        return False


def side_enabled():
    return _side["enabled"]


def side_scratch(nbytes):
    """scratch of the side stream (never shared with the main stream's)"""
    torch = _torch()
    cur = _side["scratch"]
    if cur is None or cur.numel() < nbytes:
        size = (max(int(nbytes), 1 << 20) + (1 << 20) - 1) & ~((1 << 20) - 1)
        _side["scratch"] = cur = torch.empty(size, dtype=torch.uint8, device=device())
    return cur.data_ptr(), cur.numel()


def side_launch(fn):
    """Run fn() with the side stream current, ordered behind everything enqueued on the current stream so far."""
    torch = _torch()
    cur = torch.cuda.current_stream()
    if _side["stream"] is None:
        _side["stream"] = torch.cuda.Stream()
    s = _side["stream"]
    s.wait_stream(cur)
    with torch.cuda.stream(s):
        fn()
    _side["dirty"] = True


def side_join():
    """The current stream waits for everything launched through side_launch()."""
    if _side["dirty"]:
        _torch().cuda.current_stream().wait_stream(_side["stream"])
        _side["dirty"] = False
