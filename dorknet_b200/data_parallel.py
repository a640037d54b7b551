"""Data-parallel training across the B200s of one box: one process per GPU (torchrun), NCCL over
NVLink 5 / NVSwitch through torch.distributed.  No analogue exists in the reference;
the parity definition is §8(e): G replicas at batch b == G independent reference runs at batch b whose
gradients are averaged before the identical optimiser step.  BatchNorm statistics stay per replica.

Mechanics
  * every grads[k] of the replica is re-pointed to a view into ONE flat fp32 buffer, laid out in
    reverse execution order and cut into a few buckets;
  * FeedForwardNetwork.backward stays unchanged: each parameter layer's `backward` is wrapped so that
    when the last layer writing into a bucket has ENQUEUED its kernels, that bucket's all_reduce(SUM)
    is issued (async_op): NCCL orders it after the kernels already on the compute stream and runs it
    on its own stream, overlapped with the rest of backward;
  * the fused optimiser kernel multiplies gradients by 1/G (grad_scale), so no separate scaling pass.
"""
import os

import numpy as np

from . import runtime
from .array import DeviceArray
from .workloads import iter_param_layers


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE /
    MASTER_ADDR / MASTER_PORT).  Returns (rank, world_size); (0, 1) when not launched by torchrun."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def plan_buckets(sizes, num_buckets):
    """Cut a list of tensor sizes (already in reverse execution order) into <= num_buckets contiguous
    groups of roughly equal bytes.  Returns a list of (first_index, last_index_exclusive)."""
    total = sum(sizes)
    if not sizes:
        return []
    num_buckets = max(1, min(num_buckets, len(sizes)))
    target = total / float(num_buckets)
    cuts, acc, start = [], 0, 0
    for i, n in enumerate(sizes):
        acc += n
        remaining_buckets = num_buckets - len(cuts) - 1
        remaining_items = len(sizes) - (i + 1)
        if (acc >= target * (len(cuts) + 1) and remaining_buckets > 0 and remaining_items >= remaining_buckets):
            cuts.append((start, i + 1))
            start = i + 1
    cuts.append((start, len(sizes)))
    return cuts


def flat_layout(sizes, align=32):
    """Offsets (in floats) of tensors packed into one buffer, each aligned to `align` floats (128 B)."""
    offs, cur = [], 0
    for n in sizes:
        offs.append(cur)
        cur += (int(n) + align - 1) // align * align
    return offs, cur


class DataParallel:
    def __init__(self, network, optimiser=None, num_buckets=3, overlap=True, process_group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.network = network
        self.optimiser = optimiser
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        self._pending = []
        self.hooks_enabled = True  # GraphedTrainStep switches the backward hooks off for its between-graphs fallback
        self.device = device  # None: the process's B200; tests pass torch.device("cpu") with the gloo backend
        # parameter layers in REVERSE execution order = the order backward produces gradients
        layers = list(iter_param_layers(network, include_skip=True))
        self.entries = []  # (layer, key)
        for layer in reversed(layers):
            if hasattr(layer, "_ensure_gpu") and device is None:
                layer._ensure_gpu()
            for k in layer.learned_params.keys():
                if hasattr(layer, "_param") and device is None:
                    layer._param(k)
                self.entries.append((layer, k))
        sizes = [int(np.prod(l.learned_params[k].shape)) for l, k in self.entries]
        self.offsets, self.total = flat_layout(sizes)
        self.sizes = sizes
        self._flatten_grads()
        self.buckets = []
        for a, b in plan_buckets(sizes, num_buckets):
            lo = self.offsets[a]
            hi = self.offsets[b - 1] + sizes[b - 1]
            last_layer = self.entries[b - 1][0]
            self.buckets.append(dict(lo=lo, hi=hi, trigger=last_layer))
        if optimiser is not None:
            optimiser.grad_scale = 1.0 / self.world
        if self.world > 1 and overlap:
            self._install_hooks()

    # -- flat gradient storage ---------------------------------------------------------------------
    def _flatten_grads(self):
        import torch
        dev = self.device if self.device is not None else runtime.device()
        self.flat = torch.zeros(max(self.total, 1), dtype=torch.float32, device=dev)
        for (layer, k), off, n in zip(self.entries, self.offsets, self.sizes):
            shape = layer.learned_params[k].shape
            layer.grads[k] = DeviceArray(self.flat[off:off + n], shape)

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s weights (and BN running stats if present)."""
        if self.world <= 1:
            return
        for layer, k in self.entries:
            self.dist.broadcast(layer.learned_params[k].t, src=src, group=self.group)

    # -- overlap: all-reduce a bucket as soon as its last gradient has been enqueued -------------------
    def _install_hooks(self):
        triggers = {}
        for b in self.buckets:
            triggers.setdefault(id(b["trigger"]), (b["trigger"], []))[1].append(b)
        for layer, bs in triggers.values():
            orig = layer.backward

            def wrapped(*a, _orig=orig, _bs=bs, **kw):
                out = _orig(*a, **kw)
                if self.hooks_enabled:
                    for b in _bs:
                        self._launch(b)
                return out
            layer.backward = wrapped

    def _launch(self, b):
        if os.environ.get("DK_DP_SKIP_ALLREDUCE") == "1":  # (diagnostics knob)
            return
        work = self.dist.all_reduce(self.flat[b["lo"]:b["hi"]], op=self.dist.ReduceOp.SUM, group=self.group,
                                    async_op=True)
        self._pending.append(work)

    def finish(self):
        """Call after network.backward(): (issue and) wait for every bucket on the compute stream."""
        if self.world <= 1 or os.environ.get("DK_DP_SKIP_ALLREDUCE") == "1":  # (diagnostics knob)
            return
        if not self.overlap or not self.hooks_enabled:
            for b in self.buckets:
                self._launch(b)
        for w in self._pending:
            w.wait()  # makes the current (compute) stream wait; the host does not block
        self._pending = []

    def drop_pending(self):
        self._pending = []

    def step(self):
        self.finish()
        if self.optimiser is not None:
            self.optimiser.update_weights()
