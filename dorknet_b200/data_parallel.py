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

Mode "p2p" (default on B200s of one box; DK_DP_MODE=nccl selects the above): no collective launch at all.  The flat
gradient buffer of every rank is cudaMalloc'ed, exported with cudaIpc and mapped by all peers (P2PExchange); the
optimiser kernel itself reads element i from every rank's buffer over NVLink, adds them in rank order and applies the
update (csrc/dp_p2p.cu: dk_opt_multi_p2p), bracketed by two flag handshakes on peer-mapped words.  torch.distributed
only carries the 64-byte handles at start-up and the initial parameter broadcast.
"""
import ctypes
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


def backward_order(network):
    """Parameter layers in the order their backward() RUNS (and so writes gradients): top-level layers last to first; inside
    a ResidualBlock layer_list[-1] ... layer_list[1], THEN the skip projection, then layer_list[0]
    (layers/residual_block.py backward: the skip path is taken before the first branch layer so that its gradient can be
    folded into that layer's kernel; a one-layer branch runs before the skip).  A bucket may only be exchanged once the LAST
    layer in this order that writes into it has been enqueued."""
    out = []
    for l in reversed(list(network.layers)):
        if hasattr(l, "layer_list"):
            ll = [m for m in l.layer_list]
            skip = getattr(l, "skip_projection", None)
            seq = list(reversed(ll[1:])) + ([skip] if skip is not None else []) + ll[:1] if len(ll) > 1 else ll + (
                [skip] if skip is not None else [])
            out.extend(m for m in seq if getattr(m, "learned_params", None))
        elif getattr(l, "learned_params", None):
            out.append(l)
    return out


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


class _RawCudaArray:
    """a cudaMalloc'ed range presented through __cuda_array_interface__ (torch.as_tensor wraps it without a copy)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class P2PExchange:
    """Peer-mapped gradient buffers + flag blocks of all ranks of one box, and the device-side dk_p2p_ctx."""

    FLAG_BYTES = 256  # ready[8], done[8], reduced[8] (uint32, indexed by writer rank) at bytes 0 / 32 / 64, epoch at 128

    def __init__(self, dist, group, nfloats):
        import torch
        from ._lib import api, P2PCtx
        self.api = api
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("p2p data parallel supports up to 8 ranks")
        self.nfloats = int(nfloats)

        def alloc(nbytes):
            ptr, handle = ctypes.c_void_p(0), (ctypes.c_ubyte * 64)()
            api.dk_p2p_alloc(int(nbytes), ctypes.byref(ptr), handle)
            return int(ptr.value), bytes(handle)
        # Every rank takes part in every collective below whatever happened locally, and the outcome is agreed on
        # (a rank that fell back to NCCL alone would leave the others spinning on its flags).
        err = None
        try:
            self.grad_ptr, gh = alloc((max(self.nfloats, 1) + 3) // 4 * 16)
            self.flag_ptr, fh = alloc(self.FLAG_BYTES)
        except Exception as e:  # noqa: BLE001
            err, gh, fh = str(e), None, None
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (err, gh, fh), group=group)
        bad = [g[0] for g in gathered if g[0] is not None]
        self._opened = []
        grad_ptrs, flag_ptrs = [0] * self.world, [0] * self.world
        if not bad:
            try:
                for p, (_, pgh, pfh) in enumerate(gathered):
                    if p == self.rank:
                        grad_ptrs[p], flag_ptrs[p] = self.grad_ptr, self.flag_ptr
                        continue
                    for h, out in ((pgh, grad_ptrs), (pfh, flag_ptrs)):
                        q = ctypes.c_void_p(0)
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                        api.dk_p2p_open(buf, ctypes.byref(q))
                        out[p] = int(q.value)
                        self._opened.append(int(q.value))
            except Exception as e:  # noqa: BLE001
                err = str(e)
        ok = torch.tensor([0 if (bad or err) else 1], dtype=torch.int32, device=runtime.device())
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError("peer mapping failed on at least one rank: %s" % (bad[0] if bad else err))
        ctx = P2PCtx()
        ctx.world, ctx.rank = self.world, self.rank
        for p in range(self.world):
            ctx.grad_delta[p] = grad_ptrs[p] - self.grad_ptr
            ctx.ready[p] = flag_ptrs[p]
            ctx.done[p] = flag_ptrs[p] + 32
            ctx.reduced[p] = flag_ptrs[p] + 64
        ctx.epoch = self.flag_ptr + 128
        ctx.grad_base = self.grad_ptr
        ctx.nfloats = (max(self.nfloats, 1) + 3) // 4 * 4
        per = (ctx.nfloats + self.world - 1) // self.world
        ctx.slice = (per + 3) // 4 * 4
        raw = np.frombuffer(ctypes.string_at(ctypes.addressof(ctx), ctypes.sizeof(ctx)), dtype=np.uint8).copy()
        self.ctx = torch.from_numpy(raw).to(runtime.device())
        self.flat = torch.as_tensor(_RawCudaArray(self.grad_ptr, max(self.nfloats, 1)), device=runtime.device())
        dist.barrier(group=group)  # everybody has mapped everybody before the first handshake

    @property
    def ctx_ptr(self):
        return self.ctx.data_ptr()

    def wait_done(self):
        self.api.dk_p2p_wait_done(self.ctx_ptr, runtime.stream())


class DataParallel:
    def __init__(self, network, optimiser=None, num_buckets=3, overlap=True, process_group=None, device=None,
                 mode=None):
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
        # parameter layers in the order backward produces their gradients (skip projections where they really run)
        assert {id(l) for l in backward_order(network)} == {id(l) for l in iter_param_layers(network, include_skip=True)}
        self.entries = []  # (layer, key)
        for layer in backward_order(network):
            if hasattr(layer, "_ensure_gpu") and device is None:
                layer._ensure_gpu()
            for k in layer.learned_params.keys():
                if hasattr(layer, "_param") and device is None:
                    layer._param(k)
                self.entries.append((layer, k))
        sizes = [int(np.prod(l.learned_params[k].shape)) for l, k in self.entries]
        self.offsets, self.total = flat_layout(sizes)
        self.sizes = sizes
        # gradient exchange: "p2p" = fused into the optimiser kernel over NVLink peer memory, "nccl" = all_reduce
        if mode is None:
            mode = os.environ.get("DK_DP_MODE", "p2p")
        self.mode = "nccl"
        self.p2p = None
        if (mode == "p2p" and self.world > 1 and device is None and optimiser is not None and self.world <= 8
                and hasattr(optimiser, "attach_p2p")):
            try:
                self.p2p = P2PExchange(dist, process_group, self.total)
                self.mode = "p2p"
            except Exception as e:  # noqa: BLE001 -- e.g. no peer access between these GPUs: NCCL still works
                import sys
                sys.stderr.write("dorknet_b200: peer-memory gradient exchange unavailable (%s); using NCCL\n" % e)
                self.p2p = None
        self._flatten_grads()
        self.buckets = []
        for a, b in plan_buckets(sizes, num_buckets):
            lo = self.offsets[a]
            hi = self.offsets[b - 1] + sizes[b - 1]
            last_layer = self.entries[b - 1][0]
            self.buckets.append(dict(lo=lo, hi=hi, trigger=last_layer))
        if optimiser is not None:
            optimiser.grad_scale = 1.0 / self.world
            if self.p2p is not None:
                optimiser.attach_p2p(self.p2p)
        if self.world > 1 and overlap and self.p2p is None:
            self._install_hooks()

    # -- flat gradient storage ---------------------------------------------------------------------
    def _flatten_grads(self):
        import torch
        dev = self.device if self.device is not None else runtime.device()
        if self.p2p is not None:
            self.flat = self.p2p.flat  # cudaMalloc'ed, mapped by every peer
        else:
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
        seen = set()
        for layer, _ in self.entries:
            nl = getattr(layer, "non_learned_params", None)
            if nl and id(layer) not in seen:
                seen.add(id(layer))
                for k in ("running_mean", "running_std"):
                    v = dict.get(nl, k)  # (no flush of a deferred forward: plain dict access)
                    if v is not None and hasattr(v, "t"):
                        self.dist.broadcast(v.t, src=src, group=self.group)

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

    def begin_step(self):
        """Kept for callers of the first release: the "every peer has read my gradients" wait is now enqueued by the
        optimiser right behind the exchange kernel (optimisers/_multi.py), so a plain loop `forward; backward; dp.step()`
        is safe without it.  Calling it is harmless (the flags already carry the current epoch: no spin)."""
        if self.p2p is not None:
            self.p2p.wait_done()

    def finish(self):
        """Call after network.backward(): (issue and) wait for every bucket on the compute stream."""
        if self.p2p is not None:
            return  # the optimiser kernel reads the peers' gradients itself
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
