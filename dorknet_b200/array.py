"""DeviceArray: the array type the layers exchange on the B200 path.

It plays the role CuPy arrays play on the reference's GPU path (the reference's layers only rely on
`.shape`, `+`, and passing arrays along; the example loops use `cp.asarray` / `cp.asnumpy`).
Storage is a torch CUDA tensor used purely as an allocation + stream handle: no torch op touches
the data on the hot path -- every computation goes through the C ABI on `.ptr`.
"""
import numpy as np

from . import runtime

_NP2TORCH = None


def _torch():
    import torch
    return torch


def _torch_dtype(dtype):
    global _NP2TORCH
    torch = _torch()
    if _NP2TORCH is None:
        _NP2TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32,
                     np.dtype(np.int64): torch.int64, np.dtype(np.uint8): torch.uint8}
    return _NP2TORCH[np.dtype(dtype)]


class DeviceArray:
    """A contiguous device buffer with a shape.  float32 unless stated."""

    __slots__ = ("t", "shape", "dtype")
    __array_priority__ = 100.0

    def __init__(self, t, shape=None, dtype=np.float32):
        self.t = t
        self.shape = tuple(int(s) for s in (shape if shape is not None else t.shape))
        self.dtype = np.dtype(dtype)

    # -- metadata ------------------------------------------------------------------------
    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return "DeviceArray(shape=%s, dtype=%s)" % (self.shape, self.dtype)

    # -- host <-> device -------------------------------------------------------------------
    def get(self):
        """Synchronous device -> host copy as a NumPy array."""
        return self.t.detach().reshape(self.shape).cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a.astype(dtype) if dtype is not None else a

    def set(self, host):
        """Host -> device copy into this buffer (shape must match)."""
        torch = _torch()
        h = np.ascontiguousarray(host, dtype=self.dtype)
        if h.size != self.size:
            raise ValueError("DeviceArray.set: size mismatch %s vs %s" % (h.shape, self.shape))
        self.t.reshape(-1).copy_(torch.from_numpy(h.reshape(-1)), non_blocking=False)
        return self

    def copy_from(self, other):
        """Device -> device copy (cudaMemcpyAsync on the current stream)."""
        self.t.reshape(-1).copy_(other.t.reshape(-1), non_blocking=True)
        return self

    def copy(self):
        out = empty(self.shape, self.dtype)
        return out.copy_from(self)

    # -- views -----------------------------------------------------------------------------
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            known = 1
            for s in shape:
                if s != -1:
                    known *= s
            shape[shape.index(-1)] = self.size // known
        n = 1
        for s in shape:
            n *= s
        if n != self.size:
            raise ValueError("cannot reshape %s into %s" % (self.shape, tuple(shape)))
        return DeviceArray(self.t, shape, self.dtype)

    # -- the one arithmetic op the reference's layers use on activations (residual join) -----
    def __add__(self, other):
        from ._lib import api
        other = asarray(other)
        if other.shape != self.shape:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (self.shape, other.shape))
        out = empty(self.shape)
        api.dk_add(self.ptr, other.ptr, out.ptr, self.size, runtime.stream())
        return out

    __radd__ = __add__


class LazyDeviceArray(DeviceArray):
    """A DeviceArray whose contents are produced on first use.

    ConvLayer.backward returns its input gradient this way: the reference's container computes the first
    layer's dX only to drop it (feed_forward_network.py:69-70), so the dgrad kernel -- the most expensive
    launch of conv0 -- runs only if somebody actually reads the result (`.ptr`, `.t`, `.get()`, `+`).  Like
    every buffer a layer hands out, it must be consumed before that layer's next backward()."""

    __slots__ = ("_thunk", "_buf")

    def __init__(self, buf, thunk):
        self._buf = buf
        self._thunk = thunk
        self.shape = buf.shape
        self.dtype = buf.dtype

    def materialise(self):
        if self._thunk is not None:
            thunk, self._thunk = self._thunk, None
            thunk()
        return self._buf

    @property
    def t(self):
        return self.materialise().t

    @t.setter
    def t(self, value):  # DeviceArray.__init__ is bypassed; nothing assigns t
        raise AttributeError("LazyDeviceArray.t is read-only")

    @property
    def is_materialised(self):
        return self._thunk is None


class LazyBNOutput(LazyDeviceArray):
    """Output of a training-mode BatchNormLayer.forward whose normalisation pass has not run yet: the statistics
    kernel has, and `scale` / `shift` per channel are on the device.  A ReLu that receives it launches ONE fused
    apply+ReLU pass instead of two (and tells the BatchNorm to mask its backward); any other consumer materialises
    the plain y = x*scale + shift on first use."""

    __slots__ = ("bn", "_consumed")

    def __init__(self, buf, thunk, bn):
        super().__init__(buf, thunk)
        self.bn = bn
        self._consumed = False

    @property
    def fusable(self):
        """True while nothing has been launched for this output and no fused consumer has taken it."""
        return self._thunk is not None and not self._consumed

    def consume(self):
        """A fused kernel (BatchNorm+ReLU, BatchNorm+join) took this output and writes its own buffer.  The plain
        y = x*scale + shift is still owed to any OTHER reader -- an identity skip of a ResidualBlock whose branch starts
        with a ReLu, a user calling .get() -- as the reference returns a real array there: on first read the statistics
        pass runs if it has not (flushing the fused consumer into its buffer), then y is applied from the saved values."""
        self._consumed = True
        bn, buf = self.bn, self._buf

        def late_reader():
            bn._flush()
            bn.apply_saved(buf, 0)
        self._thunk = late_reader


class LazyReluOutput(LazyDeviceArray):
    """Output of a ReLu that follows a training-mode BatchNorm whose normalisation pass is still deferred: nothing has
    been launched yet.  Any consumer that reads it gets relu(batchnorm(x)) from ONE fused pass, as before; a
    PointwiseConvLayer with stride s > 1 instead asks the BatchNorm for just the pixels it reads
    (`bn.fused_relu_apply_strided`) and re-points the thunk to a plain apply from the saved statistics, so the
    full-size activation is only ever produced if somebody else wants it."""

    __slots__ = ("bn", "relu")

    def __init__(self, buf, thunk, bn, relu):
        super().__init__(buf, thunk)
        self.bn = bn
        self.relu = relu

    def replace_thunk(self, thunk):
        self._thunk = thunk


class LazyDWOutput(LazyDeviceArray):
    """Output of a training-mode DepthwiseConvLayer.forward that has not been launched yet.  Any reader gets the plain
    depthwise forward on first use; a BatchNormLayer that is being folded into the pointwise GEMM behind it instead asks the
    layer for the variant that also accumulates the BatchNorm statistics while the outputs are in registers
    (dk_dwconv_fwd_bn), so the statistics cost no pass over the activation."""

    __slots__ = ("layer",)

    def __init__(self, buf, thunk, layer):
        super().__init__(buf, thunk)
        self.layer = layer

    def launched_by_consumer(self):
        self._thunk = None


class ZeroSumGrad(DeviceArray):
    """The input gradient a BatchNormLayer.backward returns: per channel it sums to zero over (N, H, W) -- exactly, in real
    arithmetic, because sum(x_hat) = 0 (batch_norm.py:125-156).  A layer that needs the channel sums of its upstream
    gradient (a bias gradient, the folded BatchNorm of bn_fold.cu) may skip that reduction when it receives one."""

    __slots__ = ()


class FoldedBNGrad(LazyDeviceArray):
    """What PointwiseConvLayer.backward returns when the BatchNorm in front of it was folded into its GEMMs (bn_fold.cu):
    the BatchNorm's own backward has already happened inside the dgrad epilogue -- `dx` is the gradient with respect to the
    BatchNorm's INPUT and bn.grads are written.  The BatchNorm's backward() recognises the object and hands `dx` on.  Anyone
    else reading it (`.ptr`, `.get()`) gets what the reference's pointwise layer returns, the gradient with respect to the
    BatchNorm's output, computed on first use by the plain dgrad GEMM."""

    __slots__ = ("bn", "dx")

    def __init__(self, buf, thunk, bn, dx):
        super().__init__(buf, thunk)
        self.bn = bn
        self.dx = dx


class LazyStridedGrad(LazyDeviceArray):
    """Input gradient of a stride-s PointwiseConvLayer: zero except at [:, :, ::s, ::s] (pointwise_convolution.py:68-72).
    Reading it (`.ptr`, `.get()`, `+`) produces the reference's zero-stuffed full-size tensor.  A consumer that can use
    the non-zero entries alone calls compact() instead and gets them as a dense [N, C, OH, OW] array -- the dgrad GEMM
    then writes a quarter of the bytes (stride 2) and the zeros are never stored or read."""

    __slots__ = ("stride", "_compact_fn", "_compact")

    def __init__(self, buf, thunk, stride, compact_fn):
        super().__init__(buf, thunk)
        self.stride = stride
        self._compact_fn = compact_fn
        self._compact = None

    def compact(self):
        if self._compact is None:
            self._compact = self._compact_fn()
        return self._compact


def empty(shape, dtype=np.float32):
    torch = _torch()
    runtime.ensure_init()
    shape = tuple(int(s) for s in shape)
    n = 1
    for s in shape:
        n *= s
    t = torch.empty(max(n, 1), dtype=_torch_dtype(dtype), device=runtime.device())
    return DeviceArray(t, shape, dtype)


def zeros(shape, dtype=np.float32):
    a = empty(shape, dtype)
    a.t.zero_()
    return a


def asarray(a, dtype=np.float32):
    """Host array -> DeviceArray (synchronous H2D copy); DeviceArray passes through."""
    if isinstance(a, DeviceArray):
        return a
    torch = _torch()
    runtime.ensure_init()
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            a = a.to(runtime.device())
        return DeviceArray(a.contiguous(), a.shape, np.float32 if a.dtype == torch.float32 else np.int32)
    h = np.ascontiguousarray(a, dtype=dtype)
    t = torch.from_numpy(h.reshape(-1) if h.ndim else h.reshape(1)).to(runtime.device())
    return DeviceArray(t, h.shape, dtype)


def asnumpy(a):
    if isinstance(a, DeviceArray):
        return a.get()
    return np.asarray(a)


class _Epoch:
    """The lifetime of one set of values in the slot arena: from the first l2 / loss kernel of a step until the next
    step overwrites them.  sealed = a snapshot of the arena (pinned host copy + event) was enqueued right after the
    step's loss kernel; DeviceScalars of a sealed epoch read THAT copy, so they keep their value for ever."""

    __slots__ = ("seq", "host", "ev")
    _count = 0

    def __init__(self):
        _Epoch._count += 1
        self.seq = _Epoch._count
        self.host = None
        self.ev = None

    @property
    def sealed(self):
        return self.host is not None

    def values(self):
        self.ev.synchronize()
        return self.host.numpy()

    def __del__(self):
        try:
            if self.host is not None:
                _SlotArena._free_hosts.append(self.host)
        except Exception:  # interpreter shutdown
            pass


class _SlotArena:
    """One small device buffer holding every one-float result slot (loss, l2 terms), so reading a
    DeviceScalar is ONE device->host copy however many terms it has."""

    SIZE = 4096
    _free_hosts = []

    def __init__(self):
        self.buf = None
        self.used = 0
        self.host = None
        self.epoch = _Epoch()

    def alloc(self):
        torch = _torch()
        runtime.ensure_init()
        if self.buf is None:
            self.buf = torch.zeros(self.SIZE, dtype=torch.float32, device=runtime.device())
            self.host = torch.zeros(self.SIZE, dtype=torch.float32).pin_memory()
        if self.used >= self.SIZE:
            raise RuntimeError("scalar slot arena exhausted")
        i = self.used
        self.used += 1
        a = DeviceArray(self.buf[i:i + 1], (1,))
        return a, i

    def read(self):
        """Synchronous snapshot of all slots (one D2H copy on the current stream)."""
        torch = _torch()
        self.host[:max(self.used, 1)].copy_(self.buf[:max(self.used, 1)], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.host.numpy()

    def read_async(self):
        """Enqueue a snapshot of all slots into its own pinned buffer on the current stream; returns (host tensor,
        event).  The host can keep launching: it reads the snapshot once the event has fired."""
        torch = _torch()
        n = max(self.used, 1)
        host = self._free_hosts.pop() if self._free_hosts else None
        if host is None or host.numel() < n:
            host = torch.empty(max(n, 256), dtype=torch.float32).pin_memory()
        host[:n].copy_(self.buf[:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev

    def seal(self):
        """End the current epoch: snapshot the arena behind everything enqueued so far and open a new epoch (called by
        the loss layer after its kernel, and by GraphedTrainStep after a replay).  Returns the sealed epoch.  Not legal
        while the stream is being captured (events recorded there never fire): callers check."""
        ep = self.epoch
        if self.buf is not None:
            ep.host, ep.ev = self.read_async()
            self.epoch = _Epoch()
        return ep


_arena = _SlotArena()


class PendingScalar:
    """Result of DeviceScalar.fetch_async(): the device -> host copy is in flight; result() waits for it only."""

    __slots__ = ("_scalar",)

    def __init__(self, scalar):
        self._scalar = scalar

    def result(self):
        return float(self._scalar)


def alloc_scalar_slot():
    """(DeviceArray one-float slot, arena index)"""
    return _arena.alloc()


def current_epoch():
    return _arena.epoch


def seal_epoch():
    return _arena.seal()


class DeviceScalar:
    """A float produced on the device: sum_i coeff_i * slot_i + const.

    The loss (losses.py:23-27) and every layer's l2 term (regularisers/l2.py:12-14) are produced by kernels into
    one-float device slots; `network.forward` adds them up with Python `+` (feed_forward_network.py:59-60), which
    here only concatenates term lists -- no kernel, no sync.

    VALUE SEMANTICS.  The slots are rewritten by every step, so a scalar belongs to an epoch of the arena (_Epoch).
    The loss layer seals the epoch right after its kernel (an asynchronous 16 KB device -> host snapshot + event), so
    the loss a training loop holds keeps the value of ITS step: float(x) waits for that snapshot only (never for later
    steps), and arithmetic between scalars of different steps -- the reference loop's
    `running = 0.9*running + 0.1*loss` (examples/imagenet_dogs_225_resnet_18_depsep.py:222-226) -- folds the older
    one into the constant, so term lists do not grow.  Scalars of a still-open epoch (a regulariser term read on its
    own, anything created under CUDA-graph capture) read the device when converted.
    """

    __slots__ = ("terms", "const", "epoch")
    __array_priority__ = 100.0

    def __init__(self, terms=(), const=0.0, epoch=None):
        self.terms = list(terms)  # [(slot, coeff)]: slot = arena index (int) or any object with .get()
        self.const = float(const)
        self.epoch = epoch if epoch is not None else (_arena.epoch if any(isinstance(s, int) for s, _ in self.terms) else None)

    def _resolved(self):
        """(const, non-arena terms) with every arena term of a SEALED epoch folded into the constant."""
        vals = self.epoch.values()
        c = np.float64(self.const)
        rest = []
        for slot, coeff in self.terms:
            if isinstance(slot, int):
                c += coeff * float(vals[slot])
            else:
                rest.append((slot, coeff))
        return float(c), rest

    def _combine(self, other, sign=1.0):
        if not isinstance(other, DeviceScalar):
            return DeviceScalar(self.terms, self.const + sign * float(other), self.epoch)
        a_terms, a_const, a_ep = self.terms, self.const, self.epoch
        b_terms, b_const, b_ep = [(s, sign * c) for s, c in other.terms], sign * other.const, other.epoch
        if a_ep is not None and b_ep is not None and a_ep is not b_ep:
            # two different steps: the older (sealed) one becomes a number
            if a_ep.sealed and (not b_ep.sealed or a_ep.seq < b_ep.seq):
                a_const, a_terms = self._resolved()
                a_ep = None
            elif b_ep.sealed:
                c, rest = other._resolved()
                b_const, b_terms, b_ep = sign * c, [(s, sign * k) for s, k in rest], None
            else:
                raise RuntimeError("DeviceScalars of two open epochs cannot be combined")
        return DeviceScalar(a_terms + b_terms, a_const + b_const, a_ep if a_ep is not None else b_ep)

    def __add__(self, other):
        return self._combine(other)

    __radd__ = __add__

    def __sub__(self, other):
        return self._combine(other, -1.0)

    def __rsub__(self, other):
        return (self * -1.0)._combine(other)

    def __mul__(self, k):
        k = float(k)
        return DeviceScalar([(s, c * k) for s, c in self.terms], self.const * k, self.epoch)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self * (1.0 / float(k))

    def rebind(self, epoch):
        """The same expression over another epoch's values (GraphedTrainStep: one captured loss, one epoch per replay)."""
        return DeviceScalar(self.terms, self.const, epoch)

    def __float__(self):
        total = np.float64(self.const)
        snap = None
        if self.epoch is not None and self.epoch.sealed:
            snap = self.epoch.values()
        else:
            from .regularisers.l2 import flush_pending
            flush_pending()
        for slot, coeff in self.terms:
            if isinstance(slot, int):
                if snap is None:
                    snap = _arena.read()
                total += coeff * float(snap[slot])
            else:
                total += coeff * float(slot.get().reshape(-1)[0])
        return float(total)

    def fetch_async(self):
        """Start the device -> host read of this value behind whatever is enqueued on the current stream and return a
        PendingScalar; the caller keeps launching work and calls .result() later (a training loop reads the loss of
        step i while step i+1 runs, instead of draining the GPU every step).  A scalar of a sealed epoch already has
        its snapshot in flight: nothing more is enqueued."""
        if self.epoch is not None and not self.epoch.sealed and self.epoch is _arena.epoch:
            import torch
            if not torch.cuda.is_current_stream_capturing():
                from .regularisers.l2 import flush_pending
                flush_pending()
                return PendingScalar(self.rebind(_arena.seal()))
        return PendingScalar(self)

    def get(self):
        return np.float32(float(self))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype or np.float32)

    def __repr__(self):
        return "DeviceScalar(%d terms)" % len(self.terms)

    def __format__(self, spec):
        return format(float(self), spec)

    # comparisons / printing in user loops (`if loss < best:`) resolve to the value
    def __lt__(self, o):
        return float(self) < float(o)

    def __gt__(self, o):
        return float(self) > float(o)

    def __le__(self, o):
        return float(self) <= float(o)

    def __ge__(self, o):
        return float(self) >= float(o)
