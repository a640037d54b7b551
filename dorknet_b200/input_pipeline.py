"""Host -> device input pipeline: pinned double-buffered host staging, the
H2D copy of batch i+1 on a side stream while batch i trains, and mixup
(data_loading/image_data_loader.py:100-112: X = lam*X_b + (1-lam)*X_a, same for the one-hot labels)
done by one kernel on the device instead of on the host."""
import numpy as np

from . import runtime
from ._lib import api
from .array import DeviceArray, empty


class HostBatchUploader:
    """Double-buffered upload of (X, y_one_hot) host batches.

    submit(X, Y) copies into pinned staging memory and enqueues the H2D copy on a side stream;
    get() returns DeviceArrays once the compute stream has been made to wait for that copy.  With two
    slots the upload of the next batch overlaps the training step of the current one."""

    def __init__(self, x_shape, y_shape, slots=2):
        import torch
        runtime.ensure_init()
        self.torch = torch
        self.stream = torch.cuda.Stream(device=runtime.device())
        self.slots = []
        for _ in range(slots):
            self.slots.append(dict(
                hx=torch.empty(int(np.prod(x_shape)), dtype=torch.float32).pin_memory(),
                hy=torch.empty(int(np.prod(y_shape)), dtype=torch.float32).pin_memory(),
                dx=empty(x_shape), dy=empty(y_shape),
                ready=torch.cuda.Event(), free=torch.cuda.Event()))
        self.x_shape, self.y_shape = tuple(x_shape), tuple(y_shape)
        self.bytes_per_batch = 4 * (int(np.prod(x_shape)) + int(np.prod(y_shape)))
        self._w = 0
        self._r = 0
        self._inflight = 0

    def submit(self, X, Y):
        torch = self.torch
        s = self.slots[self._w % len(self.slots)]
        self._w += 1
        # the consumer of this slot's previous contents must be done before it is overwritten
        self.stream.wait_event(s["free"])
        s["hx"].copy_(torch.from_numpy(np.ascontiguousarray(X, np.float32).reshape(-1)))
        s["hy"].copy_(torch.from_numpy(np.ascontiguousarray(Y, np.float32).reshape(-1)))
        with torch.cuda.stream(self.stream):
            s["dx"].t.copy_(s["hx"], non_blocking=True)
            s["dy"].t.copy_(s["hy"], non_blocking=True)
            s["ready"].record(self.stream)
        self._inflight += 1

    def pin(self, X, Y):
        """Copy a host batch into freshly allocated pinned memory once (what a data loader's output buffers are);
        returns the pair to pass to submit_from_pinned()."""
        torch = self.torch
        hx = torch.empty(int(np.prod(self.x_shape)), dtype=torch.float32).pin_memory()
        hy = torch.empty(int(np.prod(self.y_shape)), dtype=torch.float32).pin_memory()
        hx.copy_(torch.from_numpy(np.ascontiguousarray(X, np.float32).reshape(-1)))
        hy.copy_(torch.from_numpy(np.ascontiguousarray(Y, np.float32).reshape(-1)))
        return hx, hy

    def submit_from_pinned(self, hx, hy):
        """Enqueue the H2D copy of an already pinned host batch into the next device slot (no host-side copy)."""
        torch = self.torch
        s = self.slots[self._w % len(self.slots)]
        self._w += 1
        self.stream.wait_event(s["free"])
        with torch.cuda.stream(self.stream):
            s["dx"].t.copy_(hx, non_blocking=True)
            s["dy"].t.copy_(hy, non_blocking=True)
            s["ready"].record(self.stream)
        self._inflight += 1

    def submit_pinned(self, slot_filler=None):
        """Like submit() when the producer already wrote the pinned staging tensors in place."""
        torch = self.torch
        s = self.slots[self._w % len(self.slots)]
        self._w += 1
        self.stream.wait_event(s["free"])
        if slot_filler is not None:
            slot_filler(s["hx"].numpy().reshape(self.x_shape), s["hy"].numpy().reshape(self.y_shape))
        with torch.cuda.stream(self.stream):
            s["dx"].t.copy_(s["hx"], non_blocking=True)
            s["dy"].t.copy_(s["hy"], non_blocking=True)
            s["ready"].record(self.stream)
        self._inflight += 1

    # ---- uint8 NHWC batches: a quarter of the PCIe bytes, converted on the device ---------------------------------
    def pin_u8(self, img, Y, img_b=None):
        """Pinned copies of decoded uint8 NHWC batch(es) and the label matrix (a loader's output buffers)."""
        torch = self.torch
        def pin(a, dt):
            t = torch.empty(a.size, dtype=dt).pin_memory()
            t.copy_(torch.from_numpy(np.ascontiguousarray(a).reshape(-1)))
            return t
        hb = pin(img_b, torch.uint8) if img_b is not None else None
        return pin(img, torch.uint8), hb, pin(np.asarray(Y, np.float32), torch.float32)

    def submit_u8_from_pinned(self, ha, hb, hy, lam=0.0, sub=128.0):
        """H2D of uint8 NHWC image batch(es) + labels on the side stream, then ONE kernel there turns them into the fp32
        NCHW network input (dk_input_u8_nhwc: transpose, - 128, mixup with `lam` when a second batch is given)."""
        torch = self.torch
        s = self.slots[self._w % len(self.slots)]
        self._w += 1
        N, C, H, W = self.x_shape
        if "ua" not in s:
            s["ua"] = torch.empty(N * H * W * C, dtype=torch.uint8, device=runtime.device())
        if hb is not None and "ub" not in s:
            s["ub"] = torch.empty(N * H * W * C, dtype=torch.uint8, device=runtime.device())
        self.stream.wait_event(s["free"])
        with torch.cuda.stream(self.stream):
            s["ua"].copy_(ha, non_blocking=True)
            if hb is not None:
                s["ub"].copy_(hb, non_blocking=True)
            s["dy"].t.copy_(hy, non_blocking=True)
            api.dk_input_u8_nhwc(s["ua"].data_ptr(), s["ub"].data_ptr() if hb is not None else None, s["dx"].ptr,
                                 float(lam), float(sub), N, C, H, W, self.stream.cuda_stream)
            s["ready"].record(self.stream)
        self._inflight += 1
        return ha.numel() + (hb.numel() if hb is not None else 0) + 4 * hy.numel()

    def get(self):
        """(X, Y) DeviceArrays of the oldest submitted batch; valid until release()."""
        if self._inflight <= 0:
            raise RuntimeError("HostBatchUploader.get() with nothing submitted")
        s = self.slots[self._r % len(self.slots)]
        self.torch.cuda.current_stream().wait_event(s["ready"])
        return s["dx"], s["dy"]

    def release(self):
        """Mark the oldest batch consumed (call after the step that used it has been enqueued)."""
        s = self.slots[self._r % len(self.slots)]
        s["free"].record(self.torch.cuda.current_stream())
        self._r += 1
        self._inflight -= 1


def mixup(Xa, Xb, lam, out=None):
    """out = lam*Xb + (1-lam)*Xa on the device (image_data_loader.py:102-110)."""
    if Xa.shape != Xb.shape:
        raise ValueError("mixup: shapes differ %s vs %s" % (Xa.shape, Xb.shape))
    if out is None:
        out = empty(Xa.shape)
    api.dk_mixup(Xa.ptr, Xb.ptr, out.ptr, float(lam), Xa.size, runtime.stream())
    return out
