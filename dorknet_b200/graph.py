"""Whole-step CUDA-graph capture.

A ResNet-18-depsep step is ~330 kernel launches of 5-100 us each; driven from Python through ctypes the host
needs longer to issue them than the B200 needs to run them.  GraphedTrainStep runs the (unchanged) layer code
once under stream capture -- forward, backward, optimiser update -- and afterwards replays the captured graph:
one launch per step.  This works because every layer keeps persistent output / gradient / cache buffers, the
workspaces have reached their final size after the warm-up step, scalar results (loss, l2 terms) live in a
device arena, and the optimiser reads its hyper-parameters from device memory.
"""
from . import runtime
from .array import DeviceArray, asarray


class GraphedTrainStep:
    """step = GraphedTrainStep(net, optimiser[, data_parallel]);  loss = step(X, Y)

    X / Y must be DeviceArrays that stay alive (static input buffers, e.g. the slots of a HostBatchUploader):
    one graph is captured per distinct (X, Y) buffer pair and replayed whenever that pair comes back.  Host
    arrays are accepted too: they are uploaded into an internal static pair first.  `warmup` eager steps run
    before the first capture (they are real training steps).  With a DataParallel wrapper the bucketed NCCL
    all-reduces are captured INSIDE the graph (issued from the wrapped layer.backward on NCCL's stream, forked
    from and joined back into the capture stream, so they overlap the rest of backward and a step stays ONE
    host launch); `dp_in_graph=False` (or DK_DP_IN_GRAPH=0, or a failed capture) falls back to one all-reduce
    between two graphs (forward+backward | optimiser)."""

    def __init__(self, network, optimiser, data_parallel=None, warmup=1, enabled=True, dp_in_graph=None):
        import os
        self.net = network
        self.opt = optimiser
        self.dp = data_parallel
        self.warmup = warmup
        self.enabled = enabled
        if dp_in_graph is None:
            dp_in_graph = os.environ.get("DK_DP_IN_GRAPH", "1") != "0"
        self.dp_in_graph = bool(dp_in_graph) and data_parallel is not None
        self._graphs = {}
        self._seen = 0
        self._static = None
        self._opt_graph = None

    def _eager(self, X, Y):
        if self.dp is not None:
            self.dp.begin_step()
        loss, _ = self.net.forward(X, Y)
        self._backward()
        if self.dp is not None:
            self.dp.step()
        else:
            self.opt.update_weights()
        return loss

    def _backward(self):
        """network.backward() with the weight gradients that are off the critical path on a side stream
        (runtime.side_region; not with NCCL buckets, whose all-reduces are issued from inside backward)"""
        if self.dp is not None and getattr(self.dp, "mode", "") == "nccl":
            return self.net.backward()
        with runtime.side_region():
            self.net.backward()

    def _inputs(self, X, Y):
        if isinstance(X, DeviceArray) and isinstance(Y, DeviceArray):
            return X, Y
        if self._static is None:
            self._static = (asarray(X), asarray(Y))
        else:
            self._static[0].set(X)
            self._static[1].set(Y)
        return self._static

    def __call__(self, X, Y):
        import torch
        X, Y = self._inputs(X, Y)
        if not self.enabled or self._seen < self.warmup:
            self._seen += 1
            return self._eager(X, Y)
        key = (X.ptr, Y.ptr, X.shape, Y.shape)
        entry = self._graphs.get(key)
        if entry is None:
            self.opt.push_hyper()
            torch.cuda.synchronize()
            entry = None
            if self.dp_in_graph:
                try:
                    entry = self._capture_whole(X, Y)
                except Exception as e:  # NCCL would not be captured on this stack: keep training, between graphs
                    import sys
                    sys.stderr.write("dorknet_b200: all-reduce inside the CUDA graph failed (%s); falling back\n" % e)
                    self.dp_in_graph = False
                    self.dp.drop_pending()
                    torch.cuda.synchronize()
            if entry is None:
                if self.dp is not None:
                    self.dp.hooks_enabled = False  # fallback: one all-reduce after the graph, not from the hooks
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    if self.dp is not None:
                        self.dp.begin_step()
                    loss, _ = self.net.forward(X, Y)
                    self._backward()
                    if self.dp is None:
                        self.opt.update_weights()
                entry = (g, loss, X, Y)
                if self.dp is not None and self._opt_graph is None:
                    og = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(og):
                        self.opt.update_weights()
                    self._opt_graph = og
            self._graphs[key] = entry
        self.opt.push_hyper()
        entry[0].replay()
        if self.dp is not None and not self.dp_in_graph:
            self.dp.finish()
            self._opt_graph.replay()
        # the captured loss names arena slots that every replay rewrites: hand out THIS step's values (an asynchronous
        # 16 KB snapshot right behind the replay; array.DeviceScalar), so a held loss keeps its value
        from .array import DeviceScalar, seal_epoch
        loss = entry[1]
        return loss.rebind(seal_epoch()) if isinstance(loss, DeviceScalar) else loss

    def _capture_whole(self, X, Y):
        """forward + backward (+ bucketed all-reduces on NCCL's stream) + optimiser in ONE graph"""
        import torch
        g = torch.cuda.CUDAGraph()
        self.dp.hooks_enabled = True
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.dp.begin_step()
            loss, _ = self.net.forward(X, Y)
            self._backward()
            self.dp.finish()  # joins NCCL's stream back into the capture stream (nothing to do in p2p mode)
            self.opt.update_weights()
        return (g, loss, X, Y)

    @property
    def num_graphs(self):
        return len(self._graphs)


class AutoGraph:
    """CUDA-graph replay behind an UNCHANGED training loop.

    The reference's loops (examples/imagenet_dogs_225_resnet_18_depsep.py:216-229) call `network.forward(X, y)`,
    `network.backward()` and `optimiser.update_weights()` one after the other from Python; driven that way the host needs
    ~11 ms to issue the ~230 kernels a B200 runs in 3.4 ms.  `AutoGraph(network, optimiser)` (or
    `dorknet_b200.dropin.accelerate(network, optimiser)`) replaces those three bound methods on the INSTANCES -- the
    container class, the reference's own included, stays untouched: after `warmup` eager calls with the same input shapes
    each of them is captured once into a CUDA graph and replayed from then on (three launches per step).  Inputs of any kind
    (host arrays, fresh device arrays) are copied into static device buffers first; a new input shape, test mode or
    `terminal_layer_name` gets its own graphs (the last, smaller batch of an epoch; `network.test`).  The returned loss is
    a DeviceScalar of the step's own epoch, the scores are the loss layer's persistent buffer -- as in eager mode.
    """

    def __init__(self, network, optimiser=None, warmup=2):
        self.net, self.opt, self.warmup = network, optimiser, max(int(warmup), 1)
        self._fwd = network.forward
        self._bwd = network.backward
        self._upd = optimiser.update_weights if optimiser is not None else None
        self._state = {}      # key -> dict(n, X, Y, fwd=(graph, loss, scores), bwd=graph)
        self._last = None     # key of the last training forward (what backward() belongs to)
        self._upd_graph = None
        self._upd_calls = 0
        network.forward = self.forward
        network.backward = self.backward
        if optimiser is not None:
            optimiser.update_weights = self.update_weights

    def remove(self):
        """Give the instances their own methods back."""
        for obj, name in ((self.net, "forward"), (self.net, "backward"), (self.opt, "update_weights")):
            if obj is not None and name in vars(obj):
                delattr(obj, name)

    @staticmethod
    def _shape(a):
        return None if a is None else tuple(int(s) for s in a.shape)

    def _stage(self, st, X, Y):
        """copy the call's inputs into this key's static device buffers"""
        import numpy as np
        for name, src in (("X", X), ("Y", Y)):
            if src is None:
                continue
            if st[name] is None:
                st[name] = asarray(np.ascontiguousarray(src, np.float32)) if not isinstance(src, DeviceArray) else src.copy()
            elif isinstance(src, DeviceArray):
                if src.ptr != st[name].ptr:
                    st[name].copy_from(src)
            else:
                st[name].set(src)
        return st["X"], st["Y"]

    def forward(self, X, y_one_hot=None, test_mode=False, terminal_layer_name=None):
        import torch
        key = (self._shape(X), self._shape(y_one_hot), bool(test_mode), terminal_layer_name)
        st = self._state.setdefault(key, {"n": 0, "X": None, "Y": None, "fwd": None, "bwd": None})
        if not test_mode:
            self._last = key
        if st["n"] < self.warmup:
            st["n"] += 1
            return self._fwd(X, y_one_hot, test_mode=test_mode, terminal_layer_name=terminal_layer_name)
        Xs, Ys = self._stage(st, X, y_one_hot)
        if st["fwd"] is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, scores = self._fwd(Xs, Ys, test_mode=test_mode, terminal_layer_name=terminal_layer_name)
            st["fwd"] = (g, loss, scores)
        g, loss, scores = st["fwd"]
        g.replay()
        from .array import DeviceScalar, seal_epoch
        if isinstance(loss, DeviceScalar):
            loss = loss.rebind(seal_epoch())
        return loss, scores

    def backward(self):
        import torch
        st = self._state.get(self._last)
        if st is None or st["fwd"] is None:
            return self._bwd()  # the forward of this step ran eagerly: so does its backward
        if st["bwd"] is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                with runtime.side_region():
                    self._bwd()
            st["bwd"] = g
        st["bwd"].replay()

    def update_weights(self):
        import torch
        st = self._state.get(self._last)
        if st is None or st["bwd"] is None:
            self._upd_calls += 1
            return self._upd()
        self.opt.push_hyper()  # (learning-rate changes reach the captured kernel through device memory)
        if self._upd_graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._upd()
            self._upd_graph = g
        self._upd_graph.replay()

    @property
    def num_graphs(self):
        return sum((s["fwd"] is not None) + (s["bwd"] is not None) for s in self._state.values()) + (self._upd_graph is not None)
