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
    before the first capture (they are real training steps).  With a DataParallel wrapper the gradient all-reduce
    runs between two graphs (forward+backward | optimiser)."""

    def __init__(self, network, optimiser, data_parallel=None, warmup=1, enabled=True):
        self.net = network
        self.opt = optimiser
        self.dp = data_parallel
        self.warmup = warmup
        self.enabled = enabled
        self._graphs = {}
        self._seen = 0
        self._static = None
        self._opt_graph = None

    def _eager(self, X, Y):
        loss, _ = self.net.forward(X, Y)
        self.net.backward()
        if self.dp is not None:
            self.dp.step()
        else:
            self.opt.update_weights()
        return loss

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
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, _ = self.net.forward(X, Y)
                self.net.backward()
                if self.dp is None:
                    self.opt.update_weights()
            entry = (g, loss, X, Y)
            self._graphs[key] = entry
            if self.dp is not None and self._opt_graph is None:
                og = torch.cuda.CUDAGraph()
                with torch.cuda.graph(og):
                    self.opt.update_weights()
                self._opt_graph = og
        self.opt.push_hyper()
        entry[0].replay()
        if self.dp is not None:
            self.dp.finish()
            self._opt_graph.replay()
        return entry[1]

    @property
    def num_graphs(self):
        return len(self._graphs)
