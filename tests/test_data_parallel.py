"""CPU, world_size 2, gloo: the host-side logic of dorknet_b200.data_parallel -- flat gradient layout in reverse
execution order, bucket plan, all-reduce triggered from inside the wrapped layer.backward (overlap) and from
finish() (no overlap), 1/G folded into the optimiser's grad_scale."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_bucket_plan_and_layout():
    from dorknet_b200.data_parallel import flat_layout, plan_buckets
    sizes = [61440, 120, 262144, 512, 512, 4608, 512, 512, 262144, 9, 4800]
    cuts = plan_buckets(sizes, 3)
    assert cuts[0][0] == 0 and cuts[-1][1] == len(sizes)
    assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])) and 1 <= len(cuts) <= 3
    assert plan_buckets([], 3) == [] and plan_buckets([5], 4) == [(0, 1)]
    offs, total = flat_layout(sizes)
    assert all(o % 32 == 0 for o in offs) and total >= sum(sizes)
    assert all(o2 >= o1 + n for o1, o2, n in zip(offs, offs[1:], sizes))


class _P:
    """stand-in for a DeviceArray parameter: shape + a torch tensor"""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.t = torch.zeros(int(np.prod(shape)))


class _FakeLayer:
    def __init__(self, name, shapes, log):
        self.layer_name = name
        self.learned_params = {k: _P(s) for k, s in shapes.items()}
        self.grads = {k: None for k in shapes}
        self.log = log

    def backward(self, upstream):
        # write this rank's gradient into the (flat-buffer backed) grads
        for k, g in self.grads.items():
            g.t.fill_(upstream + (len(k)))
        self.log.append("bwd " + self.layer_name)
        return upstream


class _FakeNet:
    def __init__(self, layers):
        self.layers = layers


class _FakeOpt:
    grad_scale = 1.0


def _worker(rank, world, port, overlap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dorknet_b200.data_parallel import DataParallel
        log = []
        layers = [_FakeLayer("a", {"weights": (40, 3), "bias": (40,)}, log), _FakeLayer("b", {"gamma": (1, 40, 1, 1)}, log),
                  _FakeLayer("c", {"weights": (16, 40)}, log), _FakeLayer("d", {"weights": (5, 16), "bias": (5,)}, log)]
        net, opt = _FakeNet(layers), _FakeOpt()
        dp = DataParallel(net, opt, num_buckets=2, overlap=overlap, device=torch.device("cpu"))
        orig_launch = dp._launch

        def logged(b):
            log.append("allreduce [%d,%d)" % (b["lo"], b["hi"]))
            orig_launch(b)
        dp._launch = logged
        assert opt.grad_scale == 1.0 / world
        # reverse execution order: d first
        assert [l.layer_name for l, _ in dp.entries] == ["d", "d", "c", "b", "a", "a"]
        for layer in reversed(layers):
            layer.backward(float(rank + 1))
        dp.finish()
        expect = {}
        for l in layers:
            for k in l.grads:
                expect[(l.layer_name, k)] = sum(float(r + 1) + (len(k)) for r in range(world))
        ok = all(torch.all(l.grads[k].t == expect[(l.layer_name, k)]).item() for l in layers for k in l.grads)
        q.put((rank, ok, log))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_gradient_allreduce_gloo(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, log in res:
        assert ok, (rank, log)
        ar = [i for i, e in enumerate(log) if e.startswith("allreduce")]
        assert len(ar) == 2
        if overlap:
            # the first bucket goes out before the last layer's backward has even run
            assert ar[0] < log.index("bwd a")
        else:
            assert ar[0] > log.index("bwd a")


class _FakeBlock:
    """ResidualBlock stand-in with the product's real backward order: layer_list[-1..1], skip projection, layer_list[0]"""

    def __init__(self, name, layer_list, skip):
        self.layer_name = name
        self.layer_list = layer_list
        self.skip_projection = skip
        self.learned_params = None
        self.grads = None

    def backward(self, upstream):
        for l in self.layer_list[:0:-1]:
            l.backward(upstream)
        if self.skip_projection is not None:
            self.skip_projection.backward(upstream)
        self.layer_list[0].backward(upstream)
        return upstream


def _worker_block(rank, world, port, num_buckets, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dorknet_b200.data_parallel import DataParallel, backward_order
        log = []
        mk = lambda n, shp: _FakeLayer(n, {"weights": shp}, log)  # noqa: E731
        # sized so that with 2..4 buckets a cut falls between the skip projection and the branch layers around it
        blk = _FakeBlock("res", [mk("b0", (64, 9)), mk("b1", (300, 64)), mk("b2", (64, 9)), mk("b3", (300, 64))],
                         mk("skip", (700, 64)))
        layers = [mk("stem", (64, 75)), blk, mk("head", (10, 300))]
        net, opt = _FakeNet(layers), _FakeOpt()
        assert [l.layer_name for l in backward_order(net)] == ["head", "b3", "b2", "b1", "skip", "b0", "stem"]
        dp = DataParallel(net, opt, num_buckets=num_buckets, overlap=True, device=torch.device("cpu"))
        assert [l.layer_name for l, _ in dp.entries] == ["head", "b3", "b2", "b1", "skip", "b0", "stem"]
        written = set()
        orig_launch = dp._launch
        bad = []

        def logged(b):
            # every tensor inside the bucket must have been written by now
            for (l, k), off, n in zip(dp.entries, dp.offsets, dp.sizes):
                if off >= b["lo"] and off + n <= b["hi"] and ("bwd " + l.layer_name) not in log:
                    bad.append((l.layer_name, b["lo"], b["hi"]))
            orig_launch(b)
        dp._launch = logged
        for layer in reversed(layers):
            layer.backward(float(rank + 1))
        dp.finish()
        expect = sum(float(r + 1) + len("weights") for r in range(world))
        ok = all(torch.all(l.grads["weights"].t == expect).item() for l, _ in dp.entries)
        q.put((rank, ok and not bad, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_buckets", [2, 3, 4])
def test_skip_projection_gradient_is_written_before_its_bucket_is_reduced(num_buckets):
    """a block's skip projection runs AFTER layer_list[-1..1]; a bucket holding its gradient must not be
    all-reduced from an earlier layer's hook (include_skip_projections=True would train on un-reduced gradients)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_block, args=(r, 2, port, num_buckets, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, bad in res:
        assert ok, (rank, bad)
