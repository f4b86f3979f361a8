"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: tuple sharding, the flat-gradient allreduce (one blocking
call, or three slices reduced from the backward-part hook) and the packed reduce of the criterion scalars."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _FakeNet(object):
    def __init__(self, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self._flat = torch.randn(1024 + 64, generator=g)
        self._flat[1024:] = 0
        self._bufs = torch.randn(128, generator=g)
        self.g = torch.randn(1024 + 64, generator=g)
        self.g[1024:] = 0
        p = torch.nn.Parameter(self._flat[:16])
        p.grad = self.g[:16]                     # .grad aliases the flat gradient buffer, as in PoseNet
        self._param_list = [p]
        self._grad_part_hook = None

    def flat_parameters(self):
        return self._flat, self.g


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geomapnet_b200.ddp import shard_tuples, allreduce_flat_, FlatDataParallel
    x = torch.arange(8 * 3 * 2).view(8, 3, 2)
    sh = shard_tuples(x, rank, world)
    ok = sh.shape[0] == 4 and torch.equal(sh, x[rank * 4:(rank + 1) * 4])       # T never split
    net = _FakeNet(rank)
    crit = torch.nn.Module()
    crit.sax = torch.nn.Parameter(torch.tensor([float(rank)])); crit.saq = torch.nn.Parameter(torch.tensor([1.0]))
    crit.sax.grad = torch.tensor([1.0 + rank]); crit.saq.grad = torch.tensor([10.0 * (rank + 1)])
    g_before = net.g.clone()
    dp = FlatDataParallel(net, crit, overlap=(os.environ.get("TEST_DDP_OVERLAP") == "1"))
    dp.broadcast_parameters()
    if dp.overlap:
        # what PoseNet._run_backward does: three parts, back to front, each followed by the hook
        assert net._grad_part_hook is not None
        for part, (lo, hi) in enumerate([(700, 1088), (300, 700), (0, 300)]):
            net._grad_part_hook(part, net.g[lo:hi])
    scale = dp.allreduce_grads()
    gathered = [torch.zeros_like(g_before) for _ in range(world)]
    dist.all_gather(gathered, g_before)
    want = sum(gathered)
    ok = ok and abs(scale - 0.5) < 1e-12
    ok = ok and torch.allclose(net.g[:1024], want[:1024], rtol=1e-6, atol=1e-6)
    ok = ok and float(net.g[1024:].abs().max()) == 0.0                          # padding stays zero
    ok = ok and abs(float(crit.sax.grad) - 3.0) < 1e-6 and abs(float(crit.saq.grad) - 30.0) < 1e-6
    # parameters were broadcast from rank 0
    ref = _FakeNet(0)
    ok = ok and torch.equal(net._flat, ref._flat) and float(crit.sax) == 0.0
    # a .grad that does not alias the reduced buffer (gradient accumulation) is refused, not silently mis-reduced
    net._param_list[0].grad = torch.zeros(16)
    try:
        dp.allreduce_grads()
        ok = False
    except RuntimeError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("overlap", ["0", "1"])
def test_flat_allreduce_world2_gloo(overlap, monkeypatch):
    monkeypatch.setenv("TEST_DDP_OVERLAP", overlap)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
