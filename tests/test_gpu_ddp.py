"""Multi-GPU data parallel on real devices (NCCL): every rank's post-allreduce gradient
equals the sum of the per-rank gradients (then 1/world in FusedAdam), parameters stay
identical across ranks after the step.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, overlap):
    try:
        _worker_body(rank, world, port, q, overlap)
    except Exception:                            # a CUDA / NCCL error in one rank: report it instead of hanging the parent
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        q.put((rank, False))
        os._exit(1)


def _worker_body(rank, world, port, q, overlap):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import weights
    from helpers import make_product_model, make_product_criterion
    from geomapnet_b200.common.optimizer import Optimizer
    from geomapnet_b200.ddp import FlatDataParallel, shard_tuples
    st = weights.make_state(7 + rank)          # different initial weights: broadcast must fix that
    cfg = dict(kind="mapnet", N=4, T=3, H=64, W=64)
    x, targ = weights.make_inputs(cfg, 7)
    xs, ts = shard_tuples(x, rank, world).cuda(), shard_tuples(targ, rank, world).cuda()
    model, net = make_product_model(st, "mapnet", "fp32")
    crit = make_product_criterion("mapnet")
    opt = Optimizer([{"params": model.parameters()}, {"params": list(crit.parameters())}], "adam", 1e-4, 5e-4)
    model.train()
    model(xs)                                   # materialise the flat buffers on the device
    dp = FlatDataParallel(model, crit, overlap=overlap)
    dp.broadcast_parameters()
    flat, _ = net.flat_parameters()
    ref0 = flat.clone(); dist.broadcast(ref0, src=0)
    ok = torch.equal(flat, ref0)
    # this rank's LOCAL gradient: a backward pass with the part hook switched off (with overlap=True the hook reduces
    # the slices in place while the pass runs, so nothing local is left to look at afterwards)
    hook = net._grad_part_hook
    net._grad_part_hook = None
    loss = crit(model(xs), ts)
    opt.learner.zero_grad()
    loss.backward()
    _, g = net.flat_parameters()
    g_local = g.clone()
    sax_local = crit.sax.grad.clone()
    del loss
    net._grad_part_hook = hook
    # the step under test (same weights, same batch: the fp32 engine's split-K atomics move a gradient by ~1e-6)
    loss = crit(model(xs), ts)
    opt.learner.zero_grad()
    loss.backward()
    _, g = net.flat_parameters()
    scale = dp.allreduce_grads()
    gathered = [torch.zeros_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    want = sum(gathered)
    ok = ok and abs(scale - 1.0 / world) < 1e-12
    ok = ok and float((g - want).abs().max()) <= 1e-5 * float(want.abs().max())
    sg = [torch.zeros_like(sax_local) for _ in range(world)]
    dist.all_gather(sg, sax_local)
    ok = ok and abs(float(crit.sax.grad) - float(sum(sg))) <= 1e-5 * abs(float(sum(sg))) + 1e-7
    ok = ok and dp.overlap == overlap and (net._grad_part_hook is not None) == overlap   # slices reduced from the backward-part hook
    g_eager = g.clone()
    # the eager pass above ran on the default stream: its autograd graph (kept alive by `loss`) pins the parameters'
    # AccumulateGrad nodes to that stream, and a capture that reuses them fails with cudaErrorStreamCaptureImplicit
    # ("delete all references to the autograd graph" -- torch's own advice; geomapnet_b200/graph.py says the same)
    del loss
    import gc
    gc.collect()
    # the same step captured as CUDA graphs (overlap: slice allreduces INSIDE graph 1, geomapnet_b200/graph.py): same weights
    # (nothing has stepped yet, and building the graphed step restores the state) -> same reduced gradient
    from geomapnet_b200.graph import GraphedTrainStep
    gstep = GraphedTrainStep(model, crit, opt, xs, ts, dp=dp, warmup=2)
    gstep.x.copy_(xs); gstep.t.copy_(ts)
    gstep.g1.replay()
    dp.allreduce_grads(slices_in_graph=overlap)
    _, g2 = net.flat_parameters()
    err = float((g2 - g_eager).abs().max()) / float(g_eager.abs().max())
    ok = ok and err <= 1e-4
    ok = ok and abs(float(crit.sax.grad) - float(sum(sg))) <= 1e-5 * abs(float(sum(sg))) + 1e-7
    for _ in range(2):
        loss = gstep(xs, ts)
    ok = ok and bool(torch.isfinite(loss).all())
    after = flat.clone(); dist.broadcast(after, src=0)
    ok = ok and torch.equal(flat, after)        # replicas stay in lock-step
    if not ok:
        print("rank", rank, "graph-vs-eager reduced gradient rel err", err, flush=True)
    del gstep                                   # graphs that hold captured NCCL work go before the communicator
    import gc
    gc.collect()
    torch.cuda.synchronize()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run_world2(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    try:
        while len(res) < 2:
            res.append(q.get(timeout=180))
            if not res[-1][1]:                   # one rank failed: its peer would wait in a collective forever
                break
    finally:
        for p in procs:
            p.join(5 if (res and not res[-1][1]) else 60)
            if p.is_alive():
                p.kill()                         # exactly the processes started above
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_flat_allreduce_nccl_world2():
    """The default: one allreduce of the whole flat buffer after the backward, outside the CUDA graphs."""
    _run_world2(False)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_flat_allreduce_nccl_world2_overlap():
    """overlap=True: three slices reduced from the backward-part hook, captured inside graph 1."""
    _run_world2(True)
