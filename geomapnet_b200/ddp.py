"""Plain data parallel over the GPUs of one box: one process per GPU, ONE
allreduce of the flat gradient buffer per step (NCCL over NVLink 5 / NVSwitch).

The reference is single-GPU (common/train.py:193-196); this is the multi-GPU
row of SURVEY.md section 8e.  Tuples (the N dimension of [N,T,3,H,W]) are sharded
across ranks, never the T frames of a tuple: the relative-pose loss only
couples frames inside a tuple (common/criterion.py:94-105).  BatchNorm
statistics stay per-rank (no SyncBN), as when running the reference at the
local batch size.  Equal local batches => mean of local L1 means == global L1
mean, so averaged gradients equal the single-process gradient.

All functions work on any device/backend (gloo on CPU in the tests).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_tuples", "allreduce_flat_", "FlatDataParallel"]


def shard_tuples(x, rank, world):
    """Rank r takes tuples [r*n, (r+1)*n) of the leading dimension (must divide evenly)."""
    N = x.shape[0]
    if N % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (N, world))
    n = N // world
    return x[rank * n:(rank + 1) * n]


def allreduce_flat_(flat_grad, extra_grads=(), group=None, average=False):
    """ONE sum-allreduce over ``flat_grad`` with the (tiny) ``extra_grads`` tensors
    (criterion scalars sax/saq/srx/srq) riding in its last padding slots.
    In-place; returns the world size.  With average=False the caller folds 1/world
    into the optimizer step (FusedAdam grad_scale)."""
    world = dist.get_world_size(group)
    extras = [g for g in extra_grads if g is not None]
    k = sum(g.numel() for g in extras)
    if k > 0:
        tail = flat_grad[flat_grad.numel() - k:]
        saved = tail.clone()
        off = 0
        for g in extras:
            tail[off:off + g.numel()] = g.reshape(-1).to(tail.dtype)
            off += g.numel()
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    if k > 0:
        off = 0
        for g in extras:
            g.copy_(tail[off:off + g.numel()].view_as(g))
            off += g.numel()
        tail.copy_(saved * world)   # padding slots: restore what the sum would have held
    if average:
        flat_grad.mul_(1.0 / world)
        for g in extras:
            g.mul_(1.0 / world)
    return world


class FlatDataParallel(object):
    """Wraps a geomapnet_b200 PoseNet/MapNet (+ criterion) for data-parallel steps.

        dp = FlatDataParallel(model, criterion)       # broadcasts rank 0's weights
        loss = criterion(model(x_local), targ_local); loss.backward()
        scale = dp.allreduce_grads()                  # one NCCL allreduce; returns 1/world
        optimizer.learner.step(grad_scale=scale)
    """

    def __init__(self, model, criterion=None, group=None, broadcast=True):
        self.model = model
        self.posenet = model.mapnet if hasattr(model, "mapnet") else model
        self.criterion = criterion
        self.group = group
        self.world = dist.get_world_size(group)
        self._broadcast_pending = broadcast

    def broadcast_parameters(self):
        flat, _ = self.posenet.flat_parameters()
        dist.broadcast(flat, src=0, group=self.group)
        dist.broadcast(self.posenet._bufs, src=0, group=self.group)
        if self.criterion is not None:
            for p in self.criterion.parameters():
                dist.broadcast(p.data, src=0, group=self.group)
        self._broadcast_pending = False

    def allreduce_grads(self):
        if self._broadcast_pending:
            raise RuntimeError("call broadcast_parameters() once after the model is on its device")
        _, gflat = self.posenet.flat_parameters()
        if gflat is None:
            raise RuntimeError("no gradient buffer: run backward() first")
        extras = []
        if self.criterion is not None:
            extras = [p.grad for p in self.criterion.parameters() if p.requires_grad and p.grad is not None]
        allreduce_flat_(gflat, extras, group=self.group, average=False)
        return 1.0 / self.world
