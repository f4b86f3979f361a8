"""Optimizer wrapper with the reference's surface + a flat-buffer fused Adam.

Mirror of /root/reference/common/optimizer.py:8-47 (``Optimizer(params, method,
base_lr, weight_decay, **kwargs)``, ``.learner``, ``.adjust_lr``, ``.mult_lr``).
``method='adam'`` builds ``FusedAdam``: one CUDA launch per contiguous run of
parameters (the whole PoseNet is one run -- its parameters are views of one flat
buffer) instead of ~10 pointwise launches x 114 tensors.
``FusedAdam`` is a torch.optim.Optimizer: param_groups / state_dict() have
torch.optim.Adam's structure (step, exp_avg, exp_avg_sq per parameter) so
common/train.py:167-176,202 (checkpoint resume) keeps working.
"""
import bisect
import ctypes

import torch
import torch.optim as optim

from .. import _lib

__all__ = ["Optimizer", "FusedAdam"]

_PAD = 64  # alignment padding (floats) between parameters in the flat buffer


def _dense(p):
    """Every element of p's memory span belongs to p exactly once (elementwise kernels may sweep the span)."""
    return p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))


class FusedAdam(optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._runs = {}
        self._scratch = None

    # -- contiguous runs --------------------------------------------------------
    def _plan(self, gi, params):
        """Contiguous runs of the group's parameters (one launch each) with their moment buffers.

        The plan is keyed on the RUN STRUCTURE -- which parameters sit in which run, i.e. parameter addresses and the
        relative placement of their gradients -- not on absolute gradient addresses: a free-standing parameter (the
        criterion's sax / saq / srx / srq) gets a freshly allocated .grad from autograd every step, and the launch
        takes the gradient pointer of the day anyway.  Keyed on absolute gradient addresses the plan was rebuilt on
        every such step; harmless eagerly (moments and step are carried over) but not while a CUDA graph is being
        captured: the rebuild's allocations, zero fills, moment copies and the step-counter fill were captured with
        the step and replayed each time, resetting those parameters' Adam state on every replay
        (tests/test_gpu_graph.py: criterion scalars 0.38 lr off after 4 replays)."""
        order = sorted(params, key=lambda p: p.data_ptr())
        runs, cur = [], None
        for p in order:
            if cur is not None:
                last = cur["params"][-1]
                gap = (p.data_ptr() - (last.data_ptr() + last.numel() * 4)) // 4
                # a gap is swept by the kernels: only tolerated inside ONE storage on both sides (the flat
                # PoseNet buffers, whose padding the package zero-fills) -- never caching-allocator slack
                # between separately allocated tensors, which may hold stale NaN / Inf
                one_storage = (p.untyped_storage().data_ptr() == last.untyped_storage().data_ptr() and
                               p.grad.untyped_storage().data_ptr() == last.grad.untyped_storage().data_ptr())
                same = (p.device == last.device and (gap == 0 or (0 < gap < _PAD and one_storage)) and
                        (p.data_ptr() - last.data_ptr()) == (p.grad.data_ptr() - last.grad.data_ptr()))
                if same:
                    cur["params"].append(p)
                    continue
            cur = {"params": [p]}
            runs.append(cur)
        key = tuple(tuple(q.data_ptr() for q in r["params"]) for r in runs)
        cached = self._runs.get(gi)
        if cached is not None and cached[0] == key:
            # still valid only while the state's moments ARE the run buffers (load_state_dict replaces them)
            def bound(r):
                ea = self.state[r["params"][0]].get("exp_avg")
                return ea is not None and ea.data_ptr() == r["m"].data_ptr()
            if all(bound(r) for r in cached[1]):
                return cached[1]
        for r in runs:
            first, last = r["params"][0], r["params"][-1]
            r["n"] = (last.data_ptr() + last.numel() * 4 - first.data_ptr()) // 4
            old = cached[1] if cached is not None else []
            m = torch.zeros(r["n"], dtype=torch.float32, device=first.device)
            v = torch.zeros(r["n"], dtype=torch.float32, device=first.device)
            for p in r["params"]:
                off = (p.data_ptr() - first.data_ptr()) // 4
                st = self.state[p]
                # moments in the parameter's own element order (channels_last for the KRSC conv weights)
                mv = torch.as_strided(m, tuple(p.shape), p.stride(), off)
                vv = torch.as_strided(v, tuple(p.shape), p.stride(), off)
                if "exp_avg" in st:          # resumed / re-planned: keep the moments
                    mv.copy_(st["exp_avg"].to(p.device)); vv.copy_(st["exp_avg_sq"].to(p.device))
                st["exp_avg"], st["exp_avg_sq"] = mv, vv
                if "step" not in st:
                    st["step"] = torch.tensor(0.0)
            r["m"], r["v"] = m, v
            # device-side step counter (keeps captured CUDA graphs correct on replay)
            st0 = self.state[r["params"][0]]
            r["step_host"] = int(float(st0["step"]))
            r["step_dev"] = torch.full((1,), r["step_host"], dtype=torch.int32, device=first.device)
            del old
        self._runs[gi] = (key, runs)
        return runs

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, max_grad_norm=0.0, clip_groups=(0,)):
        """grad_scale multiplies every gradient first (1/world_size after a sum-allreduce);
        max_grad_norm > 0 applies clip_grad_norm_ (fused) to the param groups listed in `clip_groups` --
        the norm is taken over those groups only and only they are scaled.  Default: group 0, the model
        (common/train.py:357-358 clips model.parameters(); the criterion's sax/saq/srx/srq groups of
        scripts/train.py:104-110 are neither counted nor scaled).  clip_groups=None: every group."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        groups = []
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or not _dense(p) or p.grad.stride() != p.stride():
                    raise RuntimeError("FusedAdam needs dense fp32 CUDA parameters with gradients in the same strides "
                                       "(contiguous, or channels_last views of a flat buffer; no CPU path)")
            groups.append((group, self._plan(gi, params), clip_groups is None or gi in clip_groups))
        if not groups:
            return loss
        dev = groups[0][1][0]["params"][0].device
        sq_ptr = None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            if max_grad_norm and max_grad_norm > 0.0:
                if self._scratch is None or self._scratch.device != dev:
                    self._scratch = torch.zeros(2048, dtype=torch.float32, device=dev)
                total = self._scratch[1024:1025]
                total.zero_()
                part = self._scratch[1025:1026]
                for group, runs, clipped in groups:
                    if not clipped:
                        continue
                    for r in runs:
                        g0 = r["params"][0].grad
                        _lib.check(L.mapnet_sqnorm(g0.data_ptr(), r["n"], self._scratch.data_ptr(),
                                                   part.data_ptr(), st), "mapnet_sqnorm")
                        total += part
                sq_ptr = total.data_ptr()
            for group, runs, clipped in groups:
                b1, b2 = group["betas"]
                for r in runs:
                    p0 = r["params"][0]
                    r["step_host"] += 1
                    stp_new = torch.tensor(float(r["step_host"]))
                    for p in r["params"]:
                        self.state[p]["step"] = stp_new
                    _lib.check(L.mapnet_adam_step_dev(
                        p0.data_ptr(), p0.grad.data_ptr(), r["m"].data_ptr(), r["v"].data_ptr(), r["n"],
                        float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                        float(group["weight_decay"]), r["step_dev"].data_ptr(), float(grad_scale),
                        ctypes.c_void_p(sq_ptr) if (sq_ptr and clipped) else None, float(max_grad_norm or 0.0), st),
                        "mapnet_adam_step_dev")
        return loss

    def advance_host_step(self, n=1):
        """After replaying a captured graph that contains step(): the device counters advanced,
        bring the host mirrors (state['step'], used by state_dict()) in line."""
        for key, runs in self._runs.values():
            for r in runs:
                r["step_host"] += n
                stp_new = torch.tensor(float(r["step_host"]))
                for p in r["params"]:
                    self.state[p]["step"] = stp_new


class Optimizer:
    """Learner + learning-rate schedule behind the reference's wrapper surface
    (``Optimizer(params, method, base_lr, weight_decay, **kwargs)``, ``.learner``, ``.adjust_lr(epoch)``,
    ``.mult_lr(f)`` -- what /root/reference/common/train.py:128,281,359 and scripts/train.py:112 use).
    'adam' is the fused flat-buffer Adam above; 'sgd' / 'rmsprop' go to torch.optim unchanged.  Only SGD has a
    schedule: the rate drops by ``lr_decay`` at every epoch listed in ``lr_stepvalues``."""

    _LEARNERS = {"sgd": optim.SGD, "adam": FusedAdam, "rmsprop": optim.RMSprop}

    def __init__(self, params, method, base_lr, weight_decay, **kwargs):
        if method not in self._LEARNERS:
            raise NotImplementedError(method)
        self.method, self.base_lr = method, base_lr
        if method == "sgd":
            self.lr_decay = kwargs.pop("lr_decay")
            self.lr_stepvalues = sorted(kwargs.pop("lr_stepvalues"))
        self.learner = self._LEARNERS[method](params, lr=base_lr, weight_decay=weight_decay, **kwargs)

    def _each_group(self, fn):
        for group in self.learner.param_groups:
            group["lr"] = fn(group["lr"])

    def adjust_lr(self, epoch):
        """Rate for `epoch`; written into every param group when there is a schedule (SGD)."""
        if self.method != "sgd":
            return self.base_lr
        drops = bisect.bisect_right(self.lr_stepvalues, epoch)      # step values already reached
        lr = self.base_lr * self.lr_decay ** drops
        self._each_group(lambda _: lr)
        return lr

    def mult_lr(self, f):
        self._each_group(lambda lr: lr * f)
