"""Batched inference + pose-graph optimisation: the evaluation loop of the reference on the GPU.

Host-side mirror of /root/reference/scripts/eval.py:137-199.  The reference runs batch size 1 ("batch_size MUST be 1",
:136-138), moves every prediction to numpy, applies qexp per pose in Python, optionally runs a dense numpy PGO per
frame and keeps the middle prediction of each tuple.  Here: ``predict`` runs the eval-mode forward in batches of tuples
through the same fprop engines as training (BatchNorm running statistics, no dropout at eval -- see note), ``post`` does
qexp + un-normalisation for all poses in one launch, ``pgo`` optimises all windows in one launch (csrc/pgo.cu), and
``pose_errors`` gives the translation / rotation errors eval.py prints (:193-199).

Note on dropout: models/posenet.py:68-69 calls F.dropout without ``training=``, so the reference drops features at eval
time too whenever droprate > 0; scripts/eval.py builds the model with the ini's dropout.  ``predict`` follows the
module (geomapnet_b200.models.PoseNet keeps that behaviour); pass a model built with droprate=0 for deterministic poses.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .common.pgo import optimize_pose_windows

__all__ = ["predict", "post", "pgo", "pose_errors", "evaluate_tuples"]


@torch.no_grad()
def predict(model, x, batch=64):
    """model(x) in eval mode, `batch` leading entries at a time.  x: [L,3,H,W] (PoseNet) or [L,T,3,H,W] (MapNet), on the
    GPU or pinned host memory.  Returns float32 [L,6] / [L,T,6] on the device."""
    was_training = model.training
    model.eval()
    outs = []
    try:
        for i in range(0, x.shape[0], batch):
            xb = x[i:i + batch]
            if not xb.is_cuda:
                xb = xb.cuda(non_blocking=True)
            outs.append(model(xb).float())
    finally:
        model.train(was_training)
    return torch.cat(outs, 0)


def post(pred6, pose_m=None, pose_s=None):
    """eval.py:163-181: (t, log q) -> (t * pose_s + pose_m, qexp(log q)) for every pose; [..., 6] float32 -> [..., 7] float64."""
    p = pred6.detach().float().contiguous()
    if not p.is_cuda:
        raise RuntimeError("geomapnet_b200.inference has no CPU path: predictions must be CUDA tensors")
    n = p.numel() // 6
    out = torch.empty(p.shape[:-1] + (7,), dtype=torch.float64, device=p.device)
    m = s = None
    if pose_m is not None:
        m = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(pose_m).reshape(3)])
        s = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(pose_s).reshape(3)])
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().mapnet_pose_post(p.data_ptr(), out.data_ptr(), n, m, s, _lib.stream_ptr()), "mapnet_pose_post")
    return out


def pgo(pred7, vos7, sax=1.0, saq=1.0, srx=1.0, srq=1.0, fc_vos=False):
    """eval.py:172-178 for all windows at once: pred7 [W,N,7], vos7 [W,N-1,7] -> [W,N,7]"""
    return optimize_pose_windows(pred7, vos7, fc_vos=fc_vos, sax=sax, saq=saq, srx=srx, srq=srq)


def pose_errors(pred7, targ7):
    """eval.py:80-81,193-199: per pose ||t_pred - t_gt|| and the quaternion angular error in degrees
    (common/pose_utils.py:358-368).  Returns (t_err [L], q_err_deg [L]) float64 on the device."""
    p, t = pred7.double(), targ7.double().to(pred7.device)
    t_err = (p[..., :3] - t[..., :3]).norm(dim=-1)
    d = (p[..., 3:] * t[..., 3:]).sum(-1).abs().clamp(-1.0, 1.0)
    return t_err, 2.0 * torch.acos(d) * 180.0 / np.pi


def evaluate_tuples(model, x, targ6, vos7=None, pose_m=None, pose_s=None, batch=32, **sigmas):
    """The whole loop of eval.py:147-199 for a MapNet: x [L,T,3,H,W], targ6 [L,T,6] (log-q targets), vos7 [L,T-1,7]
    or None (no PGO).  Keeps the middle prediction of every tuple.  Returns dict(pred7 [L,7], targ7 [L,7], t_err, q_err)."""
    out6 = predict(model, x, batch)                     # [L,T,6]
    T = out6.shape[1]
    pred = post(out6)                                   # normalised translations, as PGO sees them (eval.py:172-178)
    targ = post(targ6.to(out6.device))
    if vos7 is not None:
        pred = pgo(pred, vos7.to(out6.device), **sigmas)
    if pose_m is not None:
        m = torch.as_tensor(np.asarray(pose_m, dtype=np.float64), device=pred.device)
        s = torch.as_tensor(np.asarray(pose_s, dtype=np.float64), device=pred.device)
        pred = torch.cat((pred[..., :3] * s + m, pred[..., 3:]), -1)
        targ = torch.cat((targ[..., :3] * s + m, targ[..., 3:]), -1)
    mid = T // 2                                        # eval.py:183-185 (py2 integer division)
    p_mid, t_mid = pred[:, mid], targ[:, mid]
    t_err, q_err = pose_errors(p_mid, t_mid)
    return dict(pred7=p_mid, targ7=t_mid, t_err=t_err, q_err=q_err)
