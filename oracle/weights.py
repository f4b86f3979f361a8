"""Deterministic seed-driven weights for PoseNet(ResNet-34) -- TEST INFRASTRUCTURE.

Both the reference modules (via ref_loader) and the product modules are loaded
from the same ``make_state(seed)`` so parity starts from identical parameters
without committing 89 MB of weights.  torch's CPU generator is deterministic
across machines for a fixed torch version (the GPU box runs the same image).

Key names/shapes follow the live reference module (torchvision resnet34 wrapped
by models/posenet.py:37-49; listed in SURVEY.md section 8b): 222 state_dict entries,
first key ``feature_extractor.conv1.weight``.
"""
from collections import OrderedDict

import torch

STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))  # (planes, blocks, stride)


def state_spec(feat_dim=2048):
    """[(name, shape, kind)] in reference state_dict order.
    kind in {conv, bn_w, bn_b, bn_rm, bn_rv, bn_nbt, fc_w, fc_b}."""
    spec = []

    def bn(prefix, c):
        spec.append((prefix + ".weight", (c,), "bn_w"))
        spec.append((prefix + ".bias", (c,), "bn_b"))
        spec.append((prefix + ".running_mean", (c,), "bn_rm"))
        spec.append((prefix + ".running_var", (c,), "bn_rv"))
        spec.append((prefix + ".num_batches_tracked", (), "bn_nbt"))

    fe = "feature_extractor."
    spec.append((fe + "conv1.weight", (64, 3, 7, 7), "conv"))
    bn(fe + "bn1", 64)
    inpl = 64
    for li, (planes, nblk, stride) in enumerate(STAGES, start=1):
        for b in range(nblk):
            p = "%slayer%d.%d." % (fe, li, b)
            s = stride if b == 0 else 1
            spec.append((p + "conv1.weight", (planes, inpl, 3, 3), "conv"))
            bn(p + "bn1", planes)
            spec.append((p + "conv2.weight", (planes, planes, 3, 3), "conv"))
            bn(p + "bn2", planes)
            if s != 1 or inpl != planes:
                spec.append((p + "downsample.0.weight", (planes, inpl, 1, 1), "conv"))
                bn(p + "downsample.1", planes)
            inpl = planes
    spec.append((fe + "fc.weight", (feat_dim, 512), "fc_w"))
    spec.append((fe + "fc.bias", (feat_dim,), "fc_b"))
    spec.append(("fc_xyz.weight", (3, feat_dim), "fc_w"))
    spec.append(("fc_xyz.bias", (3,), "fc_b"))
    spec.append(("fc_wpqr.weight", (3, feat_dim), "fc_w"))
    spec.append(("fc_wpqr.bias", (3,), "fc_b"))
    return spec


def make_state(seed=7, feat_dim=2048, dtype=torch.float32):
    """Kaiming-normal conv/fc weights (fan_in, gain sqrt(2), as
    models/posenet.py:58-63 does when pretrained=False) but with NON-trivial BN
    affine parameters, running stats and biases so every term of the math is
    exercised by the parity tests."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    st = OrderedDict()
    for name, shape, kind in state_spec(feat_dim):
        if kind in ("conv", "fc_w"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            std = (2.0 / fan_in) ** 0.5
            t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        elif kind == "bn_w":
            t = 1.0 + 0.2 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind in ("bn_b", "fc_b"):
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "bn_rm":
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "bn_rv":
            t = 1.0 + 0.2 * torch.rand(shape, generator=g, dtype=torch.float32)
        elif kind == "bn_nbt":
            st[name] = torch.zeros((), dtype=torch.int64)
            continue
        else:
            raise AssertionError(kind)
        st[name] = t.to(dtype)
    return st


def make_inputs(cfg, seed=7):
    """Synthetic inputs per SURVEY.md section 8d.  cfg keys: kind in
    {posenet, mapnet, online, online_gps}, N, T (frames per tuple as the model
    sees them), H, W.  Returns (x, targ) float32 CPU tensors."""
    import math
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) * 1000003 + 17)
    kind, N, H, W = cfg["kind"], cfg["N"], cfg["H"], cfg["W"]

    def logq(n):
        v = torch.randn(n, 3, generator=g)
        v = v / v.norm(dim=1, keepdim=True)
        ang = (torch.rand(n, 1, generator=g) - 0.5) * math.pi  # U(-pi/2, pi/2)
        return v * ang

    if kind == "posenet":
        x = torch.randn(N, 3, H, W, generator=g)
        targ = torch.cat([torch.randn(N, 3, generator=g), logq(N)], dim=1)
        return x, targ
    T = cfg["T"]
    x = torch.randn(N, T, 3, H, W, generator=g)

    def traj(n, t):
        p0 = torch.cat([torch.randn(n, 1, 3, generator=g), logq(n).view(n, 1, 3)], dim=2)
        d = 0.05 * torch.randn(n, t, 6, generator=g)
        d[:, 0] = 0
        return p0 + d.cumsum(dim=1)

    if kind == "mapnet":
        return x, traj(N, T)
    half = T // 2
    if kind == "online":      # [T/2 abs poses || T/2-1 VOs]  (composite.py:117-126)
        vo = 0.05 * torch.randn(N, half - 1, 6, generator=g)
        return x, torch.cat([traj(N, half), vo], dim=1)
    if kind == "online_gps":  # [T/2 abs poses || T/2 abs poses]
        return x, torch.cat([traj(N, half), traj(N, half)], dim=1)
    raise ValueError(kind)
