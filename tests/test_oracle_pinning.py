"""The CPU oracle (oracle/mapnet_oracle.py) pinned against outputs of the REFERENCE:
 * everywhere: the committed goldens in tests/golden/ (made by oracle/make_goldens.py
   from /root/reference's own modules);
 * in the build container: the reference modules executed live through oracle.ref_loader."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import mapnet_oracle as O
from oracle import ref_loader, weights

SV = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)


def _golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "step_%s.npz" % name))
    return g, ast.literal_eval(str(g["cfg"]))


@pytest.mark.parametrize("name", ["posenet_tiny", "posenet_ragged", "mapnet_tiny", "online_tiny", "online_gps_tiny",
                                  "posenet_7scenes_b4"])
def test_oracle_step_matches_reference_golden(golden_dir, name):
    g, cfg = _golden(golden_dir, name)
    st = weights.make_state(int(g["seed"]))
    x, targ = weights.make_inputs(cfg, int(g["seed"]))
    if "x" in g.files:
        assert np.array_equal(x.numpy(), g["x"]), "input generator drifted"
    assert np.array_equal(targ.numpy(), g["targ"])
    r = O.train_step(cfg["kind"], st, x, targ, SV, lr=cfg.get("lr", 1e-4), weight_decay=cfg.get("wd", 5e-4),
                     max_grad_norm=cfg.get("clip", 0.0), filter_nans=cfg["kind"].startswith("online"))
    # same torch build on both sides -> the restatement reproduces the reference bit for bit
    # (tolerances only guard against thread-count dependent reduction order)
    assert abs(float(r["loss"]) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(r["pred"].numpy().reshape(g["pred"].shape), g["pred"], rtol=1e-5, atol=1e-6)
    for i, n in enumerate(g["grad_names"]):
        t = r["grads"][str(n)].double()
        assert abs(float(t.norm()) - g["grad_norm"][i]) <= 1e-4 * g["grad_norm"][i] + 1e-12, str(n)
    for i, n in enumerate(g["post_names"]):
        t = r["new_state"][str(n)].double()
        assert abs(float(t.norm()) - g["post_norm"][i]) <= 1e-5 * g["post_norm"][i] + 1e-12, str(n)
    for i, n in enumerate(g["sgrad_names"]):
        if not np.isnan(g["sgrads"][i]):
            assert abs(float(r["sgrads"][str(n)]) - g["sgrads"][i]) <= 1e-5 * abs(g["sgrads"][i]) + 1e-7


def test_oracle_pose_math_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pose_math.npz"))
    v = torch.tensor(g["qexp_in"])
    np.testing.assert_allclose(O.qexp_t(v).numpy(), g["qexp_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(O.qlog_t(torch.tensor(g["qexp_out"])).numpy(), g["qlog_out"], rtol=1e-6, atol=1e-7)
    for key in ("n5t3", "n16t5", "n3t2", "n4t7"):
        p = torch.tensor(g["vos_in_" + key]).requires_grad_(True)
        np.testing.assert_allclose(O.calc_vos_simple(p).detach().numpy(), g["vos_simple_" + key], rtol=0, atol=0)
        vv = O.calc_vos(p)
        np.testing.assert_allclose(vv.detach().numpy(), g["vos_" + key], rtol=1e-5, atol=1e-6)
        (vv * torch.tensor(g["vos_w_" + key])).sum().backward()
        np.testing.assert_allclose(p.grad.numpy(), g["vos_grad_" + key], rtol=1e-4, atol=1e-5)


def test_known_answer_properties():
    """SURVEY.md section 4 (iii),(iv),(vi): qlog(qexp(v)) == v for |v| < pi; |qexp(v)| = 1;
    calc_vos_simple is a plain difference."""
    g = torch.Generator().manual_seed(1)
    v = torch.randn(100, 3, generator=g)
    v = v / v.norm(dim=1, keepdim=True) * (torch.rand(100, 1, generator=g) * 3.0)
    q = O.qexp_t(v)
    assert float((q.norm(dim=1) - 1).abs().max()) < 1e-6
    assert float((O.qlog_t(q) - v).abs().max()) < 1e-5
    p = torch.randn(4, 5, 6, generator=g)
    assert torch.equal(O.calc_vos_simple(p), p[:, 1:] - p[:, :-1])


def test_flop_model():
    macs, conv1 = O.conv_macs_per_image(256, 256)
    assert macs == 4784652288                      # SURVEY.md section 8d, hook-measured on the live module
    assert abs(O.train_flops_per_image(256, 256) - 28.400e9) < 0.01e9


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("kind,cfg", [("posenet", dict(kind="posenet", N=2, H=64, W=64)),
                                      ("online", dict(kind="online", N=2, T=4, H=64, W=64))])
def test_oracle_matches_live_reference(kind, cfg):
    ns = ref_loader.load()
    st = weights.make_state(3)
    x, targ = weights.make_inputs(cfg, 3)
    model = ref_loader.build_reference_model(st, "posenet" if kind == "posenet" else "mapnet")
    model.train()
    if kind == "posenet":
        crit = ns.PoseNetCriterion(sax=0.0, saq=-3.0, learn_beta=True)
    else:
        crit = ns.MapNetOnlineCriterion(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0, learn_beta=True, learn_gamma=True)
    loss, out, grads, cg = ref_loader.reference_step(model, crit, x, targ)
    r = O.train_step(kind, st, x, targ, SV)
    assert abs(float(r["loss"]) - loss) <= 1e-6 * abs(loss)
    assert float((out - r["pred"].view_as(out)).abs().max()) <= 1e-6
    for k, v in grads.items():
        k2 = k.replace("mapnet.", "", 1)
        assert float((v - r["grads"][k2]).norm()) <= 1e-5 * float(v.norm()) + 1e-12, k


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_oracle_nan_filter_matches_live_reference_hook():
    """models/posenet.py:28-34,50-51 with a NaN-producing gradient: degenerate state (fc_wpqr = 0 -> every predicted
    rotation is the identity -> qlog's backward gives NaN, the case the hook exists for) through the reference
    MapNet + MapNetOnlineCriterion with filter_nans=True, against the oracle.  The hook zeroes NaNs of the Linear's
    bias / input / weight gradients: whole weight rows and whole samples drop out, not single entries."""
    ns = ref_loader.load()
    st = weights.make_state(3)
    st["fc_wpqr.weight"] = torch.zeros_like(st["fc_wpqr.weight"])
    st["fc_wpqr.bias"] = torch.zeros_like(st["fc_wpqr.bias"])
    cfg = dict(kind="online", N=2, T=4, H=64, W=64)
    x, targ = weights.make_inputs(cfg, 3)
    model = ref_loader.build_reference_model(st, "mapnet", filter_nans=True)
    model.train()
    crit = ns.MapNetOnlineCriterion(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0, learn_beta=True, learn_gamma=True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss, out, grads, cg = ref_loader.reference_step(model, crit, x, targ, do_step=False)
    r = O.train_step("online", st, x, targ, SV, filter_nans=True, do_step=False)
    # the unfiltered gradient really is NaN in the rotation half
    r_nofilter = O.train_step("online", st, x, targ, SV, filter_nans=False, do_step=False)
    assert bool(torch.isnan(r_nofilter["grads"]["fc_wpqr.weight"]).any())
    n_zero_rows = 0
    for k, v in grads.items():
        k2 = k.replace("mapnet.", "", 1)
        assert bool(torch.isfinite(v).all()) and bool(torch.isfinite(r["grads"][k2]).all()), k
        assert float((v - r["grads"][k2]).norm()) <= 1e-5 * float(v.norm()) + 1e-12, k
    for j in range(3):
        n_zero_rows += int(bool((grads["mapnet.fc_wpqr.weight"][j] == 0).all()))
    assert n_zero_rows >= 1, "the degenerate input was meant to wipe at least one fc_wpqr row"


def test_baseline_config0_plumbing_single_frame():
    """BASELINE.json configs[0] / SURVEY.md section 8d "Config 1 (plumbing)": PoseNet ResNet-34 on ONE synthetic 256x256 frame,
    PoseNetCriterion(sax=0, saq=-3), seed 7, on the CPU: a [1,6] output and a finite loss -- through the oracle everywhere,
    and against the reference modules executed live where the reference tree exists."""
    st = weights.make_state(7)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 3, 256, 256, generator=g)
    targ = torch.randn(1, 6, generator=g)
    r = O.train_step("posenet", st, x, targ, SV, do_step=False)
    assert tuple(r["pred"].shape) == (1, 6) and bool(torch.isfinite(r["pred"]).all())
    assert np.isfinite(float(r["loss"]))
    # the criterion value follows from the prediction: exp(-sax) L1(t) + sax + exp(-saq) L1(q) + saq (criterion.py:42-52)
    want = float((r["pred"][:, :3] - targ[:, :3]).abs().mean() + 0.0
                 + np.exp(3.0) * (r["pred"][:, 3:] - targ[:, 3:]).abs().mean() - 3.0)
    assert abs(float(r["loss"]) - want) <= 1e-5 * abs(want)
    if ref_loader.available():
        ns = ref_loader.load()
        model = ref_loader.build_reference_model(st, "posenet")
        model.train()
        crit = ns.PoseNetCriterion(sax=0.0, saq=-3.0, learn_beta=True)
        loss, out, _, _ = ref_loader.reference_step(model, crit, x, targ, do_step=False)
        assert tuple(out.shape) == (1, 6)
        assert abs(loss - float(r["loss"])) <= 1e-6 * abs(loss)
        assert float((out - r["pred"]).abs().max()) <= 1e-6
