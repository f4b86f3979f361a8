"""Shared helpers for the parity tests (test infrastructure)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SVALS = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, "step_%s.npz" % name), allow_pickle=False)
    cfg = ast.literal_eval(str(g["cfg"]))
    return g, cfg


def make_product_model(state, kind, precision, filter_nans=False, droprate=0.0):
    import torchvision
    from geomapnet_b200.models.posenet import PoseNet, MapNet
    fe = torchvision.models.resnet34(weights=None)
    net = PoseNet(fe, droprate=droprate, pretrained=False, filter_nans=filter_nans, precision=precision)
    net.load_state_dict({k: v.clone() for k, v in state.items()})
    model = net if kind == "posenet" else MapNet(net)
    return model.cuda(), net


def make_product_criterion(kind, learn=True):
    from geomapnet_b200.common.criterion import PoseNetCriterion, MapNetCriterion, MapNetOnlineCriterion
    if kind == "posenet":
        c = PoseNetCriterion(sax=SVALS["sax"], saq=SVALS["saq"], learn_beta=learn)
    else:
        kw = dict(sax=SVALS["sax"], saq=SVALS["saq"], srx=SVALS["srx"], srq=SVALS["srq"],
                  learn_beta=learn, learn_gamma=learn)
        c = MapNetCriterion(**kw) if kind == "mapnet" else MapNetOnlineCriterion(gps_mode=(kind == "online_gps"), **kw)
    return c.cuda()


def product_step(model, net, crit, x, targ, lr=1e-4, wd=5e-4, clip=0.0, do_step=True):
    """The 10 lines of common/train.py:339-361 around the product modules."""
    from geomapnet_b200.common.optimizer import Optimizer
    params = [{"params": model.parameters()}]
    cp = [p for p in crit.parameters() if p.requires_grad]
    if cp:
        params.append({"params": cp})
    opt = Optimizer(params=params, method="adam", base_lr=lr, weight_decay=wd)
    xv = x.cuda().requires_grad_(True)
    out = model(xv)
    loss = crit(out, targ.cuda())
    opt.learner.zero_grad()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    sgrads = {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in crit.named_parameters()}
    if do_step:
        if clip > 0:
            opt.learner.step(max_grad_norm=clip)
        else:
            opt.learner.step()
    return loss.detach(), out.detach(), grads, sgrads


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def compare_with_golden(g, loss, pred, grads, sgrads, net, crit, tol, report=None):
    """Returns dict of measured errors; asserts against tol = dict(loss, pred, grad, post)."""
    errs = {}
    errs["loss"] = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    errs["pred"] = rel_err(pred.cpu().numpy().reshape(g["pred"].shape), g["pred"])
    gn, worst = 0.0, None
    gh = 0.0
    for i, name in enumerate(g["grad_names"]):
        t = grads[str(name)].double().cpu()
        ref_norm = float(g["grad_norm"][i])
        e = abs(float(t.norm()) - ref_norm) / (ref_norm + 1e-30)
        if e > gn:
            gn, worst = e, str(name)
        head = t.flatten()[:8].numpy()
        ref_head = g["grad_head"][i][:head.size]
        gh = max(gh, float(np.abs(head - ref_head).max() / (ref_norm / np.sqrt(max(t.numel(), 1)) + 1e-30)))
    errs["grad_norm"] = gn
    errs["grad_norm_worst"] = worst
    errs["grad_head_vs_rms"] = gh
    # element-wise comparison of the named gradient tensors (strided sample of the large ones): relative L2 error
    gf, gf_worst, per = 0.0, None, {}
    if "grad_full_names" in g.files:
        for i, name in enumerate(g["grad_full_names"]):
            ref = g["grad_full_%d" % i].astype(np.float64)
            t = grads[str(name)].double().cpu().flatten().numpy()
            stride = max(1, t.size // 40000)            # oracle/make_goldens.py: sample_stride
            t = t[::stride]
            assert t.shape == ref.shape, (str(name), t.shape, ref.shape)
            e = float(np.linalg.norm(t - ref) / (np.linalg.norm(ref) + 1e-30))
            per[str(name)] = e
            if e > gf:
                gf, gf_worst = e, str(name)
    errs["grad_full"] = gf
    errs["grad_full_worst"] = gf_worst
    errs["grad_full_per_tensor"] = per
    sd = net.state_dict()
    pn = 0.0
    for i, name in enumerate(g["post_names"]):
        t = sd[str(name)].double().cpu()
        pn = max(pn, abs(float(t.norm()) - float(g["post_norm"][i])) / (float(g["post_norm"][i]) + 1e-30))
    errs["post_norm"] = pn
    se = 0.0
    for i, name in enumerate(g["sgrad_names"]):
        ref = float(g["sgrads"][i])
        if not np.isnan(ref):
            se = max(se, abs(float(sgrads[str(name)]) - ref) / (abs(ref) + 1e-12))
    errs["sgrad"] = se
    if report is not None:
        report.append(errs)
    assert errs["loss"] <= tol["loss"], errs
    assert errs["pred"] <= tol["pred"], errs
    assert errs["grad_norm"] <= tol["grad"], errs
    assert errs["grad_head_vs_rms"] <= tol["grad_head"], errs
    assert errs["grad_full"] <= tol.get("grad_full", 1e9), errs
    assert errs["post_norm"] <= tol["post"], errs
    assert errs["sgrad"] <= tol["sgrad"], errs
    return errs
