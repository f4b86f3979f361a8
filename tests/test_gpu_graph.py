"""CUDA-graph replay of the training step gives the same parameters as the eager step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_step_matches_eager():
    """Same four batches through (a) the eager modules, (b) the eager modules again, (c) GraphedTrainStep.

    What can go wrong in the graph path is structural -- a stale input buffer, a host-side counter (Adam bias
    correction, dropout offset, num_batches_tracked) frozen at its capture value, construction leaking its warm-up
    steps into the training state -- and every one of those moves the parameters by a sizeable FRACTION of a step.
    What is NOT a defect is fp32 atomics-order noise: Adam's first steps are sign-like (every element moves by
    +-lr, an element whose total gradient is ~0 goes either way), and at a large learning rate this randomly
    initialised 34-layer BatchNorm network is chaotic (at lr 1e-3 the loss goes 32.8 -> 108 -> 89 -> 71 and two
    EAGER runs already differ by 1.6e-4 at step 4; measured on B200, the graph run landed anywhere between 1.5e-4
    and 8e-3 of them -- max-norm bounds scaled by one noise sample made this test fail in 2 of 4 runs).  So the
    step size is small (lr 1e-5: mapnet++_7Scenes.ini:18), and the tight checks are on the MEAN parameter deviation
    relative to the mean distance the parameters travelled, where the handful of flipped elements average out."""
    from oracle import weights
    from helpers import make_product_model, make_product_criterion
    from geomapnet_b200.common.optimizer import Optimizer
    from geomapnet_b200.graph import GraphedTrainStep
    st = weights.make_state(7)
    cfg = dict(kind="mapnet", N=2, T=3, H=64, W=64)
    batches = [weights.make_inputs(cfg, 20 + i) for i in range(4)]
    lr = 1e-5
    finals = []
    for use_graph in (False, False, True):
        model, net = make_product_model(st, "mapnet", "fp32")
        crit = make_product_criterion("mapnet")
        opt = Optimizer([{"params": model.parameters()}, {"params": list(crit.parameters())}], "adam", lr, 5e-4)
        model.train()
        losses = []
        if use_graph:
            x0, t0 = batches[0]
            # construction warms up and captures but must leave the training state untouched
            g = GraphedTrainStep(model, crit, opt, x0.cuda(), t0.cuda(), max_grad_norm=5.0, warmup=2)
            for k, v in net.state_dict().items():
                assert torch.equal(v.cpu(), st[k]), "GraphedTrainStep construction changed %s" % k
            for x, t in batches:
                losses.append(float(g(x.cuda(), t.cuda())))
            sd = opt.learner.state_dict()
            assert float(sd["state"][0]["step"]) == len(batches)
        else:
            for x, t in batches:
                out = model(x.cuda()); loss = crit(out, t.cuda())
                opt.learner.zero_grad(); loss.backward(); opt.learner.step(max_grad_norm=5.0)
                losses.append(float(loss.detach()))
        torch.cuda.synchronize()
        finals.append((losses, {k: v.detach().cpu().clone() for k, v in net.state_dict().items()},
                       [float(p.detach()) for p in crit.parameters()]))
    (la, sa, ca), (l2, s2, c2), (lb, sb, cb) = finals
    assert abs(la[0] - lb[0]) <= 1e-6 * abs(la[0])          # before any update: identical
    pkeys = [k for k in sa if sa[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))]
    bkeys = [k for k in sa if k.endswith(("running_mean", "running_var"))]

    def mean_dev(u, v, keys):
        return sum(float((u[k] - v[k]).abs().sum()) for k in keys) / sum(u[k].numel() for k in keys)

    moved = mean_dev(sa, st, pkeys)                           # ~ 4 sign-like steps of lr for most elements
    noise, dev = mean_dev(sa, s2, pkeys), mean_dev(sa, sb, pkeys)
    bmoved, bdev = mean_dev(sa, st, bkeys), mean_dev(sa, sb, bkeys)
    ln = max(abs(a - b) / abs(a) for a, b in zip(la, l2))
    lg = max(abs(a - b) / abs(a) for a, b in zip(la, lb))
    print("graph-vs-eager: parameters moved %.3e (lr %.0e x %d steps); mean deviation eager/eager %.3e, graph/eager %.3e; "
          "BN buffers moved %.3e, graph/eager %.3e; loss deviation eager/eager %.2e, graph/eager %.2e"
          % (moved, lr, len(batches), noise, dev, bmoved, bdev, ln, lg))
    assert moved > 1.5 * lr                                   # the optimizer really stepped every time
    assert dev <= 0.05 * moved, (moved, noise, dev)           # a frozen counter / stale buffer costs >= 25 %
    assert bdev <= 0.02 * bmoved, (bmoved, bdev)              # running statistics: momentum updates of 4 forwards
    assert lg <= 2e-3, (la, l2, lb)
    for k in sa:
        if not sa[k].dtype.is_floating_point:
            assert torch.equal(sa[k], sb[k]), k               # num_batches_tracked
    for a, b in zip(ca, cb):
        assert abs(a - b) <= 0.05 * len(batches) * lr + 1e-9, (ca, cb)
