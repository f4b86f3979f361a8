"""CUDA-graph replay of the training step gives the same parameters as the eager step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_step_matches_eager():
    from oracle import weights
    from helpers import make_product_model, make_product_criterion
    from geomapnet_b200.common.optimizer import Optimizer
    from geomapnet_b200.graph import GraphedTrainStep
    st = weights.make_state(7)
    cfg = dict(kind="mapnet", N=2, T=3, H=64, W=64)
    batches = [weights.make_inputs(cfg, 20 + i) for i in range(4)]
    finals = []
    for use_graph in (False, False, True):
        model, net = make_product_model(st, "mapnet", "fp32")
        crit = make_product_criterion("mapnet")
        opt = Optimizer([{"params": model.parameters()}, {"params": list(crit.parameters())}], "adam", 1e-3, 5e-4)
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
                losses.append(float(loss))
        torch.cuda.synchronize()
        finals.append((losses, {k: v.detach().cpu().clone() for k, v in net.state_dict().items()},
                       [float(p) for p in crit.parameters()]))
    (la, sa, ca), (l2, s2, c2), (lb, sb, cb) = finals
    assert abs(la[0] - lb[0]) <= 1e-6 * abs(la[0])          # before any update: identical
    # Yardstick: two EAGER runs differ by fp32 atomics-order noise, amplified by Adam's sign-like first
    # steps (an element whose gradient is ~0 moves by +-lr either way).  ONE noise sample is a poor bound
    # for the worst element of another sample (the ratio of two such maxima is heavy-tailed: on max-norms
    # alone this test failed in one of three B200 runs of the same fp32 code), so the tight checks are on MEAN deviations,
    # where the handful of flipped elements average out, and the max-norm checks only guard against gross
    # errors (a stale counter, a missed kernel, a wrong input buffer are all O(1e-2) or worse).
    fkeys = [k for k in sa if sa[k].dtype.is_floating_point]
    noise = max(abs(a - b) / abs(a) for a, b in zip(la, l2))
    gl = max(abs(a - b) / abs(a) for a, b in zip(la, lb))

    def mean_dev(u, v):
        num = sum(float((u[k] - v[k]).abs().sum()) for k in fkeys)
        den = sum(float(u[k].abs().sum()) for k in fkeys)
        return num / den

    def max_dev(u, v):
        return max(float((u[k] - v[k]).abs().max() / (u[k].abs().max() + 1e-12)) for k in fkeys)

    mn, mg = mean_dev(sa, s2), mean_dev(sa, sb)
    xn, xg = max_dev(sa, s2), max_dev(sa, sb)
    print("graph-vs-eager: loss noise %.2e graph %.2e | mean param dev noise %.2e graph %.2e | max param dev noise %.2e graph %.2e"
          % (noise, gl, mn, mg, xn, xg))
    assert gl <= 5 * noise + 2e-4, (la, l2, lb)
    assert mg <= 5 * mn + 1e-6, (mn, mg)
    assert xg <= 20 * xn + 5e-3, (xn, xg)
    for k in sa:
        if not sa[k].dtype.is_floating_point:
            assert torch.equal(sa[k], sb[k]), k
    for a, b in zip(ca, cb):
        assert abs(a - b) <= 5e-3 * max(1.0, abs(a)), (ca, cb)
