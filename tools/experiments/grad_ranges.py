"""CPU experiment: dynamic range of the backward conv operands (d loss / d conv output) per layer, relative to the
gradient that enters the trunk -- decides whether ONE power-of-two scale can place every gradient tensor of a step
inside fp16's window for the strict tensor-core mode (split fp16 planes: abs. error max(2^-22 |x|, 2^-25))."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import weights, mapnet_oracle as O

stats = []
orig = O._conv
def conv(x, w, stride, pad, emulate):
    y = orig(x, w, stride, pad, emulate)
    idx = len(stats)
    stats.append(None)
    def hook(g, idx=idx, shape=tuple(y.shape)):
        a = g.abs()
        nz = a[a > 0]
        stats[idx] = (shape, float(a.max()), float((g * g).mean().sqrt()), float(nz.min()) if nz.numel() else 0.0,
                      float((a > 0).float().mean()))
    y.register_hook(hook)
    return y
O._conv = conv
name = sys.argv[1] if len(sys.argv) > 1 else "b8"
cfg = {"b8": dict(kind="posenet", N=8, H=256, W=256), "tiny": dict(kind="posenet", N=4, H=64, W=64),
       "mapnet": dict(kind="mapnet", N=4, T=3, H=128, W=128)}[name]
st = weights.make_state(7)
x, targ = weights.make_inputs(cfg, 7)
r = O.train_step(cfg["kind"], st, x, targ, dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0), do_step=False)
print("loss", float(r["loss"]))
amax_all = max(s[1] for s in stats)
for i, s in enumerate(stats):
    print("conv %2d %-22s amax %.3e rms %.3e  amax/rms %6.1f  rms/amax_all %.2e  minnz %.1e nz %.2f" %
          (i, s[0], s[1], s[2], s[1] / s[2], s[2] / amax_all, s[3], s[4]))
print("global amax %.3e; rms range %.3e .. %.3e (ratio %.1f)" % (amax_all, min(s[2] for s in stats), max(s[2] for s in stats),
      max(s[2] for s in stats) / min(s[2] for s in stats)))
