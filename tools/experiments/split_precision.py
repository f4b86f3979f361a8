"""CPU experiment: which operand split lets a tensor-core conv meet 1e-4 on the pose?
Emulates the conv arithmetic of candidate schemes inside the oracle's forward (fp64 products/accumulate of the
rounded operand planes, so only operand representation + dropped cross terms are modelled) and compares the
pose / loss against the fp64 run of the same graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import weights, mapnet_oracle as O

def split(t, dt, n):
    planes, r = [], t.clone()
    for _ in range(n):
        p = r.to(dt).to(t.dtype)
        planes.append(p); r = r - p
    return planes

def make_conv(scheme):
    def conv(x, w, stride, pad, emulate):
        if scheme == "fp32":
            return F.conv2d(x, w, None, stride, pad)
        dt = torch.bfloat16 if scheme.startswith("bf16") else torch.float16
        n = 3 if scheme.endswith("x6") else 2
        xs, ws = split(x, dt, n), split(w, dt, n)
        if scheme.endswith("x3"): pairs = [(0, 0), (0, 1), (1, 0)]
        elif scheme.endswith("x4"): pairs = [(0, 0), (0, 1), (1, 0), (1, 1)]
        elif scheme.endswith("x6"): pairs = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
        elif scheme.endswith("x1"): pairs = [(0, 0)]
        y = 0
        for a, b in reversed(pairs):
            y = y + F.conv2d(xs[a].double(), ws[b].double(), None, stride, pad)
        return y.to(x.dtype)
    return conv

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "b8"
    cfg = dict(kind="posenet", N=8, H=256, W=256) if name == "b8" else dict(kind="posenet", N=4, H=64, W=64)
    st = weights.make_state(7)
    x, targ = weights.make_inputs(cfg, 7)
    sv = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)
    st64 = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in st.items()}
    with torch.no_grad():
        p64 = O.posenet_forward(st64, x.double(), training=True)
        l64 = O.posenet_criterion(p64, targ.double(), torch.tensor([0.0]).double(), torch.tensor([-3.0]).double())
    orig = O._conv
    for scheme in sys.argv[2:] or ["fp32", "bf16x3", "bf16x4", "fp16x3", "bf16x6"]:
        O._conv = make_conv(scheme)
        t0 = time.time()
        with torch.no_grad():
            p = O.posenet_forward(st, x, training=True)
            l = O.posenet_criterion(p, targ, torch.tensor([0.0]), torch.tensor([-3.0]))
        ep = float((p.double() - p64).abs().max() / p64.abs().max())
        el = abs(float(l) - float(l64)) / abs(float(l64))
        print("%-8s pose err %.3e  loss err %.3e  (%.1fs)" % (scheme, ep, el, time.time() - t0), flush=True)
    O._conv = orig

main()
