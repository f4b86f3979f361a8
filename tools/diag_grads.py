"""Diagnostic (GPU box): per-tensor gradient error of the product vs the CPU oracle."""
import sys, os, ast
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
torch.set_num_threads(16)
from oracle import weights, mapnet_oracle as O
from helpers import load_golden, make_product_model, make_product_criterion, product_step

name = sys.argv[1]; prec = sys.argv[2]; dt = sys.argv[3] if len(sys.argv) > 3 else "f32"
g, cfg = load_golden(name)
st = weights.make_state(7); x, targ = weights.make_inputs(cfg, 7)
sv = dict(sax=0.0, saq=-3.0, srx=0.0, srq=-3.0)
if dt == "f64":
    st_o = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in st.items()}
    r = O.train_step(cfg["kind"], st_o, x.double(), targ.double(), sv, do_step=False)
else:
    r = O.train_step(cfg["kind"], st, x, targ, sv, do_step=False)
model, net = make_product_model(st, cfg["kind"], prec)
crit = make_product_criterion(cfg["kind"])
model.train()
loss, pred, grads, sg = product_step(model, net, crit, x, targ, do_step=False)
print(name, prec, "loss", float(loss), float(r["loss"]), "pred relerr", float((pred.cpu().double() - r["pred"].double().view_as(pred.cpu())).abs().max() / r["pred"].abs().max()))
for n, t in r["grads"].items():
    a = grads[n].double().cpu(); b = t.double()
    print("%-52s ref_norm %.3e  relL2 %.2e  normratio %.5f  cos %.6f" % (n, float(b.norm()), float((a - b).norm() / b.norm()), float(a.norm() / b.norm()), float((a * b).sum() / (a.norm() * b.norm()))))
