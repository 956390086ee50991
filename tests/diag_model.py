"""Diagnostic (test infrastructure, run by hand on a GPU box: python tests/diag_model.py): product D (kernels) vs the oracle on
cuda fp32 (TF32 off) at the module-output level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.append(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat"))
import torch, numpy as np
from oracle import contrad_oracle as O
from contrad_b200.models.gan import get_architecture
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for B in (18, 12, 96):
    gen_w = torch.Generator().manual_seed(99)
    sd_d = O.make_d_state(generator=gen_w)
    G, D = get_architecture("sndcgan", (32, 32, 3)); D.load_state_dict(sd_d); D.cuda().train()
    sd_o = {k: v.clone().cuda() for k, v in sd_d.items()}
    O.set_requires_grad(sd_o, True)
    torch.manual_seed(B)
    x = torch.rand(B, 3, 32, 32, device="cuda")
    xo = x.clone().requires_grad_(True); xp = x.clone().requires_grad_(True)
    out_o, aux_o = O.d_sndcgan_forward(sd_o, xo, sg_linear=True)
    out_p, aux_p = D(xp, sg_linear=True, projection=True, projection2=True, penultimate=True)
    print("B=%d fwd rel err: d %.2e  proj %.2e  proj2 %.2e  feat %.2e" % (B, rel(out_p, out_o), rel(aux_p["projection"], aux_o["projection"]), rel(aux_p["projection2"], aux_o["projection2"]), rel(aux_p["penultimate"], aux_o["penultimate"])))
    for which in ("d", "projection", "projection2", "all"):
        gd = torch.randn_like(out_o); g1 = torch.randn_like(aux_o["projection"]); g2 = torch.randn_like(aux_o["projection2"])
        if which == "d": g1 = g1 * 0; g2 = g2 * 0
        if which == "projection": gd = gd * 0; g2 = g2 * 0
        if which == "projection2": gd = gd * 0; g1 = g1 * 0
        for v in O.trainable(sd_o).values(): v.grad = None
        xo.grad = None
        ((out_o * gd).sum() + (aux_o["projection"] * g1).sum() + (aux_o["projection2"] * g2).sum()).backward(retain_graph=True)
        D.zero_grad(); xp.grad = None
        ((out_p * gd).sum() + (aux_p["projection"] * g1).sum() + (aux_p["projection2"] * g2).sum()).backward(retain_graph=True)
        named = dict(D.named_parameters())
        errs = {}
        for k, v in O.trainable(sd_o).items():
            if v.grad is None or float(v.grad.norm()) == 0: continue
            g = named[k].grad
            errs[k] = rel(g, v.grad) if g is not None else float("nan")
        print("  upstream=%-11s dx err %.2e |" % (which, rel(xp.grad, xo.grad)), " ".join("%s:%.1e" % (k.replace("weight_orig", "w").replace("projection", "pj"), e) for k, e in errs.items()))
