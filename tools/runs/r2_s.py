import sys, os
sys.path.insert(0, "/root/repo"); sys.path.append("/root/repo/contrad_b200/compat")
from types import SimpleNamespace
import torch
from oracle import stylegan2_oracle as SO
from contrad_b200 import precision
from contrad_b200.models.gan import get_architecture
from contrad_b200.training.gan import stylegan2 as T
sys.path.insert(0, "/root/repo/tests")
size = 32
def states():
    sd_d = SO.make_d_state(size, small32=True, d_hidden=512, generator=torch.Generator().manual_seed(41))
    gen = torch.Generator().manual_seed(43)
    for k in sd_d:
        if k.endswith(".bias") and sd_d[k].abs().sum() == 0:
            sd_d[k] = 0.1 * torch.randn(sd_d[k].shape, generator=gen)
    return sd_d
for n in (16, 64):
    sd_d = states()
    torch.manual_seed(44)
    real2, fake = torch.rand(2 * n, 3, size, size), torch.rand(n, 3, size, size)
    leaf = {k: (v.clone().requires_grad_(True) if not k.endswith(".kernel") else v) for k, v in sd_d.items()}
    d_loss_o, pen_o, _, _ = SO.gd_losses(leaf, size, real2, fake)
    r1_o = SO.r1_penalty(leaf, real2[:n], size).mean()
    parts_o = {}
    for name, term in (("con", d_loss_o), ("dis", pen_o), ("r1", 0.05 * r1_o)):
        gs = torch.autograd.grad(term, [v for v in leaf.values() if v.requires_grad], retain_graph=True, allow_unused=True)
        parts_o[name] = float(torch.stack([g.double().pow(2).sum() for g in gs if g is not None]).sum().sqrt())
    for mode in (0, "full"):
        _, D = get_architecture("stylegan2", (size, size, 3))
        D.load_state_dict(sd_d, strict=True); D.cuda().train()
        with precision.strict(mode):
            d_all, view_r, view_f = T.discriminate(D, real2.cuda(), fake.cuda())
            P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
            d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
            r1 = T.r1_loss(D, real2[:n].cuda(), lambda t: t)
            out = {}
            for name, term in (("con", d_loss), ("dis", aux["penalty"]), ("r1", 0.05 * r1)):
                gs = torch.autograd.grad(term, list(D.parameters()), retain_graph=True, allow_unused=True)
                out[name] = float(torch.stack([g.double().pow(2).sum() for g in gs if g is not None]).sum().sqrt())
        print("n=%d mode=%s" % (n, mode), {k: "%.2e (ref %.4g)" % (abs(out[k] - parts_o[k]) / parts_o[k], parts_o[k]) for k in out},
              "d_loss %.2e r1 %.2e" % (abs(float(d_loss) - float(d_loss_o)) / float(d_loss_o), abs(float(r1) - float(r1_o)) / float(r1_o)), flush=True)
