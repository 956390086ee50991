"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the loss-kernel bindings of contrad_b200.kernels (rownorm,
contrastive, GAN losses), so that the host logic of the StyleGAN2 step can be exercised without a GPU.  The
definitions follow include/contrad_b200.h; backward passes come from autograd of the forward definitions."""
import contextlib

import torch
import torch.nn.functional as F


def rownorm_fwd(x, eps=1e-12):
    n = x.norm(dim=1).clamp_min(eps)
    return x / n[:, None], 1.0 / n


def rownorm_bwd(dy, y, inv, out=None, round_out=False):
    dot = (dy * y).sum(1, keepdim=True)
    return (dy - y * dot) * inv[:, None]


def _contrastive(z, n, mode, temperature):
    sim = z @ z.t() / temperature
    sim = sim - torch.diag(torch.diagonal(sim)) + torch.diag(torch.full((z.shape[0],), -5e4))
    lsm = F.log_softmax(sim, dim=1)
    if mode == 0:
        return -(lsm[:n, n:2 * n].diag() + lsm[n:2 * n, :n].diag()).sum() / (2 * n)
    rows = lsm[2 * n:3 * n]
    mask = torch.zeros_like(rows)
    mask[:, 2 * n:] = 1.0
    mask[torch.arange(n), 2 * n + torch.arange(n)] = 0.0
    mask = mask / mask.sum(1, keepdim=True)
    return -(mask * rows).sum(1).mean()


def contrastive_fwd(z, n, mode, temperature):
    return _contrastive(z, n, mode, temperature).reshape(1), torch.zeros(1)


def contrastive_bwd(z, n, mode, temperature, lse, gscale):
    zz = z.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        (dz,) = torch.autograd.grad(_contrastive(zz, n, mode, temperature), zz)
    return dz * gscale.reshape(())


def gan_d_loss(d_real, d_gen, kind, g_real=None, g_gen=None):
    assert kind == "nonsat"
    dr = d_real.detach().clone().requires_grad_(True)
    dg = d_gen.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = F.softplus(dg).mean() + F.softplus(-dr).mean()
        g_r, g_g = torch.autograd.grad(loss, [dr, dg])
    if g_real is not None:
        g_r = g_real.copy_(g_r)
    if g_gen is not None:
        g_g = g_gen.copy_(g_g)
    return torch.stack([loss.detach(), d_real.mean(), d_gen.mean()]), g_r, g_g


def gan_g_loss(d_gen, kind):
    assert kind == "nonsat"
    dg = d_gen.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = F.softplus(-dg).mean()
        (g,) = torch.autograd.grad(loss, dg)
    return loss.detach().reshape(1), g


NAMES = ("rownorm_fwd", "rownorm_bwd", "contrastive_fwd", "contrastive_bwd", "gan_d_loss", "gan_g_loss")


@contextlib.contextmanager
def patched():
    import sys
    from contrad_b200 import kernels as K
    me = sys.modules[__name__]
    saved = [(n, getattr(K, n)) for n in NAMES]
    try:
        for n, _ in saved:
            setattr(K, n, getattr(me, n))
        yield
    finally:
        for n, fn in saved:
            setattr(K, n, fn)
