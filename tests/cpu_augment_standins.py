"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the augmentation bindings of contrad_b200.kernels (rows f3 / f4) and
for the GAN-loss kernels of every loss kind, so that the HOST logic around them (random-draw order, view bookkeeping,
autograd wiring, the baseline training modes) can be exercised without a GPU.  The arithmetic is the oracle's, which is
pinned on the reference (tests/test_oracle_golden.py)."""
import contextlib

import torch

from oracle import contrad_oracle as O


def _unpack(params):
    return {k: params[i] for i, k in enumerate(O.PARAM_FIELDS)}


def augment_simclr_fwd(x, params, order):
    return O.augment_simclr(x, _unpack(params), order).detach()


def augment_simclr_bwd(x, dy, params, order):
    xx = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        (dx,) = torch.autograd.grad(O.augment_simclr(xx, _unpack(params), order), xx, dy)
    return dx


def augment_needs_large_path(H, W):
    return False


def augment_simclr_mixed_fwd(x_u8, n_u8_views, x_f32, params, order):
    parts = []
    if n_u8_views:
        reps, rem = divmod(n_u8_views, x_u8.shape[0])
        assert rem == 0
        parts += [O.to_tensor_u8(x_u8)] * reps
    if x_f32 is not None and x_f32.shape[0]:
        parts.append(x_f32.detach())
    y = O.augment_simclr(torch.cat(parts, dim=0), _unpack(params), order)
    return y, torch.zeros(y.shape[0], 3)


def shift_flip(x, params, padding_mode, adjoint=False):
    if not adjoint:
        return O.shift_flip(x, params, padding_mode)
    probe = torch.zeros_like(x, requires_grad=True)
    with torch.enable_grad():
        (dx,) = torch.autograd.grad(O.shift_flip(probe, params, padding_mode), probe, x)
    return dx


def noise_clamp_fwd(x, noise, sigma):
    return O.gaussian_noise(x, noise, sigma)


def noise_clamp_bwd(x, noise, dy, sigma):
    u = x + noise * sigma
    return dy * ((u >= 0) & (u <= 1)).to(dy.dtype)


_DIFFAUG_STAGES = ("color", "translation", "cutout")


def diffaug(x, params, flags, adjoint=False):
    stages = tuple(s for k, s in enumerate(_DIFFAUG_STAGES) if flags & (1 << k))
    if not adjoint:
        return O.diffaug(x, params, stages)
    probe = torch.zeros_like(x, requires_grad=True)
    with torch.enable_grad():
        (dx,) = torch.autograd.grad(O.diffaug(probe, params, stages), probe, x)
    return dx


def gan_d_loss(d_real, d_gen, kind, g_real=None, g_gen=None):
    dr = d_real.detach().clone().requires_grad_(True)
    dg = d_gen.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = O.gan_d_loss(dr, dg, kind)
        g_r, g_g = torch.autograd.grad(loss, [dr, dg])
    if g_real is not None:
        g_r = g_real.copy_(g_r)
    if g_gen is not None:
        g_g = g_gen.copy_(g_g)
    return torch.stack([loss.detach(), d_real.mean(), d_gen.mean()]), g_r, g_g


def gan_g_loss(d_gen, kind):
    dg = d_gen.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = O.gan_g_loss(dg, kind)
        (g,) = torch.autograd.grad(loss, dg)
    return loss.detach().reshape(1), g


NAMES = ("augment_simclr_fwd", "augment_simclr_bwd", "augment_needs_large_path", "augment_simclr_mixed_fwd", "shift_flip",
         "noise_clamp_fwd", "noise_clamp_bwd", "diffaug", "gan_d_loss", "gan_g_loss")


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
