"""Generate tests/golden/stylegan2_small.pt + stylegan2_keys.json by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, imported through oracle/ref_import.py; its two CUDA extensions are JIT-built by its own
`models/gan/stylegan2/op/__init__.py`, the CPU code path `upfirdn2d_native` is what executes here) with fixed seeds.

    TORCH_CUDA_ARCH_LIST=10.0 python tests/golden/make_golden_sg2.py

The StyleGAN2 `stylegan2` (small32) networks have 20+ M parameters each, so the fixture does not store weights: they
are re-drawn from a seeded generator by oracle.stylegan2_oracle.make_{d,g}_state (same keys / shapes as the reference
modules - checked here by load_state_dict(strict=True)) and only inputs, explicit random draws and the reference's
outputs / gradients are stored.
"""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_import                  # noqa: E402
from oracle import stylegan2_oracle as SO      # noqa: E402

SIZE = 32
W_SEED_D, W_SEED_G = 101, 202


def grad_norms(named_params):
    return {k: float(p.grad.norm()) for k, p in named_params if p.grad is not None}


def main():
    ref_import.activate()
    from models.gan import get_architecture
    from models.gan.stylegan2.op import upfirdn2d as ref_upfirdn2d
    from training.criterion import nt_xent
    from training.gan.contrad import supcon_fake

    out = {"size": SIZE, "w_seed_d": W_SEED_D, "w_seed_g": W_SEED_G}
    G, D = get_architecture("stylegan2", (SIZE, SIZE, 3))
    keys = {"D": {k: list(v.shape) for k, v in D.state_dict().items()},
            "G": {k: list(v.shape) for k, v in G.state_dict().items()}}
    G512, D512 = get_architecture("stylegan2_512", (512, 512, 3))
    keys["D512"] = {k: list(v.shape) for k, v in D512.state_dict().items()}
    keys["G512"] = {k: list(v.shape) for k, v in G512.state_dict().items()}
    del G512, D512
    with open(os.path.join(HERE, "stylegan2_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)

    sd_d = SO.make_d_state(SIZE, small32=True, d_hidden=512, generator=torch.Generator().manual_seed(W_SEED_D))
    sd_g = SO.make_g_state(SIZE, small32=True, generator=torch.Generator().manual_seed(W_SEED_G))
    # non-trivial biases / noise weights so that every term of the forward is exercised (the reference initialises
    # them to zero)
    gen = torch.Generator().manual_seed(303)
    for sd in (sd_d, sd_g):
        for k in sd:
            if k.endswith(".bias") and sd[k].abs().sum() == 0:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
            if k.endswith("noise.weight"):
                sd[k] = 0.1 * torch.randn(1, generator=gen)
    out["bias_seed"] = 303
    D.load_state_dict(sd_d, strict=True)
    G.load_state_dict(sd_g, strict=True)
    D.train(); G.train()

    # ---- upfirdn2d cases (CPU path of the reference op)
    torch.manual_seed(1)
    cases = []
    for (h, w, k, up, down, pad) in ((9, 7, (1, 3, 3, 1), 1, 1, (2, 2)), (8, 8, (1, 3, 3, 1), 1, 2, (1, 1)),
                                     (5, 6, (1, 3, 3, 1), 2, 1, (2, 1)), (9, 9, (1, 3, 3, 1), 1, 1, (1, 1)),
                                     (6, 5, (1, 2, 1), 2, 3, (0, 3)), (8, 8, (1, 3, 3, 1), 1, 1, (-1, 2))):
        x = torch.randn(2, 3, h, w, requires_grad=True)
        kern = SO.make_kernel(k) * (up ** 2)
        y = ref_upfirdn2d(x, kern, up=up, down=down, pad=pad)
        dy = torch.randn_like(y)
        (dx,) = torch.autograd.grad(y, x, dy)
        cases.append({"x": x.detach(), "kernel": kern, "up": up, "down": down, "pad": pad, "y": y.detach(), "dy": dy,
                      "dx": dx})
    out["upfirdn2d"] = cases

    # ---- discriminator forward / backward
    torch.manual_seed(21)
    B = 4
    x = torch.rand(B, 3, SIZE, SIZE)
    c_d, c1, c2 = torch.randn(B, 1), torch.randn(B, 128), torch.randn(B, 128)
    xr = x.clone().requires_grad_(True)
    D.zero_grad()
    d, aux = D(xr, projection=True, projection2=True, penultimate=True)
    loss = (d * c_d).sum() + (aux["projection"] * c1).sum() + (aux["projection2"] * c2).sum()
    loss.backward()
    out["d_case"] = {"x": x, "c_d": c_d, "c1": c1, "c2": c2, "d": d.detach(), "projection": aux["projection"].detach(),
                     "projection2": aux["projection2"].detach(), "penultimate": aux["penultimate"].detach(),
                     "dx": xr.grad.clone(), "grad_norms": grad_norms(D.named_parameters()),
                     "grad_from_rgb": D.layers[0][0].weight.grad.clone(),
                     "grad_last_bias": D.last_conv[1].bias.grad.clone()}
    # sg_linear: the `linear` head must not reach the backbone
    D.zero_grad()
    d, aux = D(x, projection=True, sg_linear=True)
    (d * c_d).sum().backward()
    out["d_case"]["sg_linear_backbone_grad_is_none_or_zero"] = bool(
        D.layers[0][0].weight.grad is None or float(D.layers[0][0].weight.grad.abs().sum()) == 0.0)

    # ---- R1 penalty (double backward)
    from train_stylegan2 import r1_loss

    class _Id(torch.nn.Module):
        def forward(self, t):
            return t

    D.zero_grad()
    xr = x.clone()
    r1_mean = r1_loss(D, xr, _Id())
    r1_mean.backward()
    images_aug = x.clone().requires_grad_(True)
    d_real = D(images_aug)
    (g_real,) = torch.autograd.grad(d_real.sum(), images_aug, create_graph=True)
    per_sample = g_real.pow(2).reshape(B, -1).sum(1)
    out["r1_case"] = {"x": x, "r1_mean": float(r1_mean), "per_sample": per_sample.detach(), "grad_x": g_real.detach(),
                      "grad_norms": grad_norms(D.named_parameters()),
                      "grad_from_rgb": D.layers[0][0].weight.grad.clone(),
                      "grad_conv1_bias": D.layers[1].conv1[1].bias.grad.clone()}

    # ---- generator (train mode: style mixing + explicit noises)
    torch.manual_seed(31)
    z = torch.randn(B, 512)
    noises = [torch.randn(*s) for s in SO.noise_shapes(SIZE, B)]
    c_img = torch.randn(B, 3, SIZE, SIZE)
    torch.manual_seed(32)
    z_mix = torch.randn(B, 512)                       # the draw Generator.sample_latent makes first (generator.py:254)
    state_after = torch.get_rng_state()
    nomix = torch.rand(B) >= 0.9
    mix_layer = torch.randint(SO.n_latent_for(SIZE), (B,)).masked_fill(nomix, SO.n_latent_for(SIZE))
    torch.manual_seed(32)
    G.zero_grad()
    img, latents = G(z, return_latents=True, style_mix=0.9, noise=noises)
    (img * c_img).sum().backward()
    out["g_case"] = {"z": z, "noises": noises, "c_img": c_img, "z_mix": z_mix, "mix_layer": mix_layer,
                     "rng_state_after_zmix": state_after, "mix_seed": 32, "image": img.detach(),
                     "latents": latents.detach(), "grad_norms": grad_norms(G.named_parameters()),
                     "grad_const": G.input.const.grad.clone(),
                     "grad_noise_w": torch.stack([G.conv1.noise.weight.grad] + [l.noise.weight.grad for l in G.layers]).clone(),
                     "grad_rgb_bias": G.to_rgbs[-1].bias.grad.clone()}
    # no style mixing
    G.zero_grad()
    img0 = G(z, style_mix=0.0, noise=noises)
    out["g_case"]["image_nomix"] = img0.detach()

    # ---- D-step losses of train_stylegan2_contraD.py (G_D.forward + _loss_D_fn) on given "augmented" batches
    torch.manual_seed(41)
    n = 4
    fake_aug = torch.rand(n, 3, SIZE, SIZE)
    real_aug2 = torch.rand(2 * n, 3, SIZE, SIZE)
    D.zero_grad()
    d_gen, aux_f = D(fake_aug, sg_linear=True, projection=True, projection2=True)
    d_rs, aux_r = D(real_aug2, sg_linear=True, projection=True, projection2=True)
    views_r, reals = F.normalize(aux_r["projection"]), F.normalize(aux_r["projection2"])
    others, fakes = F.normalize(aux_f["projection"]), F.normalize(aux_f["projection2"])
    simclr = nt_xent(views_r[:n], views_r[n:], temperature=0.1)
    sup = supcon_fake(reals[:n], reals[n:], fakes, temperature=0.1)
    d_real = d_rs[:n]
    penalty = F.softplus(d_gen).mean() + F.softplus(-d_real).mean()
    (simclr + sup + penalty).backward()
    out["dstep_case"] = {"fake_aug": fake_aug, "real_aug2": real_aug2, "d_loss": float(simclr + sup),
                         "penalty": float(penalty), "d_real": float(d_real.mean()), "d_gen": float(d_gen.mean()),
                         "grad_norms": grad_norms(D.named_parameters())}
    D.zero_grad()
    g_l = F.softplus(-D(fake_aug)).mean()
    out["dstep_case"]["g_loss"] = float(g_l)

    # ---- the same D-step at n = 16 (inputs + scalars + gradient norms only): more samples per parameter gradient
    torch.manual_seed(51)
    n = 16
    fake_aug = torch.rand(n, 3, SIZE, SIZE).half().float()          # stored as fp16: quantise before use
    real_aug2 = torch.rand(2 * n, 3, SIZE, SIZE).half().float()
    D.zero_grad()
    d_gen, aux_f = D(fake_aug, sg_linear=True, projection=True, projection2=True)
    d_rs, aux_r = D(real_aug2, sg_linear=True, projection=True, projection2=True)
    views_r, reals = F.normalize(aux_r["projection"]), F.normalize(aux_r["projection2"])
    others, fakes = F.normalize(aux_f["projection"]), F.normalize(aux_f["projection2"])
    simclr = nt_xent(views_r[:n], views_r[n:], temperature=0.1)
    sup = supcon_fake(reals[:n], reals[n:], fakes, temperature=0.1)
    d_real = d_rs[:n]
    penalty = F.softplus(d_gen).mean() + F.softplus(-d_real).mean()
    r1 = r1_loss(D, real_aug2[:n], _Id())
    (simclr + sup + penalty + 0.05 * r1).backward()
    total = math.sqrt(sum(float(p.grad.double().pow(2).sum()) for p in D.parameters() if p.grad is not None))
    out["dstep16_case"] = {"fake_aug": fake_aug.half(), "real_aug2": real_aug2.half(), "d_loss": float(simclr + sup),
                           "penalty": float(penalty), "r1": float(r1), "d_real": float(d_real.mean()),
                           "d_gen": float(d_gen.mean()), "grad_norms": grad_norms(D.named_parameters()),
                           "total_grad_norm": total}

    torch.save(out, os.path.join(HERE, "stylegan2_small.pt"))
    print("wrote stylegan2_small.pt (%.1f KB)" % (os.path.getsize(os.path.join(HERE, "stylegan2_small.pt")) / 1024))


if __name__ == "__main__":
    main()
