"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, imported through oracle/ref_import.py) on CPU with fixed seeds.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt / *.json

The reference cannot travel to the GPU box, the fixtures can.  Each fixture stores the inputs,
the augmentation parameters (drawn by oracle.sample_simclr_params from the same seeds -- the
reference's own draws are not observable, so agreement of the *outputs* is what pins the
sampler) and the reference's outputs.  tests/test_oracle_golden.py replays them against the
oracle; the GPU parity tests replay them against the CUDA path.
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_import            # noqa: E402
from oracle import contrad_oracle as O   # noqa: E402


def seed_all(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)


def t2l(t):
    return t.detach().clone()


def gen_augment(gin):
    from augment import get_augment
    cases = []
    want_orders = {0, 1}
    specs = [(5, 32), (5, 32), (1, 64), (3, 16)]
    seed = 0
    for batch, size in specs:
        while True:
            seed += 1
            seed_all(seed)
            x = torch.rand(batch, 3, size, size)
            dy = torch.randn(batch, 3, size, size)
            params, order = O.sample_simclr_params(batch, size, size)
            if size == 32 and order not in want_orders:
                continue
            want_orders.discard(order)
            break
        aug = get_augment(mode="simclr")
        seed_all(seed)
        x_ref = torch.rand(batch, 3, size, size)
        _ = torch.randn(batch, 3, size, size)
        assert torch.equal(x, x_ref)
        x_ref.requires_grad_(True)
        y_ref = aug(x_ref)
        (y_ref * dy).sum().backward()
        cases.append({"seed": seed, "x": t2l(x), "dy": t2l(dy), "order": order,
                      "params": O.pack_params(params), "y": t2l(y_ref), "dx": t2l(x_ref.grad)})
        print("augment case B=%d size=%d seed=%d order=%d" % (batch, size, seed, order))
    torch.save({"cases": cases, "fields": list(O.PARAM_FIELDS)}, os.path.join(HERE, "augment_simclr.pt"))


def gen_contrastive():
    from training.criterion import nt_xent
    from training.gan.contrad import supcon_fake
    import torch.nn.functional as F
    cases = []
    for n, seed in ((8, 11), (16, 12), (5, 13)):
        seed_all(seed)
        a = F.normalize(torch.randn(n, 128)).requires_grad_(True)
        b = F.normalize(torch.randn(n, 128)).requires_grad_(True)
        c = F.normalize(torch.randn(n, 128)).requires_grad_(True)
        l1 = nt_xent(a, b, temperature=0.1)
        g1 = torch.autograd.grad(l1, [a, b])
        l2 = supcon_fake(a, b, c, temperature=0.1)
        g2 = torch.autograd.grad(l2, [a, b, c])
        l3 = nt_xent(a, b, temperature=0.5)
        cases.append({"n": n, "a": t2l(a), "b": t2l(b), "c": t2l(c),
                      "nt_xent": float(l1), "nt_xent_grads": [t2l(g) for g in g1],
                      "supcon": float(l2), "supcon_grads": [t2l(g) for g in g2],
                      "nt_xent_t05": float(l3)})
    torch.save({"cases": cases}, os.path.join(HERE, "contrastive.pt"))
    print("contrastive cases:", [(c["n"], c["nt_xent"], c["supcon"]) for c in cases])


def _grad_norms(module):
    return {k: float(p.grad.double().norm()) for k, p in module.named_parameters() if p.grad is not None}


def gen_small_models(gin):
    """Reference D_SNDCGAN / G_SNDCGAN at reduced width (ndf=4, d_hidden=16; ngf=4, nz=16) with the
    reference's own initialisation; the state_dicts are stored."""
    from augment import get_augment
    from models.gan.sndcgan import D_SNDCGAN, G_SNDCGAN
    from training.gan import contrad as ref_contrad
    out = {}
    for loss_kind, seed in (("nonsat", 21), ("hinge", 22)):
        seed_all(seed)
        D = D_SNDCGAN(image_size=(32, 32, 3), ndf=4, mlp_linear=True, d_hidden=16)
        G = G_SNDCGAN(image_size=(32, 32, 3), ngf=4, nz=16)
        D.train(); G.train()
        sd_d0 = {k: t2l(v) for k, v in D.state_dict().items()}
        sd_g0 = {k: t2l(v) for k, v in G.state_dict().items()}
        n = 4
        P = SimpleNamespace(augment_fn=get_augment(mode="simclr"), temp=0.1, lbd_a=1.0, distributed=False)
        options = {"loss": loss_kind}

        seed_all(seed + 100)
        images = torch.rand(n, 3, 32, 32)
        z_d = G.sample_latent(n)
        # ---- D step (train_gan.py:152-159)
        for p in G.parameters(): p.requires_grad = False
        for p in D.parameters(): p.requires_grad = True
        with torch.no_grad():
            gen = G(z_d)
        d_loss, aux = ref_contrad.loss_D_fn(P, D, options, images, gen)
        (d_loss + aux["penalty"]).backward()
        d_rec = {"l_con": float(d_loss), "l_dis": float(aux["penalty"]), "d_real": float(aux["d_real"]),
                 "d_gen": float(aux["d_gen"]), "grad_norms": _grad_norms(D), "gen": t2l(gen)}
        u_after = {k: t2l(v) for k, v in D.state_dict().items() if k.endswith(("weight_u", "weight_v"))}
        # ---- G step (train_gan.py:169-175), D frozen, fresh latent
        for p in G.parameters(): p.requires_grad = True
        for p in D.parameters(): p.requires_grad = False
        D.zero_grad()
        z_g = G.sample_latent(n)
        gen2 = G(z_g)
        g_loss = ref_contrad.loss_G_fn(P, D, options, images, gen2)
        g_loss.backward()
        g_rec = {"l_gen": float(g_loss), "grad_norms": _grad_norms(G)}

        # replay the RNG stream with the oracle's sampler to obtain the explicit draws
        seed_all(seed + 100)
        images2 = torch.rand(n, 3, 32, 32)
        z_d2 = O.sample_latent(n, 16)
        aug_d, order_d = O.sample_simclr_params(3 * n, 32, 32)
        z_g2 = O.sample_latent(n, 16)
        aug_g, order_g = O.sample_simclr_params(n, 32, 32)
        assert torch.equal(images, images2) and torch.equal(z_d, z_d2) and torch.equal(z_g, z_g2)
        out[loss_kind] = {
            "sd_d": sd_d0, "sd_g": sd_g0, "images": t2l(images), "z_d": t2l(z_d), "z_g": t2l(z_g),
            "aug_d": O.pack_params(aug_d), "order_d": order_d, "aug_g": O.pack_params(aug_g), "order_g": order_g,
            "d_step": d_rec, "uv_after_d_step": u_after, "g_step": g_rec,
        }
        print("small models [%s]: l_con=%.6f l_dis=%.6f l_gen=%.6f" % (loss_kind, d_rec["l_con"], d_rec["l_dis"], g_rec["l_gen"]))
    torch.save(out, os.path.join(HERE, "sndcgan_small.pt"))


def gen_config1(gin):
    """BASELINE config 1: SNDCGAN+ContraD c10_b512.gin hyper-parameters, batch 64, synthetic 32x32,
    two full steps of the reference loop (train_gan.py:141-179) on CPU.  Weights come from
    oracle.make_d_state/make_g_state(seed) loaded into the reference modules, so only scalars are stored."""
    from augment import get_augment
    from models.gan import get_architecture
    from training.gan import contrad as ref_contrad
    import torch.optim as optim
    gin.parse_config_file(os.path.join(ref_import.REFERENCE_ROOT, "configs/gan/cifar10/c10_b512.gin"))
    n = 64
    gen_w = torch.Generator().manual_seed(1234)
    sd_d = O.make_d_state(generator=gen_w)
    sd_g = O.make_g_state(generator=gen_w)
    G, D = get_architecture("sndcgan", (32, 32, 3))
    D.load_state_dict(sd_d); G.load_state_dict(sd_g)
    opt_G = optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_D = optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    P = SimpleNamespace(augment_fn=get_augment(mode="simclr"), temp=0.1, lbd_a=1.0, distributed=False)
    options = {"loss": "nonsat", "warmup": 3000, "lr": 2e-4, "lr_d": 2e-4}
    seed_all(7)
    records = []
    for step in (1, 2):
        G.train(); D.train()
        for opt in (opt_G, opt_D):
            ratio = min(1., (step + 1) / options["warmup"])
            for pg in opt.param_groups:
                pg["lr"] = ratio * options["lr"]
        for p in G.parameters(): p.requires_grad = False
        for p in D.parameters(): p.requires_grad = True
        images = torch.rand(n, 3, 32, 32)
        with torch.no_grad():
            gen = G(G.sample_latent(n))
        d_loss, aux = ref_contrad.loss_D_fn(P, D, options, images, gen)
        loss = d_loss + aux["penalty"]
        opt_D.zero_grad(); loss.backward()
        d_gn = float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in D.parameters() if p.grad is not None)))
        opt_D.step()
        for p in G.parameters(): p.requires_grad = True
        for p in D.parameters(): p.requires_grad = False
        gen = G(G.sample_latent(n))
        g_loss = ref_contrad.loss_G_fn(P, D, options, images, gen)
        opt_G.zero_grad(); g_loss.backward()
        g_gn = float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in G.parameters() if p.grad is not None)))
        opt_G.step()
        rec = {"step": step, "l_con": float(d_loss), "l_dis": float(aux["penalty"]), "d_real": float(aux["d_real"]),
               "d_gen": float(aux["d_gen"]), "l_gen": float(g_loss), "d_grad_norm": d_gn, "g_grad_norm": g_gn}
        print("config1", rec)
        records.append(rec)
    with open(os.path.join(HERE, "config1_scalars.json"), "w") as f:
        json.dump({"weights_seed": 1234, "data_seed": 7, "batch": n, "steps": records,
                   "note": "reference modules, torch %s CPU fp32" % torch.__version__}, f, indent=1)


def gen_spectral_norm():
    from torch.nn.utils import spectral_norm
    import torch.nn as nn
    seed_all(31)
    conv = spectral_norm(nn.Conv2d(6, 10, 3, 1, 1))
    lin = spectral_norm(nn.Linear(40, 7))
    out = {}
    for name, m in (("conv", conv), ("lin", lin)):
        m.train()
        rec = {"weight_orig": t2l(m.weight_orig), "bias": t2l(m.bias), "u0": t2l(m.weight_u), "v0": t2l(m.weight_v)}
        x = torch.randn(2, 6, 5, 5) if name == "conv" else torch.randn(3, 40)
        y = m(x)
        y.pow(2).sum().backward()
        rec.update({"x": t2l(x), "w_hat": t2l(m.weight), "u1": t2l(m.weight_u), "v1": t2l(m.weight_v),
                    "y": t2l(y), "grad_weight_orig": t2l(m.weight_orig.grad)})
        out[name] = rec
    torch.save(out, os.path.join(HERE, "spectral_norm.pt"))
    print("spectral norm fixtures written")


def gen_augment_hq():
    """`simclr_hq` / `simclr_hq_cutout` (augment/__init__.py:115-133) through the reference chain; kornia is absent, its
    two entry points come from contrad_b200/compat/kornia (parity unpinned for the blur)."""
    from augment import get_augment
    cases = []
    for (batch, size, seed, mode) in ((6, 32, 1, "simclr_hq"), (5, 32, 2, "simclr_hq_cutout"), (2, 64, 3, "simclr_hq_cutout")):
        seed_all(seed)
        x = torch.rand(batch, 3, size, size)
        dy = torch.randn(batch, 3, size, size)
        params, order = O.sample_simclr_params(batch, size, size)
        hq = O.sample_hq_params(batch, size, size, cutout=mode.endswith("cutout"))
        aug = get_augment(mode=mode)
        seed_all(seed)
        x_ref = torch.rand(batch, 3, size, size)
        _ = torch.randn(batch, 3, size, size)
        assert torch.equal(x, x_ref)
        x_ref.requires_grad_(True)
        y_ref = aug(x_ref)
        (y_ref * dy).sum().backward()
        cases.append({"mode": mode, "seed": seed, "x": t2l(x), "dy": t2l(dy), "order": order, "params": O.pack_params(params),
                      "hq": hq, "length": 15, "y": t2l(y_ref), "dx": t2l(x_ref.grad)})
        print("augment_hq case %s B=%d size=%d seed=%d" % (mode, batch, size, seed))
    torch.save({"cases": cases}, os.path.join(HERE, "augment_hq.pt"))


def gen_snresnet18():
    """D_SNResNet18 (models/gan/snresnet.py) through the reference module on CPU: outputs, input gradient and
    per-parameter gradient norms for seeded weights (oracle.make_d_resnet18_state) - only inputs / outputs are stored."""
    from models.gan import get_architecture
    _, D = get_architecture("snresnet18", (32, 32, 3))
    sd = O.make_d_resnet18_state(generator=torch.Generator().manual_seed(77))
    missing = D.load_state_dict(sd, strict=True)
    D.train()
    seed_all(78)
    x = torch.rand(6, 3, 32, 32)
    c_d, c1, c2 = torch.randn(6, 1), torch.randn(6, 128), torch.randn(6, 128)
    xr = x.clone().requires_grad_(True)
    d, aux = D(xr, projection=True, projection2=True, penultimate=True)
    ((d * c_d).sum() + (aux["projection"] * c1).sum() + (aux["projection2"] * c2).sum()).backward()
    out = {"w_seed": 77, "x": t2l(x), "c_d": c_d, "c1": c1, "c2": c2, "d": t2l(d), "projection": t2l(aux["projection"]),
           "projection2": t2l(aux["projection2"]), "penultimate": t2l(aux["penultimate"]), "dx": t2l(xr.grad),
           "grad_norms": {k: float(p.grad.norm()) for k, p in D.named_parameters() if p.grad is not None},
           "grad_conv1": t2l(D.conv1.weight_orig.grad), "keys": {k: list(v.shape) for k, v in D.state_dict().items()},
           "uv_after": {k: t2l(v) for k, v in D.state_dict().items() if k.endswith(("conv1.weight_u", "l1.weight_v"))}}
    torch.save(out, os.path.join(HERE, "snresnet18.pt"))
    print("snresnet18: d[0]=%.6f" % float(d[0]))


def gen_augment_aux(gin):
    """Rows f3 / f4: HorizontalFlipRandomCrop ('hfrt'), RandomCrop (all three padding modes), Gaussian noise through the
    reference classes; the explicit draws come from replaying the seed with the oracle's samplers (agreement of the
    OUTPUTS pins them).  The uint8 case stores bytes and the reference chain applied to ToTensor(bytes)."""
    from augment import get_augment
    from augment.spatial import HorizontalFlipRandomCrop, RandomCrop
    from augment import Gaussian
    cases = []
    specs = [("hfrt", 6, 32, 32, 4, 32, "reflection"), ("hfrt", 5, 32, 32, 4, 32, "zeros"),
             ("hfrt", 5, 32, 32, 6, 32, "border"), ("crop", 4, 32, 32, 4, 32, "reflection"),
             ("hfrt", 3, 64, 64, 4, 32, "reflection"), ("hfrt", 3, 20, 28, 5, 28, "reflection"),
             ("hfrt", 3, 20, 28, 5, 28, "zeros")]
    for k, (kind, batch, h, w, max_pixels, width, pad) in enumerate(specs):
        seed = 300 + k
        cls = HorizontalFlipRandomCrop if kind == "hfrt" else RandomCrop
        layer = cls(max_pixels=max_pixels, width=width, padding_mode=pad)
        seed_all(seed)
        x = torch.rand(batch, 3, h, w)
        dy = torch.randn(batch, 3, h, w)
        params = O.sample_shift_flip(batch, max_pixels, width, flip=(kind == "hfrt"))
        seed_all(seed)
        x_ref = torch.rand(batch, 3, h, w)
        _ = torch.randn(batch, 3, h, w)
        x_ref.requires_grad_(True)
        y_ref = layer(x_ref)
        (y_ref * dy).sum().backward()
        cases.append({"kind": kind, "seed": seed, "max_pixels": max_pixels, "width": width, "padding_mode": pad,
                      "x": t2l(x), "dy": t2l(dy), "params": t2l(params), "y": t2l(y_ref), "dx": t2l(x_ref.grad)})
        print("shift_flip case %s B=%d %dx%d pad=%s" % (kind, batch, h, w, pad))
    noise_cases = []
    for k, (batch, size, sigma) in enumerate(((4, 32, 0.12), (2, 30, 0.5))):
        seed = 320 + k
        seed_all(seed)
        x = torch.rand(batch, 3, size, size)
        dy = torch.randn(batch, 3, size, size)
        noise = torch.randn(batch, 3, size, size)
        seed_all(seed)
        x_ref = torch.rand(batch, 3, size, size)
        _ = torch.randn(batch, 3, size, size)
        x_ref.requires_grad_(True)
        y_ref = Gaussian(sigma=sigma)(x_ref)
        (y_ref * dy).sum().backward()
        noise_cases.append({"seed": seed, "sigma": sigma, "x": t2l(x), "dy": t2l(dy), "noise": t2l(noise),
                            "y": t2l(y_ref), "dx": t2l(x_ref.grad)})
    # uint8 input (row f3): reference = simclr()(cat([ToTensor(bytes)] * 2 + [fakes]))
    u8_cases = []
    for k, (n, m, size) in enumerate(((4, 4, 32), (2, 3, 64), (2, 2, 48))):
        seed = 340 + k
        gen = torch.Generator().manual_seed(seed)
        x_u8 = torch.randint(0, 256, (n, 3, size, size), generator=gen, dtype=torch.uint8)
        fakes = torch.rand(m, 3, size, size, generator=gen)
        dy = torch.randn(2 * n + m, 3, size, size, generator=gen)
        seed_all(seed)
        params, order = O.sample_simclr_params(2 * n + m, size, size)
        aug = get_augment(mode="simclr")
        seed_all(seed)
        fk = fakes.clone().requires_grad_(True)
        x_f = x_u8.to(torch.float32).div(255)
        y_ref = aug(torch.cat([x_f, x_f, fk], dim=0))
        (y_ref * dy).sum().backward()
        u8_cases.append({"seed": seed, "x_u8": x_u8, "fakes": t2l(fakes), "dy": dy, "params": O.pack_params(params),
                         "order": order, "y": t2l(y_ref), "d_fakes": t2l(fk.grad)})
        print("uint8 case n=%d m=%d size=%d order=%d" % (n, m, size, order))
    torch.save({"shift_flip": cases, "noise": noise_cases, "uint8": u8_cases}, os.path.join(HERE, "augment_aux.pt"))


def gen_diffaug(gin):
    """third_party/diffaug.DiffAugment through the reference for the registry's policy ('color,cutout') and the other
    canonical-order policies; draws replayed with oracle.sample_diffaug."""
    from third_party.diffaug import DiffAugment
    cases = []
    specs = [("color,cutout", 5, 32, 32), ("color,translation,cutout", 4, 32, 32), ("translation", 3, 16, 24),
             ("color", 3, 20, 20), ("cutout", 4, 33, 31), ("color,cutout", 2, 64, 64)]
    for k, (policy, batch, h, w) in enumerate(specs):
        seed = 500 + k
        seed_all(seed)
        x = torch.rand(batch, 3, h, w)
        dy = torch.randn(batch, 3, h, w)
        params = O.sample_diffaug(batch, h, w, stages=tuple(policy.split(",")))
        seed_all(seed)
        x_ref = torch.rand(batch, 3, h, w)
        _ = torch.randn(batch, 3, h, w)
        x_ref.requires_grad_(True)
        y_ref = DiffAugment(x_ref, policy=policy)
        (y_ref * dy).sum().backward()
        cases.append({"policy": policy, "seed": seed, "x": t2l(x), "dy": t2l(dy), "params": t2l(params), "y": t2l(y_ref),
                      "dx": t2l(x_ref.grad)})
        print("diffaug case %s B=%d %dx%d" % (policy, batch, h, w))
    torch.save({"cases": cases}, os.path.join(HERE, "diffaug.pt"))


def gen_baselines(gin):
    """training/gan/{std,aug,aug_both}.py with penalty none / cr / bcr and `--aug hfrt`, and simclr_only, through the
    reference modules at reduced width (ndf=4, d_hidden=16): losses, penalty and the D gradient norms."""
    from augment import get_augment
    from models.gan.sndcgan import D_SNDCGAN
    from importlib import import_module
    out = []
    combos = [("std", "none", "nonsat"), ("std", "cr", "hinge"), ("std", "bcr", "nonsat"), ("aug", "none", "lsgan"),
              ("aug_both", "bcr", "wgan"), ("aug", "cr", "nonsat")]
    for k, (mode, penalty, loss_kind) in enumerate(combos):
        seed = 400 + k
        seed_all(seed)
        D = D_SNDCGAN(image_size=(32, 32, 3), ndf=4, mlp_linear=True, d_hidden=16)
        D.train()
        sd_d0 = {kk: t2l(v) for kk, v in D.state_dict().items()}
        n = 4
        P = SimpleNamespace(augment_fn=get_augment(mode="hfrt"), temp=0.1, lbd_a=1.0, distributed=False, penalty=penalty)
        options = {"loss": loss_kind, "lbd": 10.0, "lbd2": 5.0}
        mod = import_module("training.gan.%s" % mode)
        seed_all(seed + 50)
        images = torch.rand(n, 3, 32, 32)
        gen = torch.rand(n, 3, 32, 32)
        d_loss, aux = mod.loss_D_fn(P, D, options, images, gen)
        (d_loss + aux["penalty"]).backward()
        rec = {"mode": mode, "penalty": penalty, "loss": loss_kind, "lbd": 10.0, "lbd2": 5.0, "sd_d": sd_d0,
               "images": t2l(images), "gen": t2l(gen), "d_loss": float(d_loss), "pen": float(aux["penalty"]),
               "d_real": float(aux["d_real"]), "d_gen": float(aux["d_gen"]), "grad_norms": _grad_norms(D)}
        # explicit draws: replay the stream (one hfrt call per P.augment_fn call, in call order)
        seed_all(seed + 50)
        _ = torch.rand(n, 3, 32, 32); _ = torch.rand(n, 3, 32, 32)
        calls = []
        if mode == "aug":
            calls.append(n)
        elif mode == "aug_both":
            calls.append(2 * n)
        if penalty == "cr":
            calls.append(n)
        elif penalty == "bcr":
            calls.append(2 * n)
        rec["aug_params"] = [O.sample_shift_flip(b, 4, 32, flip=True) for b in calls]
        g_loss = mod.loss_G_fn(P, D, options, images, gen)
        rec["g_loss"] = float(g_loss)
        if mode == "aug_both":
            rec["aug_params_g"] = None      # drawn after the D step; the G loss is checked by its own seed below
        out.append(rec)
        print("baseline %s/%s/%s: d_loss=%.6f pen=%.6f" % (mode, penalty, loss_kind, rec["d_loss"], rec["pen"]))
    torch.save({"cases": out}, os.path.join(HERE, "baseline_modes.pt"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gin = ref_import.activate()
    todo = args.only.split(",") if args.only else ["augment", "augment_hq", "contrastive", "sn", "small", "config1", "snresnet18",
                                                     "augment_aux", "baselines", "diffaug"]
    if "augment" in todo:
        gen_augment(gin)
    if "augment_hq" in todo:
        gen_augment_hq()
    if "snresnet18" in todo:
        gen_snresnet18()
    if "contrastive" in todo:
        gen_contrastive()
    if "sn" in todo:
        gen_spectral_norm()
    if "small" in todo:
        gen_small_models(gin)
    if "config1" in todo:
        gen_config1(gin)
    if "augment_aux" in todo:
        gen_augment_aux(gin)
    if "baselines" in todo:
        gen_baselines(gin)
    if "diffaug" in todo:
        gen_diffaug(gin)


if __name__ == "__main__":
    main()
