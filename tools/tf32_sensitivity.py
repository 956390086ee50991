#!/usr/bin/env python
"""CPU study: which TF32 roundings move the gradient norms of the ContraD step away from the fp32 reference.

The product rounds every tensor-core operand to TF32 (10-bit mantissa, round-to-nearest): activations and data
gradients by the epilogue that produces them, weights by the packing kernels.  This script emulates exactly that on the
CPU oracle (oracle/contrad_oracle.py, fp32) by wrapping F.conv2d / F.linear / F.conv_transpose2d with straight-through
roundings that can be switched per operand class and per network, and reports the relative deviation of
L_con / L_dis / L_gen / |grad D| / |grad G| from the unrounded fp32 run on identical inputs.  It decides where the
error-compensated (split hi + lo, "3xTF32") operands are worth their cost.  TEST / DESIGN infrastructure only.

    python tools/tf32_sensitivity.py --n 64 --seeds 3
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import contrad_oracle as O  # noqa: E402


def round_tf32(t):
    """cvt.rna.tf32.f32: round to nearest (ties away) on the 13 dropped mantissa bits."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class RoundST(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return round_tf32(x)

    @staticmethod
    def backward(ctx, g):
        return g


class GradRound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return round_tf32(g)


class Emu(object):
    """cfg[net] = set of operand classes to round: 'act', 'w', 'grad'.  net = 'd' | 'g' decided by a flag the driver
    flips around the generator's forward.  The first D layer (3 -> 64, SIMT fp32 in the product) is never rounded on
    its input / weight; its incoming gradient is (the product rounds g before the first-layer wgrad)."""

    def __init__(self, cfg):
        self.cfg, self.net = cfg, "d"
        self.orig = (F.conv2d, F.linear, F.conv_transpose2d)

    def _wrap(self, op, first_layer_check):
        def fn(x, w, b=None, *a, **k):
            flags = self.cfg.get(self.net, ())
            first = first_layer_check(w)
            only = self.cfg.get("only_shapes")
            if only is not None and tuple(w.shape) not in only:
                flags = ()
            skip = self.cfg.get("skip_shapes")
            if skip is not None and tuple(w.shape) in skip:
                flags = ()
            if "act" in flags and not first:
                x = RoundST.apply(x)
            if "w" in flags and not first:
                w = RoundST.apply(w)
            y = op(x, w, None, *a, **k)
            if "grad" in flags:
                y = GradRound.apply(y)
            if b is not None:
                y = y + (b.view(1, -1, 1, 1) if y.dim() == 4 else b)
            return y
        return fn

    def __enter__(self):
        conv, lin, convt = self.orig
        O.F.conv2d = self._wrap(conv, lambda w: w.dim() == 4 and w.shape[1] == 3)
        O.F.linear = self._wrap(lin, lambda w: False)
        O.F.conv_transpose2d = self._wrap(convt, lambda w: False)
        return self

    def __exit__(self, *exc):
        O.F.conv2d, O.F.linear, O.F.conv_transpose2d = self.orig
        return False


def run(n, seed, cfg):
    gen_w = torch.Generator().manual_seed(1000 + seed)
    sd_d, sd_g = O.make_d_state(generator=gen_w), O.make_g_state(generator=gen_w)
    np.random.seed(seed); torch.manual_seed(seed)
    images = torch.rand(n, 3, 32, 32)
    z_d = O.sample_latent(n); aug_d = O.sample_simclr_params(3 * n, 32, 32)
    z_g = O.sample_latent(n); aug_g = O.sample_simclr_params(n, 32, 32)
    emu = Emu(cfg)
    g_fwd = O.g_sndcgan_forward

    def g_forward(sd, z, *a, **k):
        emu.net = "g"
        try:
            return g_fwd(sd, z, *a, **k)
        finally:
            emu.net = "d"

    O.g_sndcgan_forward = g_forward
    try:
        with emu:
            og = O.Adam(O.trainable(sd_g).values(), 2e-4); od = O.Adam(O.trainable(sd_d).values(), 2e-4)
            return O.train_step(sd_g, sd_d, og, od, images, z_d, z_g, aug_d, aug_g, step=1)
    finally:
        O.g_sndcgan_forward = g_fwd


CONFIGS = {
    "product (everything rounded)": {"d": ("act", "w", "grad"), "g": ("act", "w", "grad")},
    "weights only": {"d": ("w",), "g": ("w",)},
    "activations only": {"d": ("act",), "g": ("act",)},
    "gradients only": {"d": ("grad",), "g": ("grad",)},
    "act + grad (weights exact)": {"d": ("act", "grad"), "g": ("act", "grad")},
    "D rounded, G exact": {"d": ("act", "w", "grad"), "g": ()},
    "G rounded, D exact": {"d": (), "g": ("act", "w", "grad")},
    "D: w only; G: all": {"d": ("w",), "g": ("act", "w", "grad")},
    "D: act+w (grad exact); G exact": {"d": ("act", "w"), "g": ()},
    "D: grad only; G exact": {"d": ("grad",), "g": ()},
}


D_SHAPES = [(128, 64, 4, 4), (128, 128, 3, 3), (256, 128, 4, 4), (256, 256, 3, 3), (512, 256, 4, 4), (512, 512, 3, 3),
            (512, 8192), (1, 512), (128, 512)]
G_SHAPES = [(8192, 128), (512, 256, 4, 4), (256, 128, 4, 4), (128, 64, 4, 4), (64, 3, 3, 3)]


def per_layer(n, seeds):
    """Round ONE layer (activation + weight + gradient operands) and leave everything else exact."""
    all_flags = ("act", "w", "grad")
    rows = {}
    for seed in range(seeds):
        ref = run(n, seed, {})
        for net, shapes in (("d", D_SHAPES), ("g", G_SHAPES)):
            for shp in shapes:
                got = run(n, seed, {net: all_flags, "only_shapes": {shp}})
                rel = abs(got["g_grad_norm"] - ref["g_grad_norm"]) / ref["g_grad_norm"]
                reld = abs(got["d_grad_norm"] - ref["d_grad_norm"]) / ref["d_grad_norm"]
                rows.setdefault((net, shp), []).append((rel, reld))
                print("seed %d  %s %-18s g_grad_norm %.1e  d_grad_norm %.1e" % (seed, net, shp, rel, reld), flush=True)
    for k, v in rows.items():
        print(k, "max g %.1e  max d %.1e" % (max(a for a, _ in v), max(b for _, b in v)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-layer", action="store_true")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if args.per_layer:
        return per_layer(args.n, args.seeds)
    keys = ("l_con_pos", "l_con_neg", "l_dis", "l_gen", "d_grad_norm", "g_grad_norm")
    table = {}
    for seed in range(args.seeds):
        ref = run(args.n, seed, {})
        for name, cfg in CONFIGS.items():
            if args.only and args.only not in name:
                continue
            got = run(args.n, seed, cfg)
            rel = {k: abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-30) for k in keys}
            table.setdefault(name, []).append(rel)
            print("seed %d  %-34s " % (seed, name) + "  ".join("%s %.1e" % (k, rel[k]) for k in keys), flush=True)
    summary = {name: {k: float(max(r[k] for r in rows)) for k in keys} for name, rows in table.items()}
    print(json.dumps(summary, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"n": args.n, "seeds": args.seeds, "max_rel_dev": summary}, f, indent=1)


if __name__ == "__main__":
    main()
