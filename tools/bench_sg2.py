"""Side measurement for SURVEY config 4 (StyleGAN2 small32 + ContraD, c10_style64.gin: batch 64, --no_lazy R1 every
step, --aug=simclr) on one B200: images/s of the full step (engine.train_step_stylegan2, eager launches).  The torch-eager yardstick for the same arithmetic is timed
by tests/test_gpu_sg2.py::test_eager_gpu_yardstick_timing (only tests may run the oracle).
Not the headline bench (bench.py keeps BASELINE.json's config 2); prints one JSON line."""
import argparse
import copy
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.append(os.path.join(REPO, "contrad_b200", "compat"))


def measure(batch=64, steps=10, warmup=3, d_reg_every=1, graph=False):
    """images/s of the full config-4 step (train_stylegan2_contraD.py:195-236 through engine.train_step_stylegan2 or
    its CUDA-graph replay) with device-resident synthetic images; returns a dict."""
    import gin
    from contrad_b200 import _capi, engine
    from contrad_b200.augment import get_augment
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.optim import FusedAdam
    from contrad_b200.training.gan import stylegan2 as T
    gin.clear_config()
    gin.parse_config("RandomResizeCropLayer.scale = (0.2, 1.0)\nColorJitterLayer.brightness = 0.4\n"
                     "ColorJitterLayer.contrast = 0.4\nColorJitterLayer.saturation = 0.4\nColorJitterLayer.hue = 0.1\n")
    torch.manual_seed(0); np.random.seed(0)
    n = batch
    G, D = get_architecture("stylegan2", (32, 32, 3))
    G.cuda(); D.cuda()
    g_ema = copy.deepcopy(G)
    GD = T.G_D(G, D, get_augment(mode="simclr").cuda())
    P = SimpleNamespace(use_warmup=True, halflife_lr=0, ema_start_k=0, accum=0.5 ** (n / 1000000.0), d_reg_every=d_reg_every,
                        lbd_r1=0.1, style_mix=0.9, temp=0.1, lbd_a=1.0, distributed=False)
    opt = {"warmup": 3000, "lr": 2e-3, "lr_d": 2e-3, "batch_size": n}
    opts = (FusedAdam(G.parameters(), lr=2e-3, betas=(0.0, 0.99)), FusedAdam(D.parameters(), lr=2e-3, betas=(0.0, 0.99)))
    images = [torch.rand(n, 3, 32, 32, device="cuda") for _ in range(4)]
    graphed = engine.GraphedStyleGAN2Step(P, opt, GD, g_ema, opts) if graph else None

    def run(k, first):
        out = None
        for s in range(first, first + k):
            if graphed is not None:
                out = graphed(images[s % 4], s)
            else:
                out = engine.train_step_stylegan2(P, opt, GD, g_ema, opts, images[s % 4], s)
        return out

    extra = 5 if graph else 0                       # eager warm-up + capture of the graphed variant
    run(warmup + extra, 1)
    torch.cuda.synchronize()
    l0 = _capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run(steps, 1 + warmup + extra)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"workload": "StyleGAN2(small32)+ContraD 32x32 b%d (c10_style64.gin), R1 every %d step(s), --aug=simclr, synthetic images"
                       % (n, d_reg_every),
           "launch": "cuda-graph replay" if graph else "eager", "ms_per_step": ms, "images_per_s": n / ms * 1e3,
           "library_launches_per_step": (_capi.launch_count() - l0) / steps, "steps": steps,
           "losses": {k: float(v) for k, v in out.items()}}
    if graphed is not None:
        graphed.release()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--d-reg-every", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="replay the step as a CUDA graph (engine.GraphedStyleGAN2Step)")
    args = ap.parse_args()
    print(json.dumps(measure(args.batch, args.steps, args.warmup, args.d_reg_every, args.graph)))


if __name__ == "__main__":
    main()
