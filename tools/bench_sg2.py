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


class _DataParallelGD(object):
    """What train_stylegan2_contraD.py:376 builds: `nn.DataParallel(G_D(G, D, augment_fn))` - calls go through the
    DataParallel wrapper (scatter / replicate / parallel_apply / gather), G and D stay reachable for the optimisers."""

    def __init__(self, gd, device_ids):
        self.G, self.D = gd.G, gd.D
        self.dp = torch.nn.DataParallel(gd, device_ids=device_ids)

    def __call__(self, *args, **kwargs):
        return self.dp(*args, **kwargs)


def measure_data_parallel(gpus=8, batch=64, size=512, arch="stylegan2_512", steps=16, warmup=2, d_reg_every=16, lbd_r1=0.5):
    """BASELINE config 5: StyleGAN2_512 + ContraD (afhq_dog_style64.gin: b64, lr 2.5e-3, RRC scale (0.08, 1), CJ 0.8 / 0.8 /
    0.8 / 0.2, --lbd_r1 0.5, lazy R1 every 16 steps, --aug=simclr) on synthetic 512x512 images, ONE process driving
    `gpus` GPUs through nn.DataParallel(G_D) exactly like train_stylegan2_contraD.py:195-236,376 (replicas only).
    `steps` = a multiple of d_reg_every so that the timed window holds its share of R1 steps."""
    import gin
    from contrad_b200 import _capi, engine
    from contrad_b200.augment import get_augment
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.optim import FusedAdam
    from contrad_b200.training.gan import stylegan2 as T
    gin.clear_config()
    gin.parse_config("RandomResizeCropLayer.scale = (0.08, 1.0)\nColorJitterLayer.brightness = 0.8\n"
                     "ColorJitterLayer.contrast = 0.8\nColorJitterLayer.saturation = 0.8\nColorJitterLayer.hue = 0.2\n")
    torch.manual_seed(0); np.random.seed(0)
    n = batch
    torch.cuda.set_device(0)
    G, D = get_architecture(arch, (size, size, 3))
    G.cuda(0); D.cuda(0)
    g_ema = copy.deepcopy(G)
    GD = _DataParallelGD(T.G_D(G, D, get_augment(mode="simclr").cuda(0)), list(range(gpus)))
    P = SimpleNamespace(use_warmup=True, halflife_lr=0, ema_start_k=0, accum=0.5 ** (n / 20000.0), d_reg_every=d_reg_every,
                        lbd_r1=lbd_r1, style_mix=0.9, temp=0.1, lbd_a=1.0, distributed=False)
    opt = {"warmup": 3000, "lr": 2.5e-3, "lr_d": 2.5e-3, "batch_size": n}
    opts = (FusedAdam(G.parameters(), lr=2.5e-3, betas=(0.0, 0.99)), FusedAdam(D.parameters(), lr=2.5e-3, betas=(0.0, 0.99)))
    images = [torch.rand(n, 3, size, size, device="cuda:0") for _ in range(2)]

    def run(k, first):
        out = None
        for s in range(first, first + k):
            out = engine.train_step_stylegan2(P, opt, GD, g_ema, opts, images[s % 2], s)
        return out

    run(warmup, 1)
    for d in range(gpus):
        torch.cuda.synchronize(d)
    l0 = _capi.launch_count()
    t0 = time.perf_counter()
    out = run(steps, 1 + warmup)
    for d in range(gpus):
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    ms = 1e3 * dt / steps
    # algorithmic FLOPs per image and step (D forward 38.3 GFLOP/img, G forward ~71 GFLOP/img at 512^2, cm = 1):
    # G step = 3 G-forwards + 2 D-forwards, D step = 1 G-forward + 9 D-forwards (3n images, forward + 2 backward passes)
    flop_img = (4 * 71.0 + 11 * 38.3) * 1e9 if size == 512 else None
    res = {"workload": "StyleGAN2_512+ContraD %dx%d b%d (afhq_dog_style64.gin), lazy R1 every %d steps, --aug=simclr, synthetic "
                       "images, nn.DataParallel over %d GPUs (one process)" % (size, size, n, d_reg_every, gpus),
           "launch": "eager (DataParallel threads)", "n_gpus": gpus, "per_replica_batch": n // gpus, "ms_per_step": ms,
           "images_per_s": n / dt * steps, "steps": steps, "warmup": warmup,
           "library_launches_per_step": (_capi.launch_count() - l0) / steps,
           "losses": {k: float(v) for k, v in out.items()}}
    if flop_img:
        res["step_tflops"] = flop_img * res["images_per_s"] / 1e12
        res["tflops_note"] = "algorithmic FLOPs 705 GFLOP per image and step (R1 steps excluded), all GPUs together"
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--d-reg-every", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="replay the step as a CUDA graph (engine.GraphedStyleGAN2Step)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"],
                    help="reference = the unmodified reference's own train loop (oracle/ref_runner.run_gpu_stylegan2)")
    ap.add_argument("--data-parallel", type=int, default=0, help="config 5: StyleGAN2_512 over this many GPUs via nn.DataParallel")
    ap.add_argument("--size", type=int, default=512)
    args = ap.parse_args()
    if args.impl == "reference":
        from oracle import ref_runner
        if args.data_parallel:      # config 5 (afhq_dog_style64.gin: lazy R1 every 16, lbd_r1 0.5, halflife_k 20)
            r = ref_runner.run_gpu_stylegan2(args.steps, args.warmup, "stylegan2_512", 512, args.batch,
                                             "configs/gan/stylegan2/afhq_dog_style64.gin", lbd_r1=0.5, no_lazy=False, halflife_k=20,
                                             device_ids=list(range(args.data_parallel)))
        else:                       # config 4 (c10_style64.gin, --no_lazy, lbd_r1 0.1, halflife_k 1000)
            r = ref_runner.run_gpu_stylegan2(args.steps, args.warmup, "stylegan2", 32, args.batch, device_ids=[0])
        r["impl"] = "reference"
        print(json.dumps(r))
        return
    if args.data_parallel:
        arch = "stylegan2_512" if args.size == 512 else "stylegan2"
        print(json.dumps(measure_data_parallel(args.data_parallel, args.batch, args.size, arch, steps=args.steps,
                                               warmup=args.warmup, d_reg_every=args.d_reg_every if args.d_reg_every > 1 else 16)))
        return
    print(json.dumps(measure(args.batch, args.steps, args.warmup, args.d_reg_every, args.graph)))


if __name__ == "__main__":
    main()
