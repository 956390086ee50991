"""Roofline points of the HBM-bound kernels added for the StyleGAN2 / high-resolution rows: achieved
GB/s = ALGORITHMIC bytes (each operand read once, each result written once) / CUDA-event time, against the measured copy
bandwidth of MEASURED_PEAKS.json.  Operands are larger than L2 (126 MB) or rotated over several buffers so that no
iteration finds its input resident.  Prints one JSON object; run on the GPU box:  python tools/bench_sg2_kernels.py"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.append(os.path.join(REPO, "contrad_b200", "compat"))


def timed(fn, sets, iters=20):
    """fn(i-th operand set); operand sets rotate so every call touches cold data."""
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    import numpy as np
    from contrad_b200 import kernels as K
    from contrad_b200 import sg2_kernels as S
    from contrad_b200.augment.layers import gaussian_taps

    peaks = {"hbm_gbs": 6650.0}
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak = peaks["hbm_gbs"]
    dev = "cuda"
    fir = torch.tensor([1., 3., 3., 1.])
    fir = (fir[None] * fir[:, None] / 64).to(dev)
    out = {"peak_gbs": peak, "peak_note": "measured copy bandwidth (MEASURED_PEAKS.json)" if os.path.exists(p) else "fallback", "kernels": {}}

    def record(name, ms, nbytes, note):
        gbs = nbytes / ms / 1e6
        out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "achieved_gbs": round(gbs, 1),
                                "frac": round(gbs / peak, 3), "shape": note}

    nset = 3
    # ---- config 4, D-step batch 192 (3 x 64), 32x32x128
    B, H, C = 192, 32, 128
    xs = [torch.randn(B, H, H, C, device=dev) for _ in range(nset)]
    ms = timed(lambda x: S.upfirdn2d(x, fir, 1, 1, (2, 2, 2, 2), round_out=True), xs)
    record("upfirdn2d blur pad(2,2)", ms, 4 * (xs[0].numel() + B * (H + 1) ** 2 * C), "[192,32,32,128] -> 33x33")
    ms = timed(lambda x: S.upfirdn2d(x, fir, 1, 2, (1, 1, 1, 1), round_out=True), xs)
    record("upfirdn2d blur + down 2", ms, 4 * (xs[0].numel() + B * (H // 2) ** 2 * C), "[192,32,32,128] -> 16x16")
    ts = [torch.randn(B, H + 1, H + 1, C, device=dev) for _ in range(nset)]
    ms = timed(lambda t: S.patch_s2_gather(t), ts)
    record("patch_s2_gather", ms, 4 * (ts[0].numel() + B * (H // 2) ** 2 * 9 * C), "[192,33,33,128] -> [192,16,16,9,128]")
    us = [torch.randn(B, H // 2, H // 2, 9, C, device=dev) for _ in range(nset)]
    ms = timed(lambda u: S.patch_s2_scatter(u), us)
    record("patch_s2_scatter", ms, 4 * (us[0].numel() + B * (H + 1) ** 2 * C), "[192,16,16,9,128] -> [192,33,33,128]")
    bias = torch.randn(C, device=dev)
    ms = timed(lambda x: S.bias_act(x, bias, 0.2, 2 ** 0.5, round_out=True), xs)
    record("bias_act fwd", ms, 8 * xs[0].numel(), "[192,32,32,128]")
    ms = timed(lambda x: S.bias_act_grad(x, xs[0], bias, 0.2, 2 ** 0.5, round_out=True), xs)
    record("bias_act grad", ms, 12 * xs[0].numel(), "[192,32,32,128]")
    s = torch.rand(B, C, device=dev)
    noise = torch.randn(B, 1, H, H, device=dev)
    nw = torch.randn(1, device=dev)
    ms = timed(lambda x: S.modulate(x, s, round_out=True), xs)
    record("modulate", ms, 8 * xs[0].numel(), "[192,32,32,128]")
    ms = timed(lambda x: S.mod_epilogue(x, s, noise, nw, bias, round_out=True), xs)
    record("mod_epilogue", ms, 8 * xs[0].numel(), "[192,32,32,128]")
    ms = timed(lambda x: S.mul_reduce(x, xs[0]), xs)
    record("mul_reduce", ms, 8 * xs[0].numel(), "[192,32,32,128]")
    ms = timed(lambda x: K.round_tf32_(x), xs)
    record("round_tf32", ms, 8 * xs[0].numel(), "[192,32,32,128]")
    del xs, ts, us
    # ---- config 5: 8 images per replica at 512x512
    B, H, C = 8, 512, 32
    xs = [torch.randn(B, H, H, C, device=dev) for _ in range(nset)]
    ms = timed(lambda x: S.upfirdn2d(x, fir, 1, 1, (2, 2, 2, 2), round_out=True), xs)
    record("upfirdn2d blur pad(2,2) 512^2", ms, 4 * (xs[0].numel() + B * (H + 1) ** 2 * C), "[8,512,512,32] -> 513x513")
    ms = timed(lambda x: S.bias_act(x, bias[:C].contiguous(), 0.2, 2 ** 0.5, round_out=True), xs)
    record("bias_act fwd 512^2", ms, 8 * xs[0].numel(), "[8,512,512,32]")
    del xs
    # ---- augmentation of 512x512 images: chain (large path), blur k=51, cutout
    B = 48
    imgs = [torch.rand(B, 3, 512, 512, device=dev) for _ in range(nset)]
    prm = torch.zeros(11, B, device=dev)
    prm[0] = 0.7; prm[1] = 0.8; prm[2] = 0.1; prm[3] = -0.1; prm[4] = 1.0; prm[5] = 1.0; prm[6] = 1.2; prm[7] = 0.05
    prm[8] = 1.1; prm[9] = 0.9; prm[10] = 0.0
    prm[4, ::2] = -1.0; prm[10, ::5] = 1.0
    ms = timed(lambda x: K.augment_simclr_large_fwd(x, prm, 0), imgs, iters=10)
    record("augment_simclr_large_fwd", ms, 8 * imgs[0].numel(), "[48,3,512,512], colour jitter on every image")
    y, means = K.augment_simclr_large_fwd(imgs[0], prm, 0)
    ms = timed(lambda x: K.augment_simclr_large_bwd(imgs[0], x, prm, 0, means), imgs, iters=10)
    record("augment_simclr_large_bwd", ms, 12 * imgs[0].numel(), "[48,3,512,512]")
    taps = gaussian_taps(51, 1.0).to(dev)
    on = torch.ones(B, device=dev)
    ms = timed(lambda x: K.gaussian_blur(x, taps, on), imgs, iters=10)
    record("gaussian_blur k=51", ms, 16 * imgs[0].numel(), "[48,3,512,512], two passes")
    cut = torch.stack([on, torch.full((B,), 200.0, device=dev), torch.full((B,), 300.0, device=dev)])
    ms = timed(lambda x: K.cutout(x, cut, 255), imgs, iters=10)
    record("cutout 255", ms, 8 * imgs[0].numel(), "[48,3,512,512]")
    print(json.dumps(out))
