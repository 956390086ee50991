"""First GPU visit: microbenchmarks of the augment kernel (HBM roofline point) and the tap GEMM."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from contrad_b200 import kernels as K
from oracle import contrad_oracle as O

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2], ts[0]

out = {}
print(torch.cuda.get_device_name(0))
for B, size in ((65536, 32), (1536, 32), (512, 32), (8192, 64)):
    np.random.seed(0); torch.manual_seed(0)
    params, order = O.sample_simclr_params(B, size, size)
    p = O.pack_params(params).cuda()
    x = torch.rand(B, 3, size, size, device="cuda")
    dy = torch.randn(B, 3, size, size, device="cuda")
    for od in (0, 1):
        med, best = timeit(lambda: K.augment_simclr_fwd(x, p, od))
        gbs = 8 * x.numel() / med / 1e6
        print("augment fwd B=%d %dx%d order=%d: %.3f ms median (%.3f best)  %.0f GB/s algorithmic" % (B, size, size, od, med, best, gbs))
        out["aug_fwd_B%d_s%d_o%d" % (B, size, od)] = {"ms": med, "gbs": gbs}
        med, best = timeit(lambda: K.augment_simclr_bwd(x, dy, p, od))
        gbs = 12 * x.numel() / med / 1e6
        print("augment bwd B=%d %dx%d order=%d: %.3f ms median  %.0f GB/s (12 B/elem)" % (B, size, size, od, med, gbs))
        out["aug_bwd_B%d_s%d_o%d" % (B, size, od)] = {"ms": med, "gbs": gbs}
    del x, dy

# copy bandwidth reference in the same process
a = torch.empty(1 << 28, device="cuda"); b = torch.empty_like(a)
med, best = timeit(lambda: b.copy_(a))
print("torch copy 1 GiB: %.3f ms -> %.0f GB/s" % (med, 2 * a.numel() * 4 / med / 1e6))
del a, b

layers = [  # B, H, Cin, Cout, ks, stride   (config 2, D-step batch 1536)
    (1536, 32, 64, 128, 4, 2), (1536, 16, 128, 128, 3, 1), (1536, 16, 128, 256, 4, 2), (1536, 8, 256, 256, 3, 1),
    (1536, 8, 256, 512, 4, 2), (1536, 4, 512, 512, 3, 1)]
for (B, H, Cin, Cout, ks, st) in layers:
    x = K.round_tf32(torch.randn(B, H, H, Cin, device="cuda"))
    w = K.round_tf32(torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.02)
    bias = torch.zeros(Cout, device="cuda")
    wm = K.pack_fwd_weight(w); wt = K.pack_dgrad_weight(w, st)
    Ho = H // st
    flops = 2.0 * B * Ho * Ho * Cout * Cin * ks * ks
    med, best = timeit(lambda: K.conv2d_nhwc_fwd(x, wm, bias, ks, st, slope=0.1, round_out=True))
    print("conv fwd  B=%d %dx%d %d->%d k%d s%d: %.3f ms  %.1f TFLOP/s" % (B, H, H, Cin, Cout, ks, st, med, flops / med / 1e9))
    out["conv_fwd_%d_%d_%d" % (H, Cin, Cout)] = {"ms": med, "tflops": flops / med / 1e9}
    dy = K.round_tf32(torch.randn(B, Ho, Ho, Cout, device="cuda"))
    med, best = timeit(lambda: K.conv2d_nhwc_dgrad(dy, wt, (B, H, H, Cin), ks, st, act_in=x, slope=0.1, round_out=True))
    print("conv dgrad B=%d %dx%d %d->%d k%d s%d: %.3f ms  %.1f TFLOP/s" % (B, H, H, Cin, Cout, ks, st, med, flops / med / 1e9))
    out["conv_dgrad_%d_%d_%d" % (H, Cin, Cout)] = {"ms": med, "tflops": flops / med / 1e9}
    med, best = timeit(lambda: K.conv2d_nhwc_wgrad(x, dy, ks, st))
    print("conv wgrad B=%d %dx%d %d->%d k%d s%d: %.3f ms  %.1f TFLOP/s" % (B, H, H, Cin, Cout, ks, st, med, flops / med / 1e9))
    out["conv_wgrad_%d_%d_%d" % (H, Cin, Cout)] = {"ms": med, "tflops": flops / med / 1e9}
    # cuDNN TF32 reference timing
    torch.backends.cudnn.allow_tf32 = True
    xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    wc = w.contiguous(memory_format=torch.channels_last)
    med, best = timeit(lambda: torch.nn.functional.conv2d(xc, wc, bias, stride=st, padding=1))
    print("   cudnn tf32 fwd (channels_last): %.3f ms  %.1f TFLOP/s" % (med, flops / med / 1e9))
    del x, w, dy, xc, wc
a = K.round_tf32(torch.randn(1536, 8192, device="cuda")); b = K.round_tf32(torch.randn(1536, 8192, device="cuda") * 0.02)
med, _ = timeit(lambda: K.gemm_nt(a, b, None, slope=0.1))
print("heads gemm 1536x1536x8192: %.3f ms %.1f TFLOP/s" % (med, 2 * 1536 * 1536 * 8192 / med / 1e9))
dh = K.round_tf32(torch.randn(1536, 1536, device="cuda"))
med, _ = timeit(lambda: K.gemm_tn_wgrad(dh, a))
print("heads wgrad 1536x1536x8192: %.3f ms %.1f TFLOP/s" % (med, 2 * 1536 * 1536 * 8192 / med / 1e9))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_r1a.json", "w"), indent=1)
