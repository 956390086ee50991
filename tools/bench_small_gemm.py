#!/usr/bin/env python
"""Micro-benchmark of the tap-GEMM entry points on the per-rank shapes of the 8-GPU run (64 images per rank: D step
B = 192, G step B = 64): CUDA-event time per launch, achieved TFLOP/s.  Run once per CB200_TAPGEMM_SPLITK setting."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from contrad_b200 import kernels as K  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3        # us


def main():
    out = []
    dev = "cuda"
    for B in (192, 64):
        for (H, Cin, Cout, ks, st) in ((32, 64, 128, 4, 2), (16, 128, 128, 3, 1), (16, 128, 256, 4, 2), (8, 256, 256, 3, 1),
                                       (8, 256, 512, 4, 2), (4, 512, 512, 3, 1)):
            x = K.round_tf32(torch.randn(B, H, H, Cin, device=dev))
            w = K.round_tf32(torch.randn(Cout, Cin, ks, ks, device=dev) * 0.02)
            wf, wt = K.pack_fwd_weight(w), K.pack_dgrad_weight(w, st)
            Ho = H // st
            dy = K.round_tf32(torch.randn(B, Ho, Ho, Cout, device=dev))
            flops = 2.0 * B * Ho * Ho * Cout * Cin * ks * ks
            t_f = timeit(lambda: K.conv2d_nhwc_fwd(x, wf, None, ks, st, slope=0.1, round_out=True))
            t_d = timeit(lambda: K.conv2d_nhwc_dgrad(dy, wt, (B, H, H, Cin), ks, st, act_in=x, slope=0.1, round_out=True))
            t_w = timeit(lambda: K.conv2d_nhwc_wgrad(x, dy, ks, st)) if Cout % 128 == 0 else float("nan")
            out.append({"B": B, "layer": "%dx%d %d->%d k%ds%d" % (H, H, Cin, Cout, ks, st), "gflop": flops / 1e9,
                        "fwd_us": t_f, "fwd_tflops": flops / t_f / 1e6, "dgrad_us": t_d, "dgrad_tflops": flops / t_d / 1e6,
                        "wgrad_us": t_w, "wgrad_tflops": flops / t_w / 1e6})
        a = K.round_tf32(torch.randn(B, 8192, device=dev))
        wc = K.round_tf32(torch.randn(1536, 8192, device=dev) * 0.02)
        flops = 2.0 * B * 1536 * 8192
        t = timeit(lambda: K.gemm_nt(a, wc, None, slope=0.1, round_out=True))
        dh = K.round_tf32(torch.randn(B, 1536, device=dev))
        wct = K.round_tf32(torch.randn(8192, 1536, device=dev) * 0.02)
        t2 = timeit(lambda: K.gemm_nt(dh, wct, None))
        t3 = timeit(lambda: K.gemm_tn_wgrad(dh, a))
        out.append({"B": B, "layer": "heads 8192->1536", "gflop": flops / 1e9, "fwd_us": t, "fwd_tflops": flops / t / 1e6,
                    "dgrad_us": t2, "dgrad_tflops": flops / t2 / 1e6, "wgrad_us": t3, "wgrad_tflops": flops / t3 / 1e6})
    for r in out:
        print("B=%3d %-26s %7.2f GF  fwd %7.1f us %6.1f TF/s | dgrad %7.1f us %6.1f TF/s | wgrad %7.1f us %6.1f TF/s"
              % (r["B"], r["layer"], r["gflop"], r["fwd_us"], r["fwd_tflops"], r["dgrad_us"], r["dgrad_tflops"], r["wgrad_us"],
                 r["wgrad_tflops"]))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
