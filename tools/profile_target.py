"""Small driver for `ncu --set full`: runs one hot kernel of the D-step a few times.
usage: python tools/profile_target.py {conv3x3|conv4x4|dgrad|wgrad|augment|heads}"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from contrad_b200 import kernels as K
which = sys.argv[1] if len(sys.argv) > 1 else "conv3x3"
B = 1536
torch.manual_seed(0)
if which in ("conv3x3", "dgrad", "wgrad"):
    H, Cin, Cout, ks, st = 16, 128, 128, 3, 1
elif which == "conv4x4":
    H, Cin, Cout, ks, st = 16, 128, 256, 4, 2
if which in ("conv3x3", "conv4x4", "dgrad", "wgrad"):
    x = K.round_tf32(torch.randn(B, H, H, Cin, device="cuda"))
    w = K.round_tf32(torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.02)
    bias = torch.zeros(Cout, device="cuda")
    wm, wt = K.pack_fwd_weight(w), K.pack_dgrad_weight(w, st)
    dy = K.round_tf32(torch.randn(B, H // st, H // st, Cout, device="cuda"))
    for _ in range(4):
        if which in ("conv3x3", "conv4x4"):
            K.conv2d_nhwc_fwd(x, wm, bias, ks, st, slope=0.1, round_out=True)
        elif which == "dgrad":
            K.conv2d_nhwc_dgrad(dy, wt, (B, H, H, Cin), ks, st, act_in=x, slope=0.1, round_out=True)
        else:
            K.conv2d_nhwc_wgrad(x, dy, ks, st)
elif which == "heads":
    a = K.round_tf32(torch.randn(B, 8192, device="cuda")); b = K.round_tf32(torch.randn(1536, 8192, device="cuda") * 0.02)
    for _ in range(4):
        K.gemm_nt(a, b, None, slope=0.1, round_out=True)
elif which == "augment":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _params import random_simclr_params
    Bn = 65536
    p, order = random_simclr_params(Bn)
    xx = torch.rand(Bn, 3, 32, 32, device="cuda")
    for _ in range(4):
        K.augment_simclr_fwd(xx, p, order)
elif which in ("conv_first_wgrad", "conv_first_fwd"):
    xx = torch.rand(B, 3, 32, 32, device="cuda")
    w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
    bias = torch.zeros(64, device="cuda")
    dy = K.round_tf32(torch.randn(B, 32, 32, 64, device="cuda"))
    for _ in range(4):
        if which == "conv_first_fwd":
            K.conv_first_fwd(xx, w, None, bias)
        else:
            K.conv_first_wgrad(xx, dy)
elif which == "upfirdn":
    from contrad_b200 import sg2_kernels as S
    fir = torch.tensor([1., 3., 3., 1.]); fir = (fir[None] * fir[:, None] / 64).cuda()
    xs = [torch.randn(192, 32, 32, 128, device="cuda") for _ in range(3)]
    for i in range(4):
        S.upfirdn2d(xs[i % 3], fir, 1, 1, (2, 2, 2, 2), round_out=True)
elif which == "bias_act":
    from contrad_b200 import sg2_kernels as S
    xs = [torch.randn(192, 32, 32, 128, device="cuda") for _ in range(3)]
    bias = torch.randn(128, device="cuda")
    for i in range(4):
        S.bias_act(xs[i % 3], bias, 0.2, 2 ** 0.5, round_out=True)
elif which == "augment_large":
    prm = torch.zeros(11, 48, device="cuda")
    prm[0] = 0.7; prm[1] = 0.8; prm[2] = 0.1; prm[3] = -0.1; prm[4] = 1.0; prm[5] = 1.0; prm[6] = 1.2; prm[7] = 0.05
    prm[8] = 1.1; prm[9] = 0.9
    xs = [torch.rand(48, 3, 512, 512, device="cuda") for _ in range(3)]
    for i in range(4):
        K.augment_simclr_large_fwd(xs[i % 3], prm, 0)
torch.cuda.synchronize()
print("done", which)
