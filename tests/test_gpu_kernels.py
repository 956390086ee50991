"""GPU parity tests of the individual sm_100a kernels, through the C ABI (ctypes).  Run with -m gpu."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import contrad_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from contrad_b200 import kernels
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return kernels


def _unpack(packed):
    return {k: packed[i] for i, k in enumerate(O.PARAM_FIELDS)}


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


# ---------------------------------------------------------------------------------- augmentation
def test_augment_matches_reference_golden(K, golden_dir):
    """fused kernel vs outputs of the unmodified reference chain (fwd) and its autograd (bwd)."""
    fx = _load(golden_dir, "augment_simclr.pt")
    for case in fx["cases"]:
        x, dy, params = case["x"].cuda(), case["dy"].cuda(), case["params"].cuda()
        y = K.augment_simclr_fwd(x, params, case["order"])
        dx = K.augment_simclr_bwd(x, dy, params, case["order"])
        assert torch.allclose(y.cpu(), case["y"], atol=2e-5, rtol=0), (y.cpu() - case["y"]).abs().max()
        assert torch.allclose(dx.cpu(), case["dx"], atol=1e-4, rtol=1e-4), (dx.cpu() - case["dx"]).abs().max()


@pytest.mark.parametrize("B,size,seed", [(96, 32, 0), (33, 32, 1), (7, 64, 2), (5, 16, 3), (1, 32, 4)])
def test_augment_matches_oracle_random(K, B, size, seed):
    import numpy as np
    np.random.seed(seed); torch.manual_seed(seed)
    x = torch.rand(B, 3, size, size)
    dy = torch.randn(B, 3, size, size)
    params, order = O.sample_simclr_params(B, size, size)
    xr = x.clone().requires_grad_(True)
    yr = O.augment_simclr(xr, params, order)
    (yr * dy).sum().backward()
    packed = O.pack_params(params).cuda()
    y = K.augment_simclr_fwd(x.cuda(), packed, order)
    dx = K.augment_simclr_bwd(x.cuda(), dy.cuda(), packed, order)
    assert torch.allclose(y.cpu(), yr.detach(), atol=2e-5, rtol=0)
    # a pixel whose contrast output sits within rounding of the clamp boundary may flip its mask
    bad = ((dx.cpu() - xr.grad).abs() > 1e-4 + 1e-4 * xr.grad.abs()).float().mean()
    assert bad < 2e-3, bad


def test_augment_full_size_properties(K):
    """BASELINE config-2 size (B=1536): identity parameters reproduce the input bit-exactly; a pure
    flip is the exact mirror; gray output has equal channels; range stays in [0,1]."""
    B = 1536
    torch.manual_seed(0)
    x = torch.rand(B, 3, 32, 32, device="cuda")
    ident = torch.zeros(11, B, device="cuda")
    ident[0] = 1; ident[1] = 1; ident[4] = 1; ident[6] = 1; ident[8] = 1; ident[9] = 1
    assert torch.equal(K.augment_simclr_fwd(x, ident, 0), x)
    flip = ident.clone(); flip[4] = -1
    assert torch.equal(K.augment_simclr_fwd(x, flip, 1), torch.flip(x, dims=[3]))
    gray = ident.clone(); gray[10] = 1
    y = K.augment_simclr_fwd(x, gray, 0)
    assert torch.equal(y[:, 0], y[:, 1]) and torch.equal(y[:, 1], y[:, 2])
    import numpy as np
    np.random.seed(5); torch.manual_seed(5)
    params, order = O.sample_simclr_params(B, 32, 32)
    y = K.augment_simclr_fwd(x, O.pack_params(params).cuda(), order)
    assert float(y.min()) >= 0.0 and float(y.max()) <= 1.0 and torch.isfinite(y).all()
    # linearity of the backward in dy (size-independent property)
    p = O.pack_params(params).cuda()
    d1 = torch.randn_like(x); d2 = torch.randn_like(x)
    lhs = K.augment_simclr_bwd(x, d1 + 2 * d2, p, order)
    rhs = K.augment_simclr_bwd(x, d1, p, order) + 2 * K.augment_simclr_bwd(x, d2, p, order)
    assert torch.allclose(lhs, rhs, atol=2e-4, rtol=1e-4)


# ---------------------------------------------------------------------------------- tensor-core GEMM
@pytest.mark.parametrize("M,N,K_", [(128, 128, 32), (256, 128, 64), (384, 64, 96), (1000, 32, 512),
                                    (1536, 1536, 8192), (77, 128, 128)])
def test_gemm_nt_tf32(K, M, N, K_):
    torch.manual_seed(M + N + K_)
    a = K.round_tf32(torch.randn(M, K_, device="cuda"))
    b = K.round_tf32(torch.randn(N, K_, device="cuda") * 0.05)
    bias = torch.randn(N, device="cuda")
    out = K.gemm_nt(a, b, bias, slope=0.1)
    ref = F.leaky_relu(a.double() @ b.double().t() + bias.double(), 0.1).float()
    err = (out - ref).abs().max() / ref.abs().max()
    assert err < 2e-5, err


def test_gemm_nt_strided_views(K):
    torch.manual_seed(3)
    big = K.round_tf32(torch.randn(300, 1536, device="cuda"))
    a = big[:, 512:1024]
    b = K.round_tf32(torch.randn(128, 512, device="cuda") * 0.05)
    outbuf = torch.zeros(300, 256, device="cuda")
    K.gemm_nt(a, b, None, out=outbuf[:, 128:])
    ref = (a.double() @ b.double().t()).float()
    assert torch.allclose(outbuf[:, 128:], ref, atol=1e-4, rtol=1e-4)
    assert torch.count_nonzero(outbuf[:, :128]) == 0


CONV_CASES = [
    # B, H, Cin, Cout, ks, stride
    (8, 32, 64, 128, 4, 2), (8, 16, 128, 128, 3, 1), (8, 16, 128, 256, 4, 2), (6, 8, 256, 256, 3, 1),
    (16, 8, 256, 512, 4, 2), (16, 4, 512, 512, 3, 1), (3, 4, 64, 64, 3, 1), (5, 8, 32, 32, 4, 2),
]


@pytest.mark.parametrize("B,H,Cin,Cout,ks,stride", CONV_CASES)
def test_conv_fwd_and_dgrad(K, B, H, Cin, Cout, ks, stride):
    torch.manual_seed(B * 1000 + H + Cin + Cout)
    x = K.round_tf32(torch.randn(B, Cin, H, H, device="cuda"))
    w = K.round_tf32(torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.05)
    bias = torch.randn(Cout, device="cuda") * 0.1
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    y = K.conv2d_nhwc_fwd(x_nhwc, K.pack_fwd_weight(w), bias, ks, stride, slope=0.1)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), bias.double(), stride=stride, padding=1), 0.1)
    ref_nhwc = ref.permute(0, 2, 3, 1).float()
    err = (y - ref_nhwc).abs().max() / ref_nhwc.abs().max()
    assert err < 2e-5, ("fwd", err)
    # data gradient, fused with lrelu'(input activation)
    dy = K.round_tf32(torch.randn(B, Cout, H // stride, H // stride, device="cuda"))
    act = torch.randn(B, H, H, Cin, device="cuda")
    dx = K.conv2d_nhwc_dgrad(dy.permute(0, 2, 3, 1).contiguous(), K.pack_dgrad_weight(w, stride),
                             (B, H, H, Cin), ks, stride, act_in=act, slope=0.1)
    ref_dx = torch.nn.grad.conv2d_input((B, Cin, H, H), w.double(), dy.double(), stride=stride, padding=1)
    ref_dx = (ref_dx.permute(0, 2, 3, 1) * torch.where(act > 0, 1.0, 0.1).double()).float()
    err = (dx - ref_dx).abs().max() / ref_dx.abs().max()
    assert err < 2e-5, ("dgrad", err)
