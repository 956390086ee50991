"""GPU parity tests of the individual sm_100a kernels, through the C ABI (ctypes).  Run with -m gpu."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import contrad_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)          # `gin` shim for contrad_b200.augment
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


@pytest.mark.parametrize("B,size,seed", [(96, 32, 0), (33, 32, 1), (7, 64, 2), (5, 16, 3), (1, 32, 4), (2500, 32, 5),
                                         (700, 64, 6)])
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


def test_augment_builds_are_bit_identical(K, golden_dir, monkeypatch):
    """CB200_AUGMENT_V=1 | 2 (csrc/augment.cu: per-thread parameter loads / index tap tables vs parameters staged through a
    shared-memory ring by twelve loader lanes / byte-offset tap tables / guard-free reciprocals): same pixel arithmetic, so
    the outputs must be bit-equal - on the reference fixtures, on grids where every CTA walks many images (with the
    per-image jitter order of the graph-replay mode), and at 64 x 64; whichever build is the default is thereby tied to
    the fixtures through the other one as well."""
    def both(x, p, order):
        monkeypatch.setenv("CB200_AUGMENT_V", "1")
        y1 = K.augment_simclr_fwd(x, p, order)
        monkeypatch.setenv("CB200_AUGMENT_V", "2")
        y2 = K.augment_simclr_fwd(x, p, order)
        monkeypatch.delenv("CB200_AUGMENT_V")
        return y1, y2

    for case in _load(golden_dir, "augment_simclr.pt")["cases"]:
        y1, y2 = both(case["x"].cuda(), case["params"].cuda(), case["order"])
        assert torch.equal(y1, y2)
        assert torch.allclose(y2.cpu(), case["y"], atol=2e-5, rtol=0)
    import numpy as np
    for B, size in ((1, 32), (889, 32), (1536, 32), (20000, 32), (700, 64)):
        np.random.seed(B); torch.manual_seed(B)
        x = torch.rand(B, 3, size, size, device="cuda")
        params, order = O.sample_simclr_params(B, size, size)
        packed = O.pack_params(params).cuda()
        y1, y2 = both(x, packed, order)
        assert torch.equal(y1, y2), (B, size, float((y1 - y2).abs().max()))
        row = torch.cat([packed, (torch.rand(1, B, device="cuda") < 0.5).float()])
        y1, y2 = both(x, row, -1)
        assert torch.equal(y1, y2), (B, size, "row 11")


def test_augment_large_path_matches_reference_golden(K, golden_dir):
    """The global-memory kernels (images > 64x64) forced onto the reference-generated small fixtures."""
    fx = _load(golden_dir, "augment_simclr.pt")
    for case in fx["cases"]:
        x, dy, params = case["x"].cuda(), case["dy"].cuda(), case["params"].cuda()
        y, means = K.augment_simclr_large_fwd(x, params, case["order"])
        dx = K.augment_simclr_large_bwd(x, dy, params, case["order"], means)
        assert torch.allclose(y.cpu(), case["y"], atol=2e-5, rtol=0), (y.cpu() - case["y"]).abs().max()
        assert torch.allclose(dx.cpu(), case["dx"], atol=1e-4, rtol=1e-4), (dx.cpu() - case["dx"]).abs().max()


@pytest.mark.parametrize("B,H,W,seed", [(3, 128, 128, 0), (2, 96, 80, 1), (5, 256, 256, 2), (2, 512, 512, 3)])
def test_augment_large_path_matches_oracle(K, B, H, W, seed):
    """Large images (up to the 512x512 of BASELINE config 5) against the CPU oracle, forward and backward, through the
    autograd Function (which picks the large path by size)."""
    import numpy as np
    from contrad_b200.functional import AugmentSimCLRFn
    np.random.seed(seed); torch.manual_seed(seed)
    x = torch.rand(B, 3, H, W)
    dy = torch.randn(B, 3, H, W)
    params, order = O.sample_simclr_params(B, H, W)
    xr = x.clone().requires_grad_(True)
    yr = O.augment_simclr(xr, params, order)
    (yr * dy).sum().backward()
    xg = x.cuda().requires_grad_(True)
    y = AugmentSimCLRFn.apply(xg, O.pack_params(params).cuda(), order)
    (y * dy.cuda()).sum().backward()
    assert torch.allclose(y.detach().cpu(), yr.detach(), atol=3e-5, rtol=0), (y.detach().cpu() - yr.detach()).abs().max()
    bad = ((xg.grad.cpu() - xr.grad).abs() > 2e-4 + 2e-4 * xr.grad.abs()).float().mean()
    assert bad < 2e-3, bad


def _unpack_hq(case):
    hq = case["hq"]
    on = hq["blur_on"].cuda()
    cut = None
    if "cut_on" in hq:
        cut = torch.stack([hq["cut_on"], hq["h_center"].float(), hq["w_center"].float()]).cuda()
    return on, cut


def test_augment_hq_matches_reference_golden(K, golden_dir):
    """simclr_hq / simclr_hq_cutout: fused chain -> separable Gaussian blur -> CutOut against the reference chain
    (dense kornia-style blur), forward and backward through the autograd Functions."""
    from contrad_b200.augment.layers import GaussianBlur, gaussian_taps
    from contrad_b200.functional import AugmentSimCLRFn, CutOutFn, GaussianBlurFn
    fx = _load(golden_dir, "augment_hq.pt")
    for case in fx["cases"]:
        x = case["x"].cuda().requires_grad_(True)
        on, cut = _unpack_hq(case)
        taps = gaussian_taps(GaussianBlur.kernel_size(x.shape[2]), case["hq"]["sigma"]).cuda()
        y = AugmentSimCLRFn.apply(x, case["params"].cuda(), case["order"])
        y = GaussianBlurFn.apply(y, taps, on)
        if cut is not None:
            y = CutOutFn.apply(y, cut, case["length"])
        (y * case["dy"].cuda()).sum().backward()
        assert torch.allclose(y.detach().cpu(), case["y"], atol=2e-5, rtol=0), (y.detach().cpu() - case["y"]).abs().max()
        assert torch.allclose(x.grad.cpu(), case["dx"], atol=1e-4, rtol=1e-4), (x.grad.cpu() - case["dx"]).abs().max()


@pytest.mark.parametrize("B,H,W,seed", [(4, 512, 512, 0), (3, 96, 80, 1), (7, 32, 32, 2)])
def test_gaussian_blur_and_cutout_vs_oracle(K, B, H, W, seed):
    """The 51-tap blur of the 512x512 configs (and odd shapes): separable kernel vs the oracle's dense k x k reflect
    correlation, its adjoint vs autograd, CutOut(255) vs the oracle mask."""
    from contrad_b200.augment.layers import GaussianBlur, gaussian_taps
    torch.manual_seed(seed)
    x, dy = torch.rand(B, 3, H, W), torch.randn(B, 3, H, W)
    sigma = 0.1 + 1.9 * float(torch.rand(()))
    on = (torch.rand(B) > 0.4).float()
    on[0] = 1.0
    xr = x.clone().requires_grad_(True)
    ref = O._blend(xr, O.gaussian_blur(xr, sigma), on)
    (ref * dy).sum().backward()
    taps = gaussian_taps(GaussianBlur.kernel_size(H), sigma).cuda()
    got = K.gaussian_blur(x.cuda(), taps, on.cuda())
    got_dx = K.gaussian_blur(dy.cuda(), taps, on.cuda(), adjoint=True)
    assert torch.allclose(got.cpu(), ref.detach(), atol=3e-6, rtol=0), (got.cpu() - ref.detach()).abs().max()
    assert torch.allclose(got_dx.cpu(), xr.grad, atol=2e-5, rtol=1e-5), (got_dx.cpu() - xr.grad).abs().max()
    length = 255 if H >= 512 else 15
    hc, wc = torch.randint(H, (B,)), torch.randint(W, (B,))
    params = torch.stack([on, hc.float(), wc.float()]).cuda()
    want = O._blend(x, O.cutout(x, hc, wc, length), on)
    assert torch.equal(K.cutout(x.cuda(), params, length).cpu(), want)


def test_augment_hq_modules_run(K):
    """get_augment('simclr_hq_cutout') end to end (gin-bound constructor arguments, device draws, staging)."""
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    import gin
    import numpy as np
    from contrad_b200.augment import get_augment
    gin.clear_config()
    gin.parse_config("RandomResizeCropLayer.scale = (0.08, 1.0)\nColorJitterLayer.brightness = 0.8\n"
                     "ColorJitterLayer.contrast = 0.8\nColorJitterLayer.saturation = 0.8\nColorJitterLayer.hue = 0.2\n"
                     "GaussianBlur.sigma_range = (0.1, 2.0)\nCutOut.length = 15\n")
    np.random.seed(0); torch.manual_seed(0)
    for mode in ("simclr_hq", "simclr_hq_cutout"):
        aug = get_augment(mode=mode).cuda()
        x = torch.rand(64, 3, 64, 64, device="cuda", requires_grad=True)
        y = aug(x)
        y.sum().backward()
        assert y.shape == x.shape and torch.isfinite(y).all() and float(y.min()) >= 0 and float(y.max()) <= 1
        assert torch.isfinite(x.grad).all() and float(x.grad.abs().sum()) > 0
    gin.clear_config()


def test_augment_per_image_order_row(K):
    """order = -1: the jitter order comes per image from row 11 of the parameter block (CUDA-graph replay path);
    forward and backward equal the launch-wide order applied image by image."""
    import numpy as np
    np.random.seed(8); torch.manual_seed(8)
    B = 40
    x = torch.rand(B, 3, 32, 32).cuda()
    dy = torch.randn(B, 3, 32, 32).cuda()
    params, _ = O.sample_simclr_params(B, 32, 32)
    packed = O.pack_params(params)
    packed[5] = 1.0                                   # colour jitter on everywhere, so the order matters
    orders = (torch.arange(B) % 3 == 0).float()
    p12 = torch.cat([packed, orders.view(1, B)]).cuda()
    y = K.augment_simclr_fwd(x, p12, -1)
    dx = K.augment_simclr_bwd(x, dy, p12, -1)
    ys = [K.augment_simclr_fwd(x, packed.cuda(), o) for o in (0, 1)]
    dxs = [K.augment_simclr_bwd(x, dy, packed.cuda(), o) for o in (0, 1)]
    sel = orders.bool().cuda().view(B, 1, 1, 1)
    assert torch.equal(y, torch.where(sel, ys[1], ys[0]))
    assert torch.allclose(dx, torch.where(sel, dxs[1], dxs[0]), atol=1e-5, rtol=1e-5)   # smem atomics reorder
    assert not torch.equal(ys[0], ys[1])


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
                                    (1536, 1536, 8192), (77, 128, 128), (40000, 256, 64), (37900, 128, 96),
                                    # split-K shapes (heads at 64 images per rank: D step 192 rows, G step 64 rows)
                                    (192, 1536, 8192), (64, 1536, 8192), (192, 512, 1536), (64, 32, 512), (192, 8192, 1024)])
def test_gemm_nt_tf32(K, M, N, K_):
    torch.manual_seed(M + N + K_)
    a = K.round_tf32(torch.randn(M, K_, device="cuda"))
    b = K.round_tf32(torch.randn(N, K_, device="cuda") * 0.05)
    bias = torch.randn(N, device="cuda")
    colsum = torch.full((N,), float("nan"), device="cuda")
    out = K.gemm_nt(a, b, bias, slope=0.1, colsum=colsum)
    ref = F.leaky_relu(a.double() @ b.double().t() + bias.double(), 0.1).float()
    err = (out - ref).abs().max() / ref.abs().max()
    assert err < 5e-5, err
    want = out.double().sum(0)
    assert ((colsum.double() - want).abs().max() / out.double().abs().sum(0).max()) < 1e-6


def test_split_k_is_deterministic_and_rearms(K):
    """The split-K path (tile list < GPU): bit-identical results call after call (partials are added in split order by
    the last CTA to arrive; the per-tile arrival counters are left at zero)."""
    torch.manual_seed(9)
    a = K.round_tf32(torch.randn(192, 8192, device="cuda"))
    b = K.round_tf32(torch.randn(1536, 8192, device="cuda") * 0.05)
    outs = [K.gemm_nt(a, b, None, slope=0.1, round_out=True).clone() for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = K.round_tf32(F.leaky_relu(a.double() @ b.double().t(), 0.1).float())
    assert ((outs[0] - ref).abs().max() / ref.abs().max()) < 1e-3          # TF32 rounding of the output included
    x = K.round_tf32(torch.randn(64, 4, 4, 512, device="cuda"))
    w = K.round_tf32(torch.randn(512, 512, 3, 3, device="cuda") * 0.05)
    ys = [K.conv2d_nhwc_fwd(x, K.pack_fwd_weight(w), None, 3, 1).clone() for _ in range(3)]
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2])


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
    # B, H, Cin, Cout, ks, stride   (the last four have >= 296 M tiles: CTA-pair / cta_group::2 kernels)
    (8, 32, 64, 128, 4, 2), (8, 16, 128, 128, 3, 1), (8, 16, 128, 256, 4, 2), (6, 8, 256, 256, 3, 1),
    (16, 8, 256, 512, 4, 2), (16, 4, 512, 512, 3, 1), (3, 4, 64, 64, 3, 1), (5, 8, 32, 32, 4, 2),
    (301, 16, 128, 128, 3, 1), (600, 16, 128, 256, 4, 2), (1201, 8, 256, 256, 3, 1), (1200, 32, 64, 128, 4, 2),
    # thin data gradients (N = Cin = 32: the 3 image channels padded to 32) on the persistent BN = 32 variant
    (700, 16, 32, 64, 3, 1), (300, 32, 32, 64, 4, 2),
    # split-K shapes: the deep layers at 64 images per rank (D step B = 192, G step B = 64)
    (192, 4, 512, 512, 3, 1), (192, 8, 256, 512, 4, 2), (64, 8, 256, 256, 3, 1), (64, 16, 128, 256, 4, 2),
    (64, 4, 512, 512, 3, 1), (64, 8, 256, 512, 4, 2),
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
    colsum = torch.full((Cin,), float("nan"), device="cuda")            # zeroed inside the call
    dx = K.conv2d_nhwc_dgrad(dy.permute(0, 2, 3, 1).contiguous(), K.pack_dgrad_weight(w, stride),
                             (B, H, H, Cin), ks, stride, act_in=act, slope=0.1, colsum=colsum)
    ref_dx = torch.nn.grad.conv2d_input((B, Cin, H, H), w.double(), dy.double(), stride=stride, padding=1)
    ref_dx = (ref_dx.permute(0, 2, 3, 1) * torch.where(act > 0, 1.0, 0.1).double()).float()
    err = (dx - ref_dx).abs().max() / ref_dx.abs().max()
    assert err < 2e-5, ("dgrad", err)
    # fused column sums (= bias gradient of the producing layer) are the sums of the values actually stored
    want = dx.double().sum(dim=(0, 1, 2))
    scale = dx.double().abs().sum(dim=(0, 1, 2)).max()
    assert ((colsum.double() - want).abs().max() / scale) < 1e-6, ("colsum", (colsum.double() - want).abs().max(), scale)


@pytest.mark.parametrize("B,H,Cin,Cout,ks,stride", CONV_CASES[:6] + [(4, 8, 128, 128, 3, 1), (130, 4, 64, 128, 4, 2)])
def test_conv_wgrad(K, B, H, Cin, Cout, ks, stride):
    if Cout % 128:
        pytest.skip("wgrad tiles Cout by 128")
    torch.manual_seed(B + H + Cin)
    x = K.round_tf32(torch.randn(B, Cin, H, H, device="cuda"))
    dy = K.round_tf32(torch.randn(B, Cout, H // stride, H // stride, device="cuda"))
    dw = K.conv2d_nhwc_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), ks, stride)
    ref = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, ks, ks), dy.double(), stride=stride, padding=1)
    ref = K.pack_fwd_weight(ref).float()
    err = (dw - ref).abs().max() / ref.abs().max()
    assert err < 5e-5, err


@pytest.mark.parametrize("M,N,K_", [(1536, 1536, 8192), (100, 128, 512), (64, 128, 32), (1000, 256, 96)])
def test_gemm_tn_wgrad(K, M, N, K_):
    torch.manual_seed(M + N)
    dy = K.round_tf32(torch.randn(M, N, device="cuda"))
    x = K.round_tf32(torch.randn(M, K_, device="cuda"))
    dw = K.gemm_tn_wgrad(dy, x)
    ref = (dy.double().t() @ x.double()).float()
    err = (dw - ref).abs().max() / ref.abs().max()
    assert err < 5e-5, err


def test_spectral_norm_matches_reference_golden(K, golden_dir):
    fx = _load(golden_dir, "spectral_norm.pt")
    for name, rec in fx.items():
        w = rec["weight_orig"].cuda()
        u, v = rec["u0"].cuda().clone(), rec["v0"].cuda().clone()
        sigma = torch.zeros(2, device="cuda")
        K.sn_power_iter(w, u, v, sigma, training=True)
        assert torch.allclose(u.cpu(), rec["u1"], atol=1e-5) and torch.allclose(v.cpu(), rec["v1"], atol=1e-5)
        w4 = w if w.dim() == 4 else w.view(w.shape[0], w.shape[1], 1, 1)
        Cout, Cin, KH, KW = w4.shape
        fwd = torch.zeros(Cout, KH * KW * Cin, device="cuda")
        K.sn_pack_weights(w4, sigma, fwd=fwd, ld_fwd=fwd.shape[1], round_out=False)
        w_hat = fwd.view(Cout, KH, KW, Cin).permute(0, 3, 1, 2).reshape(rec["w_hat"].shape)
        assert torch.allclose(w_hat.cpu(), rec["w_hat"], atol=1e-6, rtol=1e-5)
        # backward: feed dL/dW_hat of the fixture's loss (sum y^2) computed by torch on the fixture tensors
        wh = rec["w_hat"].clone().requires_grad_(True)
        if name == "conv":
            y = F.conv2d(rec["x"], wh, rec["bias"], padding=1)
        else:
            y = F.linear(rec["x"], wh, rec["bias"])
        y.pow(2).sum().backward()
        g4 = wh.grad if wh.grad.dim() == 4 else wh.grad.view(Cout, Cin, 1, 1)
        g_packed = g4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().cuda()
        dw = torch.empty_like(w4)
        K.sn_weight_bwd(g_packed, g_packed.shape[1], w4, u, v, sigma, dw)
        assert torch.allclose(dw.cpu().view(rec["grad_weight_orig"].shape), rec["grad_weight_orig"], atol=1e-4, rtol=1e-3)


def test_sn_pack_dgrad_layouts(K):
    torch.manual_seed(0)
    w = torch.randn(64, 32, 4, 4, device="cuda")
    dg = torch.zeros(4 * 32, 4 * 64, device="cuda")
    K.sn_pack_weights(w, None, dgrad=dg, dgrad_mode=2, round_out=False)
    assert torch.equal(dg, K.pack_dgrad_weight(w, 2))
    w3 = torch.randn(64, 32, 3, 3, device="cuda")
    dg = torch.zeros(32, 9 * 64, device="cuda")
    fw = torch.zeros(64, 9 * 32, device="cuda")
    K.sn_pack_weights(w3, None, fwd=fw, ld_fwd=9 * 32, dgrad=dg, dgrad_mode=1, round_out=False)
    assert torch.equal(dg, K.pack_dgrad_weight(w3, 1)) and torch.equal(fw, K.pack_fwd_weight(w3))
    # transposed linear pack (mode 3): rows (h,w,c), columns offset
    wl = torch.randn(48, 32 * 4 * 4, device="cuda")
    t = torch.zeros(4 * 4 * 32, 100, device="cuda")
    K.sn_pack_weights(wl.view(48, 32, 4, 4), None, dgrad=t, dgrad_mode=3, ldt=100, col0=20, round_out=False)
    ref = wl.view(48, 32, 4, 4).permute(2, 3, 1, 0).reshape(512, 48)
    assert torch.equal(t[:, 20:68], ref) and torch.count_nonzero(t[:, :20]) == 0


@pytest.mark.parametrize("B,H", [(8, 32), (3, 16), (2, 64)])
def test_conv_first_layer(K, B, H):
    torch.manual_seed(B)
    x = torch.rand(B, 3, H, H, device="cuda")
    w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
    bias = torch.randn(64, device="cuda") * 0.1
    sigma = torch.tensor([2.0, 0.5], device="cuda")
    y = K.conv_first_fwd(x, w, sigma, bias, slope=0.1, round_out=False)
    ref = F.leaky_relu(F.conv2d(x.double() * 2 - 1, w.double() * 0.5, bias.double(), padding=1), 0.1)
    assert torch.allclose(y, ref.permute(0, 2, 3, 1).float(), atol=1e-5, rtol=1e-5)
    dy = torch.randn(B, H, H, 64, device="cuda")
    dw, db = K.conv_first_wgrad(x, dy)
    ref_dw = torch.nn.grad.conv2d_weight(x.double() * 2 - 1, (64, 3, 3, 3), dy.permute(0, 3, 1, 2).double(), padding=1)
    assert torch.allclose(dw.view(64, 3, 3, 3), ref_dw.float(), atol=1e-3, rtol=1e-4)
    assert torch.allclose(db, dy.sum(dim=(0, 1, 2)), atol=1e-3, rtol=1e-4)
    # data gradient through the padded tensor-core path
    wt = torch.zeros(32, 9 * 64, device="cuda")
    K.sn_pack_weights(w, sigma, dgrad=wt, dgrad_mode=1, round_out=True)
    dyr = K.round_tf32(dy)
    dpad = K.conv2d_nhwc_dgrad(dyr, wt, (B, H, H, 32), 3, 1)
    dx = K.conv_first_dgrad_finish(dpad)
    ref_dx = 2 * torch.nn.grad.conv2d_input((B, 3, H, H), K.round_tf32(w * 0.5).double(), dyr.permute(0, 3, 1, 2).double(), padding=1)
    assert torch.allclose(dx, ref_dx.float(), atol=1e-4, rtol=1e-4)


# ---------------------------------------------------------------------------------- losses
def test_contrastive_matches_reference_golden(K, golden_dir):
    fx = _load(golden_dir, "contrastive.pt")
    for case in fx["cases"]:
        n = case["n"]
        a, b, c = case["a"].cuda(), case["b"].cuda(), case["c"].cuda()
        one = torch.ones(1, device="cuda")
        z = torch.cat([a, b], 0)
        loss, lse = K.contrastive_fwd(z, n, 0, 0.1)
        assert abs(float(loss) - case["nt_xent"]) < 2e-5 * abs(case["nt_xent"])
        dz = K.contrastive_bwd(z, n, 0, 0.1, lse, one)
        ref = torch.cat(case["nt_xent_grads"], 0)
        assert torch.allclose(dz.cpu(), ref, atol=2e-6, rtol=2e-4)
        z3 = torch.cat([a, b, c], 0)
        loss, lse = K.contrastive_fwd(z3, n, 1, 0.1)
        assert abs(float(loss) - case["supcon"]) < 2e-5 * abs(case["supcon"])
        dz = K.contrastive_bwd(z3, n, 1, 0.1, lse, one * 0.5)
        ref = torch.cat(case["supcon_grads"], 0) * 0.5
        assert torch.allclose(dz.cpu(), ref, atol=2e-6, rtol=2e-4)
        loss, _ = K.contrastive_fwd(z, n, 0, 0.5)
        assert abs(float(loss) - case["nt_xent_t05"]) < 2e-5


@pytest.mark.parametrize("n", [64, 512, 100])
def test_contrastive_full_size_vs_oracle(K, n):
    torch.manual_seed(n)
    a, b, c = (F.normalize(torch.randn(n, 128, device="cuda")).requires_grad_(True) for _ in range(3))
    one = torch.ones(1, device="cuda")
    l1 = O.nt_xent(a, b, 0.1); g1 = torch.autograd.grad(l1, [a, b])
    l2 = O.supcon_fake(a, b, c, 0.1); g2 = torch.autograd.grad(l2, [a, b, c])
    z = torch.cat([a, b], 0).detach()
    loss, lse = K.contrastive_fwd(z, n, 0, 0.1)
    assert abs(float(loss) - float(l1)) < 1e-5 * abs(float(l1))
    assert torch.allclose(K.contrastive_bwd(z, n, 0, 0.1, lse, one), torch.cat(g1, 0), atol=1e-7, rtol=1e-3)
    z3 = torch.cat([a, b, c], 0).detach()
    loss, lse = K.contrastive_fwd(z3, n, 1, 0.1)
    assert abs(float(loss) - float(l2)) < 1e-5 * abs(float(l2))
    assert torch.allclose(K.contrastive_bwd(z3, n, 1, 0.1, lse, one), torch.cat(g2, 0), atol=1e-7, rtol=1e-3)


def test_rownorm_and_gan_losses(K):
    torch.manual_seed(0)
    big = torch.randn(300, 384, device="cuda")
    x = big[:, 128:256]
    xr = x.clone().requires_grad_(True)
    yr = F.normalize(xr)
    dy = torch.randn(300, 128, device="cuda")
    (yr * dy).sum().backward()
    y, inv = K.rownorm_fwd(x)
    assert torch.allclose(y, yr.detach(), atol=1e-6)
    assert torch.allclose(K.rownorm_bwd(dy, y, inv), xr.grad, atol=1e-5, rtol=1e-4)
    for kind in ("nonsat", "hinge", "wgan", "lsgan"):
        d = torch.randn(3 * 64, 1, device="cuda")
        dr = d.clone().requires_grad_(True)
        ref = O.gan_d_loss(dr[:64], dr[128:], kind)
        ref.backward()
        out, g_r, g_g = K.gan_d_loss(d[:64, 0], d[128:, 0], kind)
        assert abs(float(out[0]) - float(ref)) < 1e-5
        assert torch.allclose(g_r, dr.grad[:64, 0], atol=1e-6) and torch.allclose(g_g, dr.grad[128:, 0], atol=1e-6)
        assert abs(float(out[1]) - float(d[:64].mean())) < 1e-6 and abs(float(out[2]) - float(d[128:].mean())) < 1e-6
        dr = d[:64].clone().requires_grad_(True)
        ref = O.gan_g_loss(dr, kind); ref.backward()
        out, g = K.gan_g_loss(d[:64, 0], kind)
        assert abs(float(out[0]) - float(ref)) < 1e-5 and torch.allclose(g, dr.grad[:, 0], atol=1e-6)
    for M, N in ((5000, 192), (512, 8192), (1536, 1536), (100003, 64), (37, 12), (3000, 1), (1, 256), (70000, 8)):
        x = torch.randn(M, N, device="cuda")
        assert torch.allclose(K.colsum(x), x.double().sum(0).float(), atol=2e-3, rtol=1e-4), (M, N)


def test_sn_batched_equals_single(K):
    """The batched power-iteration / pack / backward launches give the single-layer results."""
    torch.manual_seed(0)
    shapes = [(64, 3, 3, 3), (128, 64, 4, 4), (512, 512, 3, 3), (512, 8192), (1, 512), (128, 512)]
    ws = [torch.randn(*s, device="cuda") * 0.05 for s in shapes]
    us = [F.normalize(torch.randn(s[0], device="cuda"), dim=0) for s in shapes]
    vs = [F.normalize(torch.randn(w.numel() // w.shape[0], device="cuda"), dim=0) for w in ws]
    u1, v1 = [u.clone() for u in us], [v.clone() for v in vs]
    u2, v2 = [u.clone() for u in us], [v.clone() for v in vs]
    s1 = [torch.zeros(2, device="cuda") for _ in ws]
    s2 = [torch.zeros(2, device="cuda") for _ in ws]
    for w, u, v, s in zip(ws, u1, v1, s1):
        K.sn_power_iter(w, u, v, s, training=True)
    K.sn_power_iter_batched(list(zip(ws, u2, v2, s2)), training=True)
    for a, b in zip(u1 + v1 + s1, u2 + v2 + s2):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)
    # eval-mode sigma from the stored vectors
    s3 = [torch.zeros(2, device="cuda") for _ in ws]
    K.sn_power_iter_batched(list(zip(ws, u2, v2, s3)), training=False)
    for a, b in zip(s2, s3):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    # pack + backward on the 4x4 stride-2 layer and the big linear layer
    w = ws[1]
    fwd_a = torch.empty(128, 16 * 64, device="cuda"); dg_a = torch.empty(4 * 64, 4 * 128, device="cuda")
    fwd_b = torch.empty_like(fwd_a); dg_b = torch.empty_like(dg_a)
    K.sn_pack_weights(w, s1[1], fwd=fwd_a, ld_fwd=1024, dgrad=dg_a, dgrad_mode=2)
    K.sn_pack_batched([dict(w4=w, sigma=s1[1], fwd=fwd_b, ld_fwd=1024, dgrad=dg_b, dgrad_mode=2)])
    assert torch.equal(fwd_a, fwd_b) and torch.equal(dg_a, dg_b)
    g = torch.randn_like(fwd_a)
    dw_a = torch.empty_like(w); dw_b = torch.empty_like(w)
    K.sn_weight_bwd(g, 1024, w, u1[1], v1[1], s1[1], dw_a)
    K.sn_weight_bwd_batched([dict(dw_hat_packed=g, ld_fwd=1024, w4=w, u=u1[1], v=v1[1], sigma=s1[1], dw=dw_b)])
    assert torch.allclose(dw_a, dw_b, atol=1e-6, rtol=1e-4)


def test_fused_adam_matches_torch_adam(K):
    from contrad_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(512, 8192), (64, 3, 3, 3), (1, 512), (77,), (128, 64, 4, 4)]
    pa = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedAdam(pa, lr=2e-4, betas=(0.5, 0.999))
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.5, 0.999))
    for step in range(3):
        for a, b in zip(pa, pb):
            g = torch.randn_like(a) * (step + 1)
            a.grad = g.clone(); b.grad = g.clone()
        for o in (oa, ob):
            for grp in o.param_groups:
                grp["lr"] = 2e-4 * (step + 1) / 3
        oa.step(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, atol=1e-7, rtol=1e-5)
    sa, sb = oa.state_dict(), ob.state_dict()
    assert set(sa["state"][0].keys()) == set(sb["state"][0].keys())
    ob.load_state_dict(sa)          # checkpoints interchange


def test_tensor_core_contrastive_path(K, golden_dir):
    """functional.ContrastiveTCFn - similarity matrix on tcgen05 (error-compensated operands), softmax / CE as warp-shuffle
    row reductions - against the reference fixtures (2e-5) and against the fused SIMT kernels at the benchmark size
    (N = 512: R = 1024 / 1536 rows) and at N = 4096 (R = 8192 / 12288); timings of both paths go to gpurun_out."""
    import json
    from contrad_b200.functional import ContrastiveFn, ContrastiveTCFn
    fx = _load(golden_dir, "contrastive.pt")
    for case in fx["cases"]:
        n = case["n"]
        a, b, c = (case[k].cuda().requires_grad_(True) for k in ("a", "b", "c"))
        l1 = ContrastiveTCFn.apply(torch.cat([a, b]), n, 0, 0.1)
        g1 = torch.autograd.grad(l1, [a, b])
        assert abs(float(l1) - case["nt_xent"]) < 2e-5 * abs(case["nt_xent"])
        for g, w in zip(g1, case["nt_xent_grads"]):
            assert torch.allclose(g.cpu(), w, atol=2e-5 * float(w.abs().max()) + 1e-7, rtol=1e-4)
        l2 = ContrastiveTCFn.apply(torch.cat([a, b, c]), n, 1, 0.1)
        g2 = torch.autograd.grad(l2, [a, b, c])
        assert abs(float(l2) - case["supcon"]) < 2e-5 * abs(case["supcon"])
        for g, w in zip(g2, case["supcon_grads"]):
            assert torch.allclose(g.cpu(), w, atol=2e-5 * float(w.abs().max()) + 1e-7, rtol=1e-4)
    report = {}
    for n in (512, 4096):
        torch.manual_seed(n)
        for mode, rows in ((0, 2 * n), (1, 3 * n)):
            z = F.normalize(torch.randn(rows, 128, device="cuda")).requires_grad_(True)
            res = {}
            for name, fn in (("simt", ContrastiveFn), ("tc", ContrastiveTCFn)):
                loss = fn.apply(z, n, mode, 0.1)
                (g,) = torch.autograd.grad(loss, z)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    l = fn.apply(z, n, mode, 0.1)
                    torch.autograd.grad(l, z)
                e1.record(); torch.cuda.synchronize()
                res[name] = (float(loss), g.detach(), e0.elapsed_time(e1) / 5)
            assert abs(res["tc"][0] - res["simt"][0]) < 2e-5 * abs(res["simt"][0]), (n, mode, res["tc"][0], res["simt"][0])
            gerr = float((res["tc"][1] - res["simt"][1]).norm() / res["simt"][1].norm())
            assert gerr < 1e-4, (n, mode, gerr)
            report["N%d_mode%d" % (n, mode)] = {"simt_ms_fwd_bwd": res["simt"][2], "tc_ms_fwd_bwd": res["tc"][2], "grad_rel_l2": gerr}
    print(report)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "contrastive_paths.json"), "w") as f:
        json.dump(report, f, indent=1)
