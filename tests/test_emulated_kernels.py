"""CPU execution of the product's SIMT kernels: the asm-free files of contrad_b200/csrc are compiled for the host against
a CUDA emulation header (tests/emu) and driven through the product's own bindings, so kernel source, launch configuration
and Python glue are checked without a GPU.  The first tests replay kernels that ARE verified on the B200 against the same
fixtures (they validate the emulator); the others extend CPU coverage to the remaining SIMT kernel families."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import contrad_oracle as O
from tests.emu import emulated

pytestmark = pytest.mark.timeout(600)

_COMPAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
if _COMPAT not in __import__("sys").path:
    __import__("sys").path.append(_COMPAT)          # `gin` shim for contrad_b200.augment


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


# ------------------------------------------------------------------ emulator validation on hardware-verified kernels
def test_emulated_light_augmentations_match_reference_fixtures(golden_dir):
    from contrad_b200 import kernels as K
    fx = _load(golden_dir, "augment_aux.pt")
    with emulated():
        for case in fx["shift_flip"]:
            y = K.shift_flip(case["x"], case["params"], case["padding_mode"])
            assert torch.equal(y, case["y"]), (case["kind"], case["padding_mode"])
            dx = K.shift_flip(case["dy"], case["params"], case["padding_mode"], adjoint=True)
            assert torch.allclose(dx, case["dx"], atol=1e-6, rtol=1e-6)
        for case in fx["noise"]:
            assert torch.equal(K.noise_clamp_fwd(case["x"], case["noise"], case["sigma"]), case["y"])
            assert torch.equal(K.noise_clamp_bwd(case["x"], case["noise"], case["dy"], case["sigma"]), case["dx"])
        for case in _load(golden_dir, "diffaug.pt")["cases"]:
            flags = sum({"color": 1, "translation": 2, "cutout": 4}[s] for s in case["policy"].split(","))
            assert torch.allclose(K.diffaug(case["x"], case["params"], flags), case["y"], atol=2e-6, rtol=0)
            assert torch.allclose(K.diffaug(case["dy"], case["params"], flags, adjoint=True), case["dx"], atol=5e-6, rtol=1e-5)


def _rel(a, b):
    a, b = a.detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def test_emulated_blur_and_cutout_vs_oracle():
    """csrc/augment_hq.cu (GPU-verified): separable reflect-padded Gaussian + adjoint, CutOut, at odd small shapes."""
    from contrad_b200 import kernels as K
    from contrad_b200.augment.layers import GaussianBlur, gaussian_taps
    with emulated():
        for B, H, W, seed in ((3, 40, 36, 1), (4, 32, 32, 2)):
            torch.manual_seed(seed)
            x, dy = torch.rand(B, 3, H, W), torch.randn(B, 3, H, W)
            sigma = 0.1 + 1.9 * float(torch.rand(()))
            on = (torch.rand(B) > 0.4).float()
            on[0] = 1.0
            xr = x.clone().requires_grad_(True)
            ref = O._blend(xr, O.gaussian_blur(xr, sigma), on)
            (ref * dy).sum().backward()
            taps = gaussian_taps(GaussianBlur.kernel_size(H), sigma)
            assert torch.allclose(K.gaussian_blur(x, taps, on), ref.detach(), atol=3e-6, rtol=0)
            assert torch.allclose(K.gaussian_blur(dy, taps, on, adjoint=True), xr.grad, atol=2e-5, rtol=1e-5)
            hc, wc = torch.randint(H, (B,)), torch.randint(W, (B,))
            params = torch.stack([on, hc.float(), wc.float()])
            assert torch.equal(K.cutout(x, params, 15), O._blend(x, O.cutout(x, hc, wc, 15), on))


# ------------------------------------------------------------------ kernel families beyond the ones above
def test_emulated_spectral_norm_matches_reference_golden(golden_dir):
    """csrc/sn_weights.cu: power iteration, packing and backward against the torch.nn.utils.spectral_norm fixture."""
    from contrad_b200 import kernels as K
    fx = _load(golden_dir, "spectral_norm.pt")
    with emulated():
        for name, rec in fx.items():
            w = rec["weight_orig"].clone()
            u, v = rec["u0"].clone(), rec["v0"].clone()
            sigma = torch.zeros(2)
            K.sn_power_iter(w, u, v, sigma, training=True)
            assert torch.allclose(u, rec["u1"], atol=1e-5) and torch.allclose(v, rec["v1"], atol=1e-5)
            w4 = w if w.dim() == 4 else w.view(w.shape[0], w.shape[1], 1, 1)
            Cout, Cin, KH, KW = w4.shape
            fwd = torch.zeros(Cout, KH * KW * Cin)
            K.sn_pack_weights(w4, sigma, fwd=fwd, ld_fwd=fwd.shape[1], round_out=False)
            w_hat = fwd.view(Cout, KH, KW, Cin).permute(0, 3, 1, 2).reshape(rec["w_hat"].shape)
            assert torch.allclose(w_hat, rec["w_hat"], atol=1e-6, rtol=1e-5)
            wh = rec["w_hat"].clone().requires_grad_(True)
            y = F.conv2d(rec["x"], wh, rec["bias"], padding=1) if name == "conv" else F.linear(rec["x"], wh, rec["bias"])
            y.pow(2).sum().backward()
            g4 = wh.grad if wh.grad.dim() == 4 else wh.grad.view(Cout, Cin, 1, 1)
            g_packed = g4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
            dw = torch.empty_like(w4)
            K.sn_weight_bwd(g_packed, g_packed.shape[1], w4, u, v, sigma, dw)
            assert torch.allclose(dw.view(rec["grad_weight_orig"].shape), rec["grad_weight_orig"], atol=1e-4, rtol=1e-3)


def test_emulated_sn_pack_layouts_and_batched_launches():
    from contrad_b200 import kernels as K
    torch.manual_seed(0)
    with emulated():
        w = torch.randn(64, 32, 4, 4)
        dg = torch.zeros(4 * 32, 4 * 64)
        K.sn_pack_weights(w, None, dgrad=dg, dgrad_mode=2, round_out=False)
        assert torch.equal(dg, K.pack_dgrad_weight(w, 2))
        w3 = torch.randn(64, 32, 3, 3)
        dg, fw = torch.zeros(32, 9 * 64), torch.zeros(64, 9 * 32)
        K.sn_pack_weights(w3, None, fwd=fw, ld_fwd=9 * 32, dgrad=dg, dgrad_mode=1, round_out=False)
        assert torch.equal(dg, K.pack_dgrad_weight(w3, 1)) and torch.equal(fw, K.pack_fwd_weight(w3))
        wl = torch.randn(48, 32 * 4 * 4)
        t = torch.zeros(4 * 4 * 32, 100)
        K.sn_pack_weights(wl.view(48, 32, 4, 4), None, dgrad=t, dgrad_mode=3, ldt=100, col0=20, round_out=False)
        assert torch.equal(t[:, 20:68], wl.view(48, 32, 4, 4).permute(2, 3, 1, 0).reshape(512, 48)) and not t[:, :20].any()
        # batched == single (power iteration, eval-mode sigma, pack, backward)
        shapes = [(64, 3, 3, 3), (32, 16, 4, 4), (48, 640), (1, 96)]
        ws = [torch.randn(*s) * 0.05 for s in shapes]
        us = [F.normalize(torch.randn(s[0]), dim=0) for s in shapes]
        vs = [F.normalize(torch.randn(w.numel() // w.shape[0]), dim=0) for w in ws]
        u1, v1, u2, v2 = [u.clone() for u in us], [v.clone() for v in vs], [u.clone() for u in us], [v.clone() for v in vs]
        s1, s2, s3 = ([torch.zeros(2) for _ in ws] for _ in range(3))
        for w, u, v, s in zip(ws, u1, v1, s1):
            K.sn_power_iter(w, u, v, s, training=True)
        K.sn_power_iter_batched(list(zip(ws, u2, v2, s2)), training=True)
        for a, b in zip(u1 + v1 + s1, u2 + v2 + s2):
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)
        for i, (w, u, v, s) in enumerate(zip(ws, u1, v1, s1)):     # against the oracle's restatement of torch's hook
            uu, vv = us[i].clone(), vs[i].clone()
            w_hat = O.spectral_normalize(w, uu, vv, training=True)
            assert torch.allclose(u, uu, atol=1e-5) and torch.allclose(v, vv, atol=1e-5)
            assert torch.allclose(w / s[0], w_hat, atol=1e-5, rtol=1e-4) and abs(float(s[0] * s[1]) - 1.0) < 1e-5
        K.sn_power_iter_batched(list(zip(ws, u2, v2, s3)), training=False)
        for a, b in zip(s2, s3):
            assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
        w = ws[1]
        fwd_a, dg_a = torch.empty(32, 16 * 16), torch.empty(4 * 16, 4 * 32)
        fwd_b, dg_b = torch.empty_like(fwd_a), torch.empty_like(dg_a)
        K.sn_pack_weights(w, s1[1], fwd=fwd_a, ld_fwd=256, dgrad=dg_a, dgrad_mode=2)
        K.sn_pack_batched([dict(w4=w, sigma=s1[1], fwd=fwd_b, ld_fwd=256, dgrad=dg_b, dgrad_mode=2)])
        assert torch.equal(fwd_a, fwd_b) and torch.equal(dg_a, dg_b)
        g = torch.randn_like(fwd_a)
        dw_a, dw_b = torch.empty_like(w), torch.empty_like(w)
        K.sn_weight_bwd(g, 256, w, u1[1], v1[1], s1[1], dw_a)
        K.sn_weight_bwd_batched([dict(dw_hat_packed=g, ld_fwd=256, w4=w, u=u1[1], v=v1[1], sigma=s1[1], dw=dw_b)])
        assert torch.allclose(dw_a, dw_b, atol=1e-6, rtol=1e-4)


def test_emulated_contrastive_losses(golden_dir):
    """csrc/losses.cu: NT-Xent / supcon-fake forward + backward against the reference fixtures and, at a batch that spans
    several column tiles, against the oracle."""
    from contrad_b200 import kernels as K
    fx = _load(golden_dir, "contrastive.pt")
    one = torch.ones(1)
    with emulated():
        for case in fx["cases"]:
            n, a, b, c = case["n"], case["a"], case["b"], case["c"]
            z = torch.cat([a, b], 0)
            loss, lse = K.contrastive_fwd(z, n, 0, 0.1)
            assert abs(float(loss) - case["nt_xent"]) < 2e-5 * abs(case["nt_xent"])
            assert torch.allclose(K.contrastive_bwd(z, n, 0, 0.1, lse, one), torch.cat(case["nt_xent_grads"], 0), atol=2e-6, rtol=2e-4)
            z3 = torch.cat([a, b, c], 0)
            loss, lse = K.contrastive_fwd(z3, n, 1, 0.1)
            assert abs(float(loss) - case["supcon"]) < 2e-5 * abs(case["supcon"])
            assert torch.allclose(K.contrastive_bwd(z3, n, 1, 0.1, lse, one * 0.5), torch.cat(case["supcon_grads"], 0) * 0.5,
                                  atol=2e-6, rtol=2e-4)
            loss, _ = K.contrastive_fwd(z, n, 0, 0.5)
            assert abs(float(loss) - case["nt_xent_t05"]) < 2e-5
        n = 44
        torch.manual_seed(n)
        a, b, c = (F.normalize(torch.randn(n, 128)).requires_grad_(True) for _ in range(3))
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


@pytest.mark.parametrize("n", [1, 2, 3, 33])
def test_emulated_contrastive_edge_cases(n):
    """Collisions and degenerate batches of the contrastive losses (training/criterion.py:24-45, training/gan/contrad.py:8-32)
    against the oracle: a single pair (NT-Xent's softmax row then has exactly one entry besides the masked diagonal),
    duplicated embeddings (every off-diagonal similarity 1: the positives are indistinguishable from the negatives),
    antipodal views (similarity -1 at the positive), and a temperature small enough that the largest logit is +-100 - the
    log-sum-exp must stay finite.  supcon_fake with ONE fake has no positive: NaN in the reference, NaN here."""
    from contrad_b200 import kernels as K
    one = torch.ones(1)
    torch.manual_seed(100 + n)
    base = F.normalize(torch.randn(n, 128))
    cases = {"random": tuple(F.normalize(torch.randn(n, 128)) for _ in range(3)),
             "duplicates": (base[:1].expand(n, 128).contiguous(),) * 3,
             "identical views": (base, base.clone(), F.normalize(torch.randn(n, 128))),
             "antipodal views": (base, -base, F.normalize(torch.randn(n, 128)))}
    with emulated():
        for name, (a, b, c) in cases.items():
            for temp in (0.1, 0.01):
                ar, br, cr = (t.clone().requires_grad_(True) for t in (a, b, c))
                l1 = O.nt_xent(ar, br, temp); g1 = torch.autograd.grad(l1, [ar, br])
                z = torch.cat([a, b], 0)
                loss, lse = K.contrastive_fwd(z, n, 0, temp)
                assert torch.isfinite(loss) and torch.isfinite(lse).all(), (name, temp)
                assert abs(float(loss) - float(l1)) < 2e-5 * max(1.0, abs(float(l1))), (name, temp, float(loss), float(l1))
                g = K.contrastive_bwd(z, n, 0, temp, lse, one)
                assert torch.allclose(g, torch.cat(g1, 0), atol=2e-5 / temp * 0.1, rtol=1e-3), (name, temp)
                ar, br, cr = (t.clone().requires_grad_(True) for t in (a, b, c))
                l2 = O.supcon_fake(ar, br, cr, temp); g2 = torch.autograd.grad(l2, [ar, br, cr])
                z3 = torch.cat([a, b, c], 0)
                loss, lse = K.contrastive_fwd(z3, n, 1, temp)
                if n == 1:      # a single fake has no positive: the reference divides its mask row by 0 (contrad.py:26) -> NaN
                    assert torch.isnan(l2) and torch.isnan(loss).all(), (name, temp, float(loss), float(l2))
                    continue
                assert torch.isfinite(loss) and torch.isfinite(lse).all(), (name, temp)
                assert abs(float(loss) - float(l2)) < 2e-5 * max(1.0, abs(float(l2))), (name, temp, float(loss), float(l2))
                g = K.contrastive_bwd(z3, n, 1, temp, lse, one)
                assert torch.allclose(g, torch.cat(g2, 0), atol=2e-5 / temp * 0.1, rtol=1e-3), (name, temp)


def test_emulated_rownorm_gan_losses_colsum_lrelu():
    from contrad_b200 import kernels as K
    torch.manual_seed(0)
    with emulated():
        big = torch.randn(70, 384)
        x = big[:, 128:256]
        xr = x.clone().requires_grad_(True)
        yr = F.normalize(xr)
        dy = torch.randn(70, 128)
        (yr * dy).sum().backward()
        y, inv = K.rownorm_fwd(x)
        assert torch.allclose(y, yr.detach(), atol=1e-6)
        assert torch.allclose(K.rownorm_bwd(dy, y, inv), xr.grad, atol=1e-5, rtol=1e-4)
        for kind in ("nonsat", "hinge", "wgan", "lsgan"):
            d = torch.randn(3 * 40, 1)
            dr = d.clone().requires_grad_(True)
            ref = O.gan_d_loss(dr[:40], dr[80:], kind)
            ref.backward()
            out, g_r, g_g = K.gan_d_loss(d[:40, 0], d[80:, 0], kind)
            assert abs(float(out[0]) - float(ref)) < 1e-5
            assert torch.allclose(g_r, dr.grad[:40, 0], atol=1e-6) and torch.allclose(g_g, dr.grad[80:, 0], atol=1e-6)
            assert abs(float(out[1]) - float(d[:40].mean())) < 1e-6 and abs(float(out[2]) - float(d[80:].mean())) < 1e-6
            dr = d[:40].clone().requires_grad_(True)
            ref = O.gan_g_loss(dr, kind); ref.backward()
            out, g = K.gan_g_loss(d[:40, 0], kind)
            assert abs(float(out[0]) - float(ref)) < 1e-5 and torch.allclose(g, dr.grad[:, 0], atol=1e-6)
        for M, N in ((500, 192), (64, 1024), (37, 12), (300, 1), (1, 256), (2000, 8)):
            x = torch.randn(M, N)
            assert torch.allclose(K.colsum(x), x.double().sum(0).float(), atol=2e-3, rtol=1e-4), (M, N)
        act, g = torch.randn(33, 64), torch.randn(33, 64)
        assert torch.equal(K.lrelu_bwd(g, act, 0.1), torch.where(act > 0, g, 0.1 * g))


def test_emulated_batchnorm_and_generator_tail():
    """csrc/gen_ops.cu: train-mode BatchNorm + ReLU forward / backward (plain NHWC and the (c,h,w)->(h,w,c) remapped
    first layer), running statistics, the tanh stage - against torch autograd."""
    from contrad_b200 import kernels as K
    torch.manual_seed(3)
    with emulated():
        for M, C, remap in ((96, 64, 0), (50, 40, 0), (24, 128, 4), (7, 8192, 16), (3000, 64, 0), (700, 256, 0), (64, 1024, 4)):
            x = torch.randn(M, C) * 2 + 0.5
            gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.1
            rm, rv = torch.zeros(C), torch.ones(C)
            sums = K.bn_stats(x)
            stats = K.bn_finalize(sums, M, rm, rv)
            y = K.bn_apply_relu(x, stats, gamma, beta, remap_s=remap, round_out=False)
            xr = x.clone().requires_grad_(True)
            bn = torch.nn.BatchNorm1d(C)
            bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
            ref = F.relu(bn(xr))
            dy = torch.randn(M, C)
            if remap:                                           # output columns are (s, c) instead of (c, s)
                perm = torch.arange(C).view(C // remap, remap).t().reshape(-1)
                ref_out, dy_ref = ref[:, perm], None
            else:
                perm, ref_out = None, ref
            assert torch.allclose(y, ref_out.detach(), atol=2e-5, rtol=1e-5), (M, C, remap)
            assert torch.allclose(rm, bn.running_mean, atol=1e-6) and torch.allclose(rv, bn.running_var, atol=1e-5, rtol=1e-5)
            (ref_out * dy).sum().backward()
            bsums = K.bn_bwd_reduce(dy, y, x, stats, remap_s=remap)
            dx = K.bn_bwd_apply(dy, y, x, stats, gamma, bsums, M, remap_s=remap, round_out=False)
            assert torch.allclose(dx, xr.grad, atol=5e-5, rtol=1e-4), (M, C, remap, float((dx - xr.grad).abs().max()))
        pre = torch.randn(6, 96, 96, 32)                          # 165 888 outputs > 592 CTAs x 256: the grid-stride loop iterates
        bias = torch.randn(3)
        pr = pre.clone().requires_grad_(True)
        br = bias.clone().requires_grad_(True)
        ref = 0.5 * torch.tanh(pr[..., :3].permute(0, 3, 1, 2) + br.view(1, 3, 1, 1)) + 0.5       # sndcgan.py:46-48,52
        out = K.g_final_fwd(pre, bias)
        assert torch.allclose(out, ref.detach(), atol=1e-6)
        dout = torch.randn_like(out)
        (ref * dout).sum().backward()
        dpre, dbias = K.g_final_bwd(dout, out)
        assert torch.allclose(dpre, pr.grad[..., :3].permute(0, 3, 1, 2), atol=1e-6)
        assert torch.allclose(dbias, br.grad, atol=1e-4, rtol=1e-4)
        t = torch.randn(1000)
        assert torch.equal(K.round_tf32_(t), K.round_tf32(t))     # kernel (cvt.rna / its C equivalent) vs the torch bit trick


def test_emulated_fused_adam_matches_torch_adam():
    from contrad_b200 import kernels as K
    torch.manual_seed(0)
    shapes = [(40, 96), (64, 3, 3, 3), (1, 512), (77,), (16, 8, 4, 4)]
    pa = [torch.randn(*s) for s in shapes]
    pb = [torch.nn.Parameter(p.clone()) for p in pa]
    ma, va = [torch.zeros_like(p) for p in pa], [torch.zeros_like(p) for p in pa]
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.5, 0.999))
    with emulated():
        for step in range(1, 4):
            grads = [torch.randn_like(p) * step for p in pa]
            for b, g in zip(pb, grads):
                b.grad = g.clone()
            lr = 2e-4 * step / 3
            for grp in ob.param_groups:
                grp["lr"] = lr
            ob.step()
            K.adam_step(list(zip(pa, grads, ma, va)), lr, 0.5, 0.999, 1e-8, step)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b.detach(), atol=1e-7, rtol=1e-5)


# ------------------------------------------------------------------ StyleGAN2-side SIMT kernels (csrc/sg2_ops.cu)
@pytest.mark.parametrize("nhwc", [True, False])
@pytest.mark.parametrize("H,W,k,up,down,pad", [(9, 7, (1, 3, 3, 1), 1, 1, (2, 2, 2, 2)), (8, 8, (1, 3, 3, 1), 1, 2, (1, 1, 1, 1)),
                                               (5, 6, (1, 3, 3, 1), 2, 1, (2, 1, 2, 1)), (6, 5, (1, 2, 1), 2, 3, (0, 3, 0, 3)),
                                               (8, 8, (1, 3, 3, 1), 1, 1, (-1, 2, -1, 2))])
def test_emulated_upfirdn2d(nhwc, H, W, k, up, down, pad):
    from contrad_b200 import sg2_kernels as S
    from tests import cpu_kernels as CK
    torch.manual_seed(H * 100 + W + up + down)
    C = 5 if not nhwc else 36
    x = torch.randn(2, H, W, C) if nhwc else torch.randn(2, C, H, W)
    fir = torch.tensor(k, dtype=torch.float32)
    fir = fir[None] * fir[:, None]
    fir = fir / fir.sum() * up * up
    fir[0, 1] += 0.01                                       # break the symmetry so that flipping is observable
    with emulated():
        for flip in (False, True):
            got = S.upfirdn2d(x, fir, up, down, pad, nhwc=nhwc, flip=flip, gain=1.5)
            want = CK.upfirdn2d(x, fir, up, down, pad, nhwc=nhwc, flip=flip, gain=1.5)
            assert got.shape == want.shape and _rel(got, want) < 1e-5
        oh = (H * up + pad[2] + pad[3] - fir.shape[0]) // down + 2
        assert _rel(S.upfirdn2d(x, fir, up, down, pad, out_hw=(oh, oh), nhwc=nhwc),
                    CK.upfirdn2d(x, fir, up, down, pad, out_hw=(oh, oh), nhwc=nhwc)) < 1e-5


def test_emulated_sg2_elementwise_and_reduction_kernels():
    from contrad_b200 import sg2_kernels as S
    from tests import cpu_kernels as CK
    torch.manual_seed(0)
    with emulated():
        for B, Ho, C in ((2, 4, 32), (3, 1, 8)):
            x = torch.randn(B, 2 * Ho + 1, 2 * Ho + 1, C)
            assert torch.equal(S.patch_s2_gather(x), CK.patch_s2_gather(x))
            u = torch.randn(B, Ho, Ho, 9, C)
            assert _rel(S.patch_s2_scatter(u), CK.patch_s2_scatter(u)) < 1e-6
        B, H, C = 4, 8, 48
        x, res, bias, g = torch.randn(B, H, H, C), torch.randn(B, H, H, C), torch.randn(C), torch.randn(B, H, H, C)
        assert _rel(S.bias_act(x, bias, 0.2, 1.4, res=res), CK.bias_act(x, bias, 0.2, 1.4, res=res)) < 1e-6
        assert _rel(S.bias_act(x, None, 0.1, 1.0), CK.bias_act(x, None, 0.1, 1.0)) < 1e-6
        assert _rel(S.bias_act_grad(g, x, bias, 0.2, 1.4), CK.bias_act_grad(g, x, bias, 0.2, 1.4)) < 1e-6
        assert _rel(S.bias_act(x, bias, 0.2, 1.4, round_out=True), CK.bias_act(x, bias, 0.2, 1.4)) < 6e-4
        s = torch.randn(B, C)
        const = torch.randn(1, H, H, C)
        assert _rel(S.modulate(x, s), CK.modulate(x, s)) < 1e-6 and _rel(S.modulate(const, s), CK.modulate(const, s)) < 1e-6
        assert _rel(S.mul_reduce(x, res), CK.mul_reduce(x, res)) < 1e-5
        assert _rel(S.mul_reduce(x, const), CK.mul_reduce(x, const.expand(B, -1, -1, -1))) < 1e-5
        noise, nw, d = torch.randn(B, 1, H, H), torch.randn(1), torch.rand(B, C) + 0.5
        assert _rel(S.mod_epilogue(x, d, noise, nw, bias), CK.mod_epilogue(x, d, noise, nw, bias)) < 1e-6
        assert _rel(S.mod_epilogue(x, None, noise, nw, bias), CK.mod_epilogue(x, None, noise, nw, bias)) < 1e-6
        assert _rel(S.noise_grad(g, noise), CK.noise_grad(g, noise)) < 1e-4
        big, big2 = torch.randn(2, 32, 32, 40), torch.randn(2, 32, 32, 40)       # P > 512: the split-P path of mul_reduce
        assert _rel(S.mul_reduce(big, big2), CK.mul_reduce(big, big2)) < 1e-4
        for Bs in (4, 3, 12):                                                   # minibatch stddev incl. second order
            Cs, Hs = 40, 4
            xs = torch.randn(Bs, Hs, Hs, Cs)
            std = CK.stddev_fwd(xs)
            assert _rel(S.stddev_fwd(xs), std) < 1e-5
            dstd, gg = torch.randn_like(std), torch.randn_like(xs)
            assert _rel(S.stddev_bwd(dstd, xs), CK.stddev_bwd(dstd, xs)) < 1e-5
            got, want = S.stddev_bwd_bwd(gg, dstd, xs), CK.stddev_bwd_bwd(gg, dstd, xs)
            assert _rel(got[0], want[0]) < 1e-4 and _rel(got[1], want[1]) < 1e-4
            assert torch.equal(S.stddev_concat(xs, std, 64), CK.stddev_concat(xs, std, 64))
            dy = torch.randn(Bs, Hs, Hs, 64)
            got, want = S.stddev_split(dy, Cs), CK.stddev_split(dy, Cs)
            assert torch.equal(got[0], want[0]) and _rel(got[1], want[1]) < 1e-5
        xi = torch.rand(3, 3, 16, 16)
        assert _rel(S.rgb_to_nhwc(xi, 32, 2.0, -1.0), CK.rgb_to_nhwc(xi, 32, 2.0, -1.0)) < 1e-6
        src, resi = torch.randn(3, 16, 16, 32), torch.randn(3, 3, 16, 16)
        assert _rel(S.nhwc_to_rgb(src, resi, 0.5), CK.nhwc_to_rgb(src, resi, 0.5)) < 1e-6
        assert _rel(S.nhwc_to_rgb(src, None, 2.0), CK.nhwc_to_rgb(src, None, 2.0)) < 1e-6
        z = torch.randn(7, 512)
        assert _rel(S.pixelnorm(z), CK.pixelnorm(z)) < 1e-5
        gr = torch.randn(6, 3, 32, 32)
        assert _rel(S.row_sqsum(gr), CK.row_sqsum(gr)) < 1e-5
        sc = torch.randn(6)
        assert _rel(S.row_scale(gr, sc, 2.0), CK.row_scale(gr, sc, 2.0)) < 1e-6
        a, b = torch.randn(1000), torch.randn(1000)
        assert _rel(S.axpby(a, b, 0.5, -2.0, 0.25), CK.axpby(a, b, 0.5, -2.0, 0.25)) < 1e-6
        assert _rel(S.axpby(a, None, 0.5, 0.0, 0.5), CK.axpby(a, None, 0.5, 0.0, 0.5)) < 1e-6
        dst = [torch.randn(n) for n in (5, 4096, 7001) * 25]                     # 75 tensors: two launches
        src_ = [torch.randn_like(t) for t in dst]
        mine = [t.clone() for t in dst]
        S.ema_lerp(list(zip(mine, src_)), 0.75)
        CK.ema_lerp(list(zip(dst, src_)), 0.75)
        assert max(_rel(a_, b_) for a_, b_ in zip(mine, dst)) < 1e-6


# ------------------------------------------------------------------ the fused SimCLR chain and the first D layer
def test_emulated_fused_augment_matches_reference_golden(golden_dir):
    """csrc/augment.cu, the headline augmentation kernels (persistent CTAs, bulk-copy ring, column mapping) and their
    backward, on the fixtures produced by the unmodified reference chain; the any-size kernels and the uint8 / mixed-
    source launch on the same data."""
    from contrad_b200 import kernels as K
    fx = _load(golden_dir, "augment_simclr.pt")
    with emulated():
        for case in fx["cases"]:
            x, dy, params = case["x"], case["dy"], case["params"]
            y = K.augment_simclr_fwd(x, params, case["order"])
            dx = K.augment_simclr_bwd(x, dy, params, case["order"])
            assert torch.allclose(y, case["y"], atol=2e-5, rtol=0), float((y - case["y"]).abs().max())
            assert torch.allclose(dx, case["dx"], atol=1e-4, rtol=1e-4), float((dx - case["dx"]).abs().max())
            y2, means = K.augment_simclr_large_fwd(x, params, case["order"])
            dx2 = K.augment_simclr_large_bwd(x, dy, params, case["order"], means)
            assert torch.allclose(y2, case["y"], atol=2e-5, rtol=0) and torch.allclose(dx2, case["dx"], atol=1e-4, rtol=1e-4)
        for case in _load(golden_dir, "augment_aux.pt")["uint8"]:
            n = case["x_u8"].shape[0]
            y, _ = K.augment_simclr_mixed_fwd(case["x_u8"], 2 * n, case["fakes"], case["params"], case["order"])
            assert torch.allclose(y, case["y"], atol=2e-5, rtol=0), float((y - case["y"]).abs().max())


def test_emulated_fused_augment_persistent_ring_vs_oracle():
    """More images than the (emulated 4-SM) grid has CTAs: every CTA walks several images through the two-slot prefetch
    ring; per-image jitter order from row 11 (order = -1, the CUDA-graph mode)."""
    from contrad_b200 import kernels as K
    np.random.seed(0); torch.manual_seed(0)
    B, size = 70, 32
    x, dy = torch.rand(B, 3, size, size), torch.randn(B, 3, size, size)
    params, order = O.sample_simclr_params(B, size, size)
    xr = x.clone().requires_grad_(True)
    yr = O.augment_simclr(xr, params, order)
    (yr * dy).sum().backward()
    packed = O.pack_params(params)
    with emulated():
        y = K.augment_simclr_fwd(x, packed, order)
        assert torch.allclose(y, yr.detach(), atol=2e-5, rtol=0)
        dx = K.augment_simclr_bwd(x, dy, packed, order)
        bad = ((dx - xr.grad).abs() > 1e-4 + 1e-4 * xr.grad.abs()).float().mean()
        assert bad < 2e-3, bad                       # a pixel within rounding of the clamp boundary may flip its mask
        with_row = torch.cat([packed, torch.full((1, B), float(order))])
        assert torch.equal(K.augment_simclr_fwd(x, with_row, -1), y)
        x_u8 = (x[:20] * 255).round().to(torch.uint8)
        y_mixed, _ = K.augment_simclr_mixed_fwd(x_u8, 40, x[40:], packed, order)
        y_cat = K.augment_simclr_fwd(torch.cat([O.to_tensor_u8(x_u8)] * 2 + [x[40:]]), packed, order)
        assert torch.equal(y_mixed, y_cat)           # bit-equal, as on the B200


def test_emulated_fused_augment_second_build_is_bit_identical(golden_dir, monkeypatch):
    """CB200_AUGMENT_V=2 (parameters staged through a 4-slot shared-memory ring by twelve loader lanes, byte-offset tap
    tables): same pixel arithmetic, so the outputs must be BIT-equal to the first build - on the reference fixtures, with
    CTAs that walk up to six images (slot wrap-around, both ring buffers, the tail without a successor), per-image
    jitter order from row 11, and at 64 x 64."""
    from contrad_b200 import kernels as K
    np.random.seed(1); torch.manual_seed(1)

    def both(x, packed, order):
        monkeypatch.setenv("CB200_AUGMENT_V", "1")
        y1 = K.augment_simclr_fwd(x, packed, order)
        monkeypatch.setenv("CB200_AUGMENT_V", "2")
        y2 = K.augment_simclr_fwd(x, packed, order)
        monkeypatch.delenv("CB200_AUGMENT_V")
        return y1, y2

    with emulated():
        for case in _load(golden_dir, "augment_simclr.pt")["cases"]:
            y1, y2 = both(case["x"], case["params"], case["order"])
            assert torch.equal(y1, y2)
            assert torch.allclose(y2, case["y"], atol=2e-5, rtol=0)
        for B, size in ((1, 32), (25, 32), (131, 32), (14, 64)):
            x = torch.rand(B, 3, size, size)
            params, order = O.sample_simclr_params(B, size, size)
            packed = O.pack_params(params)
            y1, y2 = both(x, packed, order)
            assert torch.equal(y1, y2), (B, size, float((y1 - y2).abs().max()))
            per_image = torch.cat([packed, (torch.rand(1, B) < 0.5).float()])
            y1, y2 = both(x, per_image, -1)
            assert torch.equal(y1, y2), (B, size, "row 11")


def test_emulated_fused_augment_on_degenerate_images(monkeypatch):
    """tests/edge_inputs.py (colour-cube corners = breakpoints of the colour wheel, saturation-0 images whose hue is
    undefined, a 0/1 checkerboard, stripes): both builds of the fused forward kernel, the backward, the any-size path and
    the uint8 / mixed-source launch against the oracle, which tests/test_oracle_golden.py pins on the unmodified reference
    chain for the very same images."""
    from contrad_b200 import kernels as K
    from tests.edge_inputs import degenerate_images
    x = degenerate_images()
    B, size = x.shape[0], x.shape[-1]
    with emulated():
        for seed in (1, 2):
            np.random.seed(seed); torch.manual_seed(seed)
            params, order = O.sample_simclr_params(B, size, size)
            packed = O.pack_params(params)
            xr = x.clone().requires_grad_(True)
            yr = O.augment_simclr(xr, params, order)
            dy = torch.randn(yr.shape)
            (yr * dy).sum().backward()
            for build in ("1", "2"):
                monkeypatch.setenv("CB200_AUGMENT_V", build)
                y = K.augment_simclr_fwd(x, packed, order)
                assert torch.allclose(y, yr.detach(), atol=2e-5, rtol=0), (build, float((y - yr).abs().max()))
            monkeypatch.delenv("CB200_AUGMENT_V")
            dx = K.augment_simclr_bwd(x, dy, packed, order)
            bad = ((dx - xr.grad).abs() > 1e-4 + 1e-4 * xr.grad.abs()).float().mean()
            assert bad < 5e-3, float(bad)                # pixels exactly ON the clamp boundary (0 / 1 images) may flip their mask
            y2, _ = K.augment_simclr_large_fwd(x, packed, order)
            assert torch.allclose(y2, yr.detach(), atol=2e-5, rtol=0)
            exact = (x == 0) | (x == 1)                  # images that ToTensor can produce from bytes
            keep = exact.flatten(1).all(1).nonzero().flatten()
            x_u8 = (x[keep] * 255).to(torch.uint8)
            y_u8, _ = K.augment_simclr_mixed_fwd(x_u8, len(keep), x[:0], packed[:, keep].contiguous(), order)
            assert torch.allclose(y_u8, yr.detach()[keep], atol=2e-5, rtol=0)


@pytest.mark.parametrize("kind", ["nonsat", "hinge", "wgan", "lsgan"])
def test_emulated_gan_losses_at_extreme_logits(kind):
    """L_dis / L_gen (training/gan/contrad.py:52-64,73-81) where torch's softplus switches branches (|x| = 20), where
    exp() overflows in fp32 (|x| > 88.7) and far beyond: values and logit gradients against torch, no inf / NaN."""
    from contrad_b200 import kernels as K
    v = torch.tensor([-1e4, -100., -88.8, -20.0001, -19.9999, -1., -1e-8, 0., 1e-8, 1., 19.9999, 20.0001, 88.8, 100., 1e4])
    scale = float(v.abs().sum()) if kind != "lsgan" else float((v ** 2).sum())
    with emulated():
        dr, dg = v.clone().requires_grad_(True), v.flip(0).clone().requires_grad_(True)
        ref = {"nonsat": lambda: F.softplus(dg).mean() + F.softplus(-dr).mean(),
               "wgan": lambda: dg.mean() - dr.mean(),
               "hinge": lambda: F.relu(1. + dg).mean() + F.relu(1. - dr).mean(),
               "lsgan": lambda: 0.5 * (((dr - 1.0) ** 2).mean() + (dg ** 2).mean())}[kind]()
        ref.backward()
        out, g_r, g_g = K.gan_d_loss(v.clone(), v.flip(0).clone(), kind)
        assert torch.isfinite(out).all() and abs(float(out[0]) - float(ref.detach())) < 1e-6 * scale
        assert torch.allclose(g_r, dr.grad, atol=1e-7, rtol=1e-6) and torch.allclose(g_g, dg.grad, atol=1e-7, rtol=1e-6)
        d = v.clone().requires_grad_(True)
        rg = {"nonsat": lambda: F.softplus(-d).mean(), "lsgan": lambda: 0.5 * ((d - 1.0) ** 2).mean()}.get(kind, lambda: -d.mean())()
        rg.backward()
        out, g = K.gan_g_loss(v.clone(), kind)
        assert torch.isfinite(out).all() and abs(float(out[0]) - float(rg.detach())) < 1e-6 * scale
        assert torch.allclose(g, d.grad, atol=1e-7, rtol=1e-6)


def test_emulated_empty_and_malformed_inputs_are_noops_or_loud_errors():
    """Edge cases at the C ABI: an empty batch is a no-op for the augmentation entry points (torch returns empty tensors
    for them too) and a CB200Error with a message for the convolution / reduction / optimiser entry points (the
    reference's step has no meaning on an empty batch: its losses are means over zero rows); malformed parameter blocks
    and unsupported shapes are refused before any launch.  Never a crash, never a silent wrong answer."""
    from contrad_b200 import kernels as K
    from contrad_b200._capi import CB200Error
    with emulated():
        p0 = torch.zeros(11, 0)
        e = torch.zeros(0, 3, 32, 32)
        assert K.augment_simclr_fwd(e, p0, 0).shape == (0, 3, 32, 32)
        assert K.augment_simclr_bwd(e, e, p0, 0).shape == (0, 3, 32, 32)
        y, means = K.augment_simclr_large_fwd(torch.zeros(0, 3, 40, 40), p0, 0)
        assert y.shape == (0, 3, 40, 40) and means.shape == (0, 3)
        for bad in (lambda: K.conv_first_fwd(e, torch.randn(64, 3, 3, 3), None, torch.zeros(64)),
                    lambda: K.conv_first_wgrad(e, torch.zeros(0, 32, 32, 64)),
                    lambda: K.colsum(torch.zeros(0, 64)),
                    lambda: K.adam_step([], 1e-3, 0.5, 0.999, 1e-8, 1),
                    lambda: K.split_tf32(torch.zeros(0, 32), 0)):
            with pytest.raises(CB200Error):
                bad()
        x = torch.rand(2, 3, 32, 32)
        with pytest.raises(AssertionError):                 # parameter block with the wrong number of rows / samples
            K.augment_simclr_fwd(x, torch.zeros(11, 3), 0)
        with pytest.raises(AssertionError):
            K.augment_simclr_fwd(x, torch.zeros(11, 2), -1)  # per-image order needs row 11
        with pytest.raises(CB200Error):                     # the fused small-image kernels stop at 64 x 64
            K.augment_simclr_fwd(torch.rand(1, 3, 128, 128), torch.zeros(11, 1), 0)
        with pytest.raises(CB200Error):                     # width must be a multiple of 4 (float4 rows)
            K.augment_simclr_fwd(torch.rand(1, 3, 30, 30), torch.zeros(11, 1), 0)


@pytest.mark.parametrize("B,H", [(3, 32), (2, 16)])
def test_emulated_conv_first_layer(B, H):
    """csrc/conv_first.cu: Conv2d(3 -> 64) with the x*2-1 input affine, its weight / bias gradient (persistent CTAs,
    bulk-copied dY tiles) and the channel extraction of the data gradient."""
    from contrad_b200 import kernels as K
    torch.manual_seed(B)
    x = torch.rand(B, 3, H, H)
    w, bias = torch.randn(64, 3, 3, 3) * 0.1, torch.randn(64) * 0.1
    sigma = torch.tensor([2.0, 0.5])
    with emulated():
        y = K.conv_first_fwd(x, w, sigma, bias, slope=0.1, round_out=False)
        ref = F.leaky_relu(F.conv2d(x.double() * 2 - 1, w.double() * 0.5, bias.double(), padding=1), 0.1)
        assert torch.allclose(y, ref.permute(0, 2, 3, 1).float(), atol=1e-5, rtol=1e-5)
        dy = torch.randn(B, H, H, 64)
        dw, db = K.conv_first_wgrad(x, dy)
        ref_dw = torch.nn.grad.conv2d_weight(x.double() * 2 - 1, (64, 3, 3, 3), dy.permute(0, 3, 1, 2).double(), padding=1)
        assert torch.allclose(dw.view(64, 3, 3, 3), ref_dw.float(), atol=1e-3, rtol=1e-4)
        assert torch.allclose(db, dy.sum(dim=(0, 1, 2)), atol=1e-3, rtol=1e-4)
        dpad = torch.randn(B, H, H, 32)
        dx = K.conv_first_dgrad_finish(dpad)
        assert torch.allclose(dx, 2 * dpad[..., :3].permute(0, 3, 1, 2), atol=1e-6)


# ------------------------------------------------------------------ the product's train step on the CPU
@pytest.mark.parametrize("strict", [False, True, "full"])
def test_product_train_step_on_cpu_vs_oracle(strict):
    """BASELINE configs[0] in spirit ("SNDCGAN+ContraD on CPU, one step, synthetic 32x32: plumbing, no GPU"): the PRODUCT's
    own modules, autograd Functions, engine.train_step and every SIMT kernel (emulated) run one complete D+G step incl.
    spectral norm and the optimiser on CPU tensors; only the tcgen05 entry points are torch stand-ins
    (tests/cpu_tc_standins.py).  Scalars, gradient norms and updated buffers against the fp32 oracle on identical weights,
    latents and augmentation draws.  Heads / generator at reduced width to keep the emulation short.

    strict = True runs the generator step in the strict precision mode (contrad_b200/precision.py: error-compensated
    "3xTF32" operands through the same GEMM entry points): the generator's gradient norm - which single-pass TF32 only
    holds to ~1e-2 at initialisation (tools/tf32_sensitivity.py) - must then meet north_star's 1e-3."""
    from types import SimpleNamespace
    import tests.cpu_tc_standins as TC
    from contrad_b200 import engine, precision
    from contrad_b200.functional import AugmentSimCLRFn
    from contrad_b200.models.gan.sndcgan import D_SNDCGAN, G_SNDCGAN
    from contrad_b200.training.gan import contrad
    n, ngf, nz, d_hidden = 6, 64, 16, 16          # ngf = 64: the last generator layer reuses the 64-channel first-layer kernels
    gen_w = torch.Generator().manual_seed(5)
    sd_d = O.make_d_state(d_hidden=d_hidden, generator=gen_w)
    sd_g = O.make_g_state(ngf=ngf, nz=nz, generator=gen_w)
    np.random.seed(11); torch.manual_seed(11)
    images = torch.rand(n, 3, 32, 32)
    z_d = O.sample_latent(n, nz); aug_d = O.sample_simclr_params(3 * n, 32, 32)
    z_g = O.sample_latent(n, nz); aug_g = O.sample_simclr_params(n, 32, 32)

    # ---- oracle
    sd_d_o = {k: v.clone() for k, v in sd_d.items()}
    sd_g_o = {k: v.clone() for k, v in sd_g.items()}
    opt_g_o, opt_d_o = O.Adam(O.trainable(sd_g_o).values(), 2e-4), O.Adam(O.trainable(sd_d_o).values(), 2e-4)
    orig_g = O.g_sndcgan_forward
    O.g_sndcgan_forward = lambda sd, z, **kw: orig_g(sd, z, ngf=ngf)
    try:
        ref = O.train_step(sd_g_o, sd_d_o, opt_g_o, opt_d_o, images, z_d, z_g, aug_d, aug_g, step=1)
    finally:
        O.g_sndcgan_forward = orig_g

    # ---- product on CPU tensors
    class Aug(torch.nn.Module):
        def __init__(self, blocks):
            super().__init__(); self.blocks = list(blocks)
        def forward(self, x):
            packed, order = self.blocks.pop(0)
            return AugmentSimCLRFn.apply(x, packed, order)

    class Gw(torch.nn.Module):
        def __init__(self, g, zs):
            super().__init__(); self.g, self.zs = g, list(zs)
        def sample_latent(self, k):
            return self.zs.pop(0)
        def forward(self, z):
            return self.g(z)
        def parameters(self, recurse=True):
            return self.g.parameters(recurse)
        def train(self, mode=True):
            self.g.train(mode); return self

    with emulated(), TC.patched(), precision.strict(strict):
        D = D_SNDCGAN((32, 32, 3), mlp_linear=True, d_hidden=d_hidden)
        G = G_SNDCGAN((32, 32, 3), ngf=ngf, nz=nz)
        D.load_state_dict(sd_d); G.load_state_dict(sd_g)
        P = SimpleNamespace(augment_fn=Aug([(O.pack_params(aug_d[0]), aug_d[1]), (O.pack_params(aug_g[0]), aug_g[1])]),
                            temp=0.1, lbd_a=1.0, distributed=False)
        opts = {"loss": "nonsat", "warmup": 3000, "lr": 2e-4}
        opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
        opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
        got = engine.train_step(P, opts, {"D": contrad.loss_D_fn, "G": contrad.loss_G_fn}, (Gw(G, [z_d, z_g]), D),
                                (opt_G, opt_D), images, 1, record_grad_norms=True)
    got = {k: float(v) for k, v in got.items()}
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
    # TF32 rounding of the GEMM operands is part of the product's arithmetic (done by the producing kernels): 1e-3 bars
    assert rel(got["d_loss"], ref["l_con_pos"] + ref["l_con_neg"]) < 1e-3, (got, ref)
    assert rel(got["d_penalty"], ref["l_dis"]) < 1e-3 and rel(got["g_loss"], ref["l_gen"]) < 1e-3, (got, ref)
    assert abs(got["d_real"] - ref["d_real"]) < 1e-3 and abs(got["d_gen"] - ref["d_gen"]) < 1e-3
    assert rel(got["d_grad_norm"], ref["d_grad_norm"]) < (1e-3 if strict == "full" else 5e-3), (got["d_grad_norm"], ref["d_grad_norm"])
    assert rel(got["g_grad_norm"], ref["g_grad_norm"]) < (1e-3 if strict else 3e-2), (got["g_grad_norm"], ref["g_grad_norm"])
    sd_now = D.state_dict()
    for k, v in sd_d_o.items():
        if k.endswith(("weight_u", "weight_v")):                      # two power iterations (D step + G step)
            assert torch.allclose(sd_now[k], v, atol=2e-4), k
    g_now = G.state_dict()
    for k in ("norm_init.running_mean", "main.1.running_var", "main.7.running_mean"):
        assert torch.allclose(g_now[k], sd_g_o[k], atol=1e-4, rtol=1e-3), k


# ------------------------------------------------------------------ the data-parallel path on two CPU ranks (gloo)
def _dp_worker(rank, world, port, payload, results):
    import torch.distributed as dist
    from types import SimpleNamespace
    import tests.cpu_tc_standins as TC
    from contrad_b200.functional import AugmentSimCLRFn
    from contrad_b200.models.gan.sndcgan import D_SNDCGAN, G_SNDCGAN
    from contrad_b200.training.gan import contrad
    from contrad_b200 import engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = payload["images"].shape[0] // world
        rows = slice(rank * n, (rank + 1) * n)
        with emulated(), TC.patched():
            D = D_SNDCGAN((32, 32, 3), mlp_linear=True, d_hidden=payload["d_hidden"])
            G = torch.nn.SyncBatchNorm.convert_sync_batchnorm(G_SNDCGAN((32, 32, 3), ngf=64, nz=payload["nz"]))
            D.load_state_dict(payload["sd_d"]); G.load_state_dict(payload["sd_g"])
            D.train(); G.train()
            engine.set_grad(G, False); engine.set_grad(D, True)
            with torch.no_grad():
                gen = G(payload["z"][rows])                          # SyncBN: batch statistics over BOTH ranks
            # the rank's columns of the full-batch draws: views [x | x | G(z)] of its own samples
            N = payload["images"].shape[0]
            cols = torch.cat([torch.arange(N)[rows] + k * N for k in range(3)])
            packed = payload["aug"][:, cols].contiguous()

            class Aug(torch.nn.Module):
                def forward(self, x):
                    return AugmentSimCLRFn.apply(x, packed, payload["order"])

            P = SimpleNamespace(augment_fn=Aug(), temp=0.1, lbd_a=1.0, distributed=True)
            d_loss, aux = contrad.loss_D_fn(P, D, {"loss": "nonsat"}, payload["images"][rows], gen)
            (d_loss + aux["penalty"]).backward()
            engine.allreduce_gradients(D)                            # train_gan.py:311-313 (DDP averages)
            gn = float(engine.grad_norm(D))
        results[rank] = {"gen": gen.clone(), "l_con": float(d_loss), "l_dis": float(aux["penalty"]), "grad_norm": gn,
                         "g_main0": D.main[0].weight_orig.grad.clone()}
    finally:
        dist.destroy_process_group()


def test_data_parallel_d_step_on_two_cpu_ranks_vs_single_process_oracle():
    """SURVEY 8e on the CPU: two gloo ranks run the PRODUCT's distributed D step (SyncBatchNorm in G, one packed all-gather
    of the embeddings, replicated full-batch contrastive losses, gradient averaging) on their halves of a batch; the
    generated images, the contrastive loss and the averaged gradient equal the single-process oracle on the whole batch
    (the reference's semantics: identical full-batch L_con on every rank, L_dis a local mean, DDP's 1/W averaging)."""
    import torch.multiprocessing as mp
    world, n_all, nz, d_hidden = 2, 6, 16, 16
    gen_w = torch.Generator().manual_seed(9)
    sd_d = O.make_d_state(d_hidden=d_hidden, generator=gen_w)
    sd_g = O.make_g_state(ngf=64, nz=nz, generator=gen_w)
    np.random.seed(3); torch.manual_seed(3)
    images = torch.rand(n_all, 3, 32, 32)
    z = O.sample_latent(n_all, nz)
    params, order = O.sample_simclr_params(3 * n_all, 32, 32)
    payload = {"sd_d": sd_d, "sd_g": sd_g, "images": images, "z": z, "aug": O.pack_params(params), "order": order,
               "nz": nz, "d_hidden": d_hidden}
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_dp_worker, args=(world, 29731, payload, results), nprocs=world, join=True)
    assert set(results.keys()) == {0, 1}

    # single-process oracle on the full batch
    sd_d_o = {k: v.clone() for k, v in sd_d.items()}
    O.set_requires_grad(sd_d_o, True)
    with torch.no_grad():
        gen_o = O.g_sndcgan_forward({k: v.clone() for k, v in sd_g.items()}, z)
    l_con_o, l_dis_o, ex = O.loss_d(sd_d_o, images, gen_o, params, order)
    n = n_all // world
    for r in range(world):
        res = results[r]
        assert torch.allclose(res["gen"], gen_o[r * n:(r + 1) * n], atol=2e-3), r         # 4 TF32 layers + tanh
        assert abs(res["l_con"] - float(l_con_o)) < 1e-3 * abs(float(l_con_o)), (res["l_con"], float(l_con_o))
    # L_dis is a mean over the rank's own samples; the two local means average to the full-batch value
    assert abs(0.5 * (results[0]["l_dis"] + results[1]["l_dis"]) - float(l_dis_o)) < 1e-3 * abs(float(l_dis_o))
    # DDP semantics: every rank back-propagates the full-batch L_con through ITS rows, then gradients are
    # averaged: backbone gradient = (1/W) * d L_con / d theta  +  d L_dis(full batch) / d theta
    (l_con_o / world + l_dis_o).backward()
    assert torch.equal(results[0]["g_main0"], results[1]["g_main0"])
    ref_g = sd_d_o["main.0.weight_orig"].grad
    err = float((results[0]["g_main0"] - ref_g).norm() / ref_g.norm())
    assert err < 0.12, err          # per-tensor bar of tests/test_gpu_model.py (LeakyReLU kinks within TF32 rounding at a tiny batch)
    cos = float((results[0]["g_main0"] * ref_g).sum() / (results[0]["g_main0"].norm() * ref_g.norm()))
    assert cos > 0.995, cos
    tot_o = O.grad_norm(sd_d_o)
    assert abs(results[0]["grad_norm"] - tot_o) < 5e-3 * tot_o, (results[0]["grad_norm"], tot_o)


def test_emulated_conv_first_wgrad_second_mapping():
    """The opt-in weight-gradient mapping of csrc/conv_first.cu (CB200_CONV_FIRST_WGRAD=2: one input channel's nine taps per
    thread, sliding register window) against torch - in a subprocess, because the variant is read from the environment
    once per process.  Several tiles per CTA (B = 40 on the emulated 4-SM device) exercise the persistent loop."""
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from tests.emu import emulated
from contrad_b200 import kernels as K
with emulated():
    for B, H in ((40, 32), (3, 16), (2, 20)):
        torch.manual_seed(B)
        x, dy = torch.rand(B, 3, H, H), torch.randn(B, H, H, 64)
        dw, db = K.conv_first_wgrad(x, dy)
        ref = torch.nn.grad.conv2d_weight(x.double() * 2 - 1, (64, 3, 3, 3), dy.permute(0, 3, 1, 2).double(), padding=1)
        assert torch.allclose(dw.view(64, 3, 3, 3), ref.float(), atol=2e-3, rtol=1e-4), (B, H, float((dw.view(64, 3, 3, 3) - ref.float()).abs().max()))
        assert torch.allclose(db, dy.sum(dim=(0, 1, 2)), atol=2e-3, rtol=1e-4), (B, H)
print("wgrad v2 ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CB200_CONV_FIRST_WGRAD="2")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert r.returncode == 0 and b"wgrad v2 ok" in r.stdout, r.stdout.decode()[-3000:]


def test_tensor_core_contrastive_formulation_matches_reference_fixtures(golden_dir):
    """functional.ContrastiveTCFn (north_star: similarity matrix as a tcgen05 GEMM, softmax / CE as warp-shuffle row
    reductions): the row kernels run emulated, the GEMMs on the torch stand-ins, against the reference's own nt_xent /
    supcon_fake values and gradients (tests/golden/contrastive.pt) at the fixtures' 2e-5."""
    import tests.cpu_tc_standins as TC
    from contrad_b200.functional import ContrastiveTCFn
    fx = _load(golden_dir, "contrastive.pt")
    with emulated(), TC.patched():
        for case in fx["cases"]:
            n = case["n"]
            a, b, c = (case[k].clone().requires_grad_(True) for k in ("a", "b", "c"))
            l1 = ContrastiveTCFn.apply(torch.cat([a, b]), n, 0, 0.1)
            g1 = torch.autograd.grad(l1, [a, b])
            assert abs(float(l1) - case["nt_xent"]) < 2e-5 * abs(case["nt_xent"]), (n, float(l1), case["nt_xent"])
            for g, w in zip(g1, case["nt_xent_grads"]):
                assert torch.allclose(g, w, atol=2e-5 * float(w.abs().max()) + 1e-7, rtol=1e-4), n
            l2 = ContrastiveTCFn.apply(torch.cat([a, b, c]), n, 1, 0.1)
            g2 = torch.autograd.grad(l2, [a, b, c])
            assert abs(float(l2) - case["supcon"]) < 2e-5 * abs(case["supcon"]), (n, float(l2), case["supcon"])
            for g, w in zip(g2, case["supcon_grads"]):
                assert torch.allclose(g, w, atol=2e-5 * float(w.abs().max()) + 1e-7, rtol=1e-4), n
            l3 = ContrastiveTCFn.apply(torch.cat([a, b]), n, 0, 0.5)
            assert abs(float(l3) - case["nt_xent_t05"]) < 2e-5 * abs(case["nt_xent_t05"])


def test_generator_latent_gradient_vs_oracle():
    """GSNDCGANFn.backward returns d loss / d z when the latent asks for it (latent optimisation); checked in
    the full strict precision mode, where the batch-of-4 BatchNorm statistics do not amplify TF32 rounding."""
    import tests.cpu_tc_standins as TC
    from contrad_b200 import precision
    from contrad_b200.models.gan.sndcgan import G_SNDCGAN
    sd_g = O.make_g_state(ngf=64, nz=16, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(1)
    z0, c = torch.empty(4, 16).uniform_(-1, 1), torch.randn(4, 3, 32, 32)
    with emulated(), TC.patched(), precision.strict("full"):
        G = G_SNDCGAN((32, 32, 3), ngf=64, nz=16)
        G.load_state_dict(sd_g); G.train()
        z = z0.clone().requires_grad_(True)
        (G(z) * c).sum().backward()
    zo = z0.clone().requires_grad_(True)
    (O.g_sndcgan_forward({k: v.clone() for k, v in sd_g.items()}, zo, ngf=64) * c).sum().backward()
    assert float((z.grad - zo.grad).norm() / zo.grad.norm()) < 2e-3
