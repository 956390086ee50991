"""GPU parity tests of the StyleGAN2 side of the path, through the C ABI.

  * every kernel of csrc/sg2_ops.cu against its torch stand-in (tests/cpu_kernels.py; fp32, 1e-5);
  * the tensor-core backed autograd families (MmNT/MmNN/MmTN, Conv3x3/Dgrad/Wgrad) incl. their double backward on
    TF32-exact inputs (1e-4);
  * ResidualDiscriminatorP, the R1 double backward, Generator (style mixing, explicit noise) and the D-step losses of
    train_stylegan2_contraD.py against the fixtures produced by the unmodified reference (stylegan2_small.pt).
    Tolerances here are TF32 tolerances (fp32 storage, TF32 multiply, fp32 accumulate - the reference's own GPU
    arithmetic): 1e-2 of the tensor's max for activations / gradients, 5e-3 on loss scalars; the measured errors are
    written to gpurun_out/sg2_parity.json.
  * one full train_step_stylegan2 iteration (G step, D step with R1, EMA) runs and updates every parameter."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import cpu_kernels as CK

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPORT = {}


@pytest.fixture(scope="module")
def S():
    from contrad_b200 import sg2_kernels
    return sg2_kernels


@pytest.fixture(scope="module")
def K():
    from contrad_b200 import kernels
    return kernels


def _rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _chk(name, a, b, tol):
    err = _rel(a, b)
    REPORT[name] = err
    assert err < tol, "%s: rel err %.3e (tol %.1e)" % (name, err, tol)


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("nhwc", [True, False])
@pytest.mark.parametrize("H,W,k,up,down,pad", [(9, 7, (1, 3, 3, 1), 1, 1, (2, 2, 2, 2)), (8, 8, (1, 3, 3, 1), 1, 2, (1, 1, 1, 1)),
                                               (5, 6, (1, 3, 3, 1), 2, 1, (2, 1, 2, 1)), (6, 5, (1, 2, 1), 2, 3, (0, 3, 0, 3)),
                                               (8, 8, (1, 3, 3, 1), 1, 1, (-1, 2, -1, 2)), (33, 33, (1, 3, 3, 1), 1, 1, (1, 1, 1, 1)),
                                               (16, 16, (1, 3, 3, 1), 2, 1, (2, 1, 2, 1))])
def test_upfirdn2d(S, nhwc, H, W, k, up, down, pad):
    torch.manual_seed(H * 100 + W + up + down)
    C = 5 if not nhwc else 36
    x = torch.randn(3, H, W, C) if nhwc else torch.randn(3, C, H, W)
    fir = torch.tensor(k, dtype=torch.float32)
    fir = fir[None] * fir[:, None]
    fir = fir / fir.sum() * up * up
    fir[0, 1] += 0.01                                       # break the symmetry so that flipping is observable
    for flip in (False, True):
        got = S.upfirdn2d(x.cuda(), fir.cuda(), up, down, pad, nhwc=nhwc, flip=flip, gain=1.5)
        want = CK.upfirdn2d(x, fir, up, down, pad, nhwc=nhwc, flip=flip, gain=1.5)
        assert got.shape == want.shape
        assert _rel(got, want) < 1e-5
    # explicit output size (the backward pass asks for the forward input size)
    oh = (H * up + pad[2] + pad[3] - fir.shape[0]) // down + 2
    got = S.upfirdn2d(x.cuda(), fir.cuda(), up, down, pad, out_hw=(oh, oh), nhwc=nhwc)
    want = CK.upfirdn2d(x, fir, up, down, pad, out_hw=(oh, oh), nhwc=nhwc)
    assert _rel(got, want) < 1e-5


@pytest.mark.parametrize("B,Ho,C", [(2, 4, 32), (3, 16, 128), (5, 1, 8)])
def test_patch_s2(S, B, Ho, C):
    torch.manual_seed(B + Ho)
    x = torch.randn(B, 2 * Ho + 1, 2 * Ho + 1, C)
    assert torch.equal(S.patch_s2_gather(x.cuda()).cpu(), CK.patch_s2_gather(x))
    u = torch.randn(B, Ho, Ho, 9, C)
    assert _rel(S.patch_s2_scatter(u.cuda()), CK.patch_s2_scatter(u)) < 1e-6
    # gather / scatter are transposes of each other
    lhs = (S.patch_s2_gather(x.cuda()) * u.cuda()).sum()
    rhs = (x.cuda() * S.patch_s2_scatter(u.cuda())).sum()
    assert abs(float(lhs - rhs)) < 1e-3 * abs(float(lhs)) + 1e-3


def test_bias_act_modulate_epilogue(S):
    torch.manual_seed(0)
    B, H, C = 6, 8, 48
    x, res, bias = torch.randn(B, H, H, C), torch.randn(B, H, H, C), torch.randn(C)
    assert _rel(S.bias_act(x.cuda(), bias.cuda(), 0.2, 1.4, res=res.cuda()), CK.bias_act(x, bias, 0.2, 1.4, res=res)) < 1e-6
    assert _rel(S.bias_act(x.cuda(), None, 0.1, 1.0), CK.bias_act(x, None, 0.1, 1.0)) < 1e-6
    g = torch.randn(B, H, H, C)
    assert _rel(S.bias_act_grad(g.cuda(), x.cuda(), bias.cuda(), 0.2, 1.4), CK.bias_act_grad(g, x, bias, 0.2, 1.4)) < 1e-6
    # TF32 rounding of the outputs is a <= 2^-11 relative perturbation
    assert _rel(S.bias_act(x.cuda(), bias.cuda(), 0.2, 1.4, round_out=True), CK.bias_act(x, bias, 0.2, 1.4)) < 6e-4
    s = torch.randn(B, C)
    assert _rel(S.modulate(x.cuda(), s.cuda()), CK.modulate(x, s)) < 1e-6
    const = torch.randn(1, H, H, C)
    assert _rel(S.modulate(const.cuda(), s.cuda()), CK.modulate(const, s)) < 1e-6
    assert _rel(S.mul_reduce(x.cuda(), res.cuda()), CK.mul_reduce(x, res)) < 1e-5
    assert _rel(S.mul_reduce(x.cuda(), const.cuda()), CK.mul_reduce(x, const.expand(B, -1, -1, -1))) < 1e-5
    noise, nw, d = torch.randn(B, 1, H, H), torch.randn(1), torch.rand(B, C) + 0.5
    assert _rel(S.mod_epilogue(x.cuda(), d.cuda(), noise.cuda(), nw.cuda(), bias.cuda()),
                CK.mod_epilogue(x, d, noise, nw, bias)) < 1e-6
    assert _rel(S.mod_epilogue(x.cuda(), None, noise.cuda(), nw.cuda(), bias.cuda()),
                CK.mod_epilogue(x, None, noise, nw, bias)) < 1e-6
    assert _rel(S.noise_grad(g.cuda(), noise.cuda()), CK.noise_grad(g, noise)) < 1e-4
    big = torch.randn(2, 64, 64, 40)                                   # P > 512: the split-P path of mul_reduce
    big2 = torch.randn(2, 64, 64, 40)
    assert _rel(S.mul_reduce(big.cuda(), big2.cuda()), CK.mul_reduce(big, big2)) < 1e-4


@pytest.mark.parametrize("B", [4, 8, 3, 12])
def test_minibatch_stddev(S, B):
    torch.manual_seed(B)
    C, H = 40, 4
    x = torch.randn(B, H, H, C)
    std = CK.stddev_fwd(x)
    assert _rel(S.stddev_fwd(x.cuda()), std) < 1e-5
    dstd = torch.randn_like(std)
    assert _rel(S.stddev_bwd(dstd.cuda(), x.cuda()), CK.stddev_bwd(dstd, x)) < 1e-5
    gg = torch.randn_like(x)
    got = S.stddev_bwd_bwd(gg.cuda(), dstd.cuda(), x.cuda())
    want = CK.stddev_bwd_bwd(gg, dstd, x)
    assert _rel(got[0], want[0]) < 1e-4 and _rel(got[1], want[1]) < 1e-4
    assert torch.equal(S.stddev_concat(x.cuda(), std.cuda(), 64).cpu(), CK.stddev_concat(x, std, 64))
    dy = torch.randn(B, H, H, 64)
    got = S.stddev_split(dy.cuda(), C)
    want = CK.stddev_split(dy, C)
    assert torch.equal(got[0].cpu(), want[0]) and _rel(got[1], want[1]) < 1e-5


def test_layout_and_misc_kernels(S):
    torch.manual_seed(5)
    x = torch.rand(5, 3, 16, 16)
    assert _rel(S.rgb_to_nhwc(x.cuda(), 32, 2.0, -1.0), CK.rgb_to_nhwc(x, 32, 2.0, -1.0)) < 1e-6
    src, res = torch.randn(5, 16, 16, 32), torch.randn(5, 3, 16, 16)
    assert _rel(S.nhwc_to_rgb(src.cuda(), res.cuda(), 0.5), CK.nhwc_to_rgb(src, res, 0.5)) < 1e-6
    assert _rel(S.nhwc_to_rgb(src.cuda(), None, 2.0), CK.nhwc_to_rgb(src, None, 2.0)) < 1e-6
    z = torch.randn(7, 512)
    assert _rel(S.pixelnorm(z.cuda()), CK.pixelnorm(z)) < 1e-5
    g = torch.randn(6, 3, 32, 32)
    assert _rel(S.row_sqsum(g.cuda()), CK.row_sqsum(g)) < 1e-5
    s = torch.randn(6)
    assert _rel(S.row_scale(g.cuda(), s.cuda(), 2.0), CK.row_scale(g, s, 2.0)) < 1e-6
    a, b = torch.randn(1000), torch.randn(1000)
    assert _rel(S.axpby(a.cuda(), b.cuda(), 0.5, -2.0, 0.25), CK.axpby(a, b, 0.5, -2.0, 0.25)) < 1e-6
    assert _rel(S.axpby(a.cuda(), None, 0.5, 0.0, 0.5), CK.axpby(a, None, 0.5, 0.0, 0.5)) < 1e-6
    dst = [torch.randn(n) for n in (5, 4096, 70001) * 30]                  # 90 tensors: two launches
    src_ = [torch.randn_like(t) for t in dst]
    d_gpu = [t.cuda() for t in dst]
    S.ema_lerp([(d, s_.cuda()) for d, s_ in zip(d_gpu, src_)], 0.75)
    CK.ema_lerp(list(zip(dst, src_)), 0.75)
    assert max(_rel(a_, b_) for a_, b_ in zip(d_gpu, dst)) < 1e-6


# ------------------------------------------------------------------------------------------------ tensor-core families
def _tf32(t, K):
    return K.round_tf32(t)


def test_mm_family_double_backward(K):
    from contrad_b200 import sg2_functional as SF
    torch.manual_seed(1)
    for (M, N, Kd) in ((70, 128, 64), (256, 32, 128), (33, 4608, 64), (512, 64, 32)):
        a = _tf32(torch.randn(M, Kd), K).cuda().requires_grad_(True)
        w = _tf32(torch.randn(N, Kd) * 0.1, K).cuda().requires_grad_(True)
        bias = torch.randn(N).cuda().requires_grad_(True)
        gy = _tf32(torch.randn(M, N), K).cuda().requires_grad_(True)
        y = SF.MmNT.apply(a, w, bias)
        ref = a.double() @ w.double().t() + bias.double()
        assert _rel(y, ref) < 1e-4
        da, dw, db = torch.autograd.grad(y, [a, w, bias], gy, create_graph=True)
        assert _rel(da, gy.double() @ w.double()) < 1e-4
        assert _rel(dw, gy.double().t() @ a.double()) < 1e-4
        assert _rel(db, gy.double().sum(0)) < 1e-4
        # second order: d/d(gy), d/dw of <da, v>
        v = _tf32(torch.randn(M, Kd), K).cuda()
        d_gy, d_w = torch.autograd.grad(da, [gy, w], v)
        assert _rel(d_gy, v.double() @ w.double().t()) < 1e-4
        assert _rel(d_w, gy.double().t() @ v.double()) < 1e-4


@pytest.mark.parametrize("B,H,Cin,Cout", [(4, 4, 544, 512), (6, 8, 128, 128), (3, 32, 128, 128)])
def test_conv3x3_family_double_backward(K, B, H, Cin, Cout):
    from contrad_b200 import sg2_functional as SF
    import torch.nn.functional as F
    torch.manual_seed(B + H)
    x = _tf32(torch.randn(B, H, H, Cin), K).cuda().requires_grad_(True)
    w = _tf32(torch.randn(Cout, Cin, 3, 3) * 0.05, K).cuda().requires_grad_(True)
    gy = _tf32(torch.randn(B, H, H, Cout), K).cuda().requires_grad_(True)
    y = SF.Conv3x3.apply(x, w)
    xd, wd, gd = (t.detach().double() for t in (x, w, gy))
    ref = F.conv2d(xd.permute(0, 3, 1, 2), wd, padding=1).permute(0, 2, 3, 1)
    assert _rel(y, ref) < 1e-4
    dx, dw = torch.autograd.grad(y, [x, w], gy, create_graph=True)
    ref_dx = torch.nn.grad.conv2d_input((B, Cin, H, H), wd, gd.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    ref_dw = torch.nn.grad.conv2d_weight(xd.permute(0, 3, 1, 2), (Cout, Cin, 3, 3), gd.permute(0, 3, 1, 2), padding=1)
    assert _rel(dx, ref_dx) < 1e-4 and _rel(dw, ref_dw) < 1e-4
    v = _tf32(torch.randn(B, H, H, Cin), K).cuda()
    d_gy, d_w = torch.autograd.grad(dx, [gy, w], v)
    vd = v.double()
    assert _rel(d_gy, F.conv2d(vd.permute(0, 3, 1, 2), wd, padding=1).permute(0, 2, 3, 1)) < 1e-4
    assert _rel(d_w, torch.nn.grad.conv2d_weight(vd.permute(0, 3, 1, 2), (Cout, Cin, 3, 3), gd.permute(0, 3, 1, 2), padding=1)) < 1e-4


# ------------------------------------------------------------------------------------------------ models vs the reference
@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN, "stylegan2_small.pt"), weights_only=False)


@pytest.fixture(scope="module")
def models(fx):
    from oracle import stylegan2_oracle as SO
    from contrad_b200.models.gan import get_architecture
    sd_d = SO.make_d_state(fx["size"], small32=True, d_hidden=512, generator=torch.Generator().manual_seed(fx["w_seed_d"]))
    sd_g = SO.make_g_state(fx["size"], small32=True, generator=torch.Generator().manual_seed(fx["w_seed_g"]))
    gen = torch.Generator().manual_seed(fx["bias_seed"])
    for sd in (sd_d, sd_g):
        for k in sd:
            if k.endswith(".bias") and sd[k].abs().sum() == 0:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
            if k.endswith("noise.weight"):
                sd[k] = 0.1 * torch.randn(1, generator=gen)
    G, D = get_architecture("stylegan2", (fx["size"], fx["size"], 3))
    D.load_state_dict(sd_d, strict=True)
    G.load_state_dict(sd_g, strict=True)
    return G.cuda().train(), D.cuda().train()


def _l2(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _norm_errs(grads, want, tag=None):
    """worst relative deviation of per-parameter gradient norms; grads: dict name -> tensor."""
    errs = {}
    for k, n in want.items():
        if n <= 0:
            continue
        assert grads.get(k) is not None, "no gradient for %s" % k
        errs[k] = abs(float(grads[k].norm()) - n) / n
    if tag:
        REPORT[tag + ".worst_keys"] = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    return max(errs.values())


def _no_noise(want):
    """NoiseInjection weights are scalars whose gradient is a sum of random-sign terms that cancels almost completely
    (the noise is independent of everything else): their relative error is unbounded under any rounding, they are
    checked against the scale of the largest of them instead (grad_noise_w)."""
    return {k: v for k, v in want.items() if not k.endswith("noise.weight")}


def _total_norm(grads):
    return float(torch.stack([g.detach().double().pow(2).sum() for g in grads if g is not None]).sum().sqrt())


class _Checks(object):
    """Collects (name, error, tolerance) triples, records them in REPORT, asserts once at the end so that one run
    reports every deviation."""

    def __init__(self, prefix):
        self.prefix, self.rows = prefix, []

    def add(self, name, err, tol):
        REPORT["%s.%s" % (self.prefix, name)] = err
        self.rows.append((name, err, tol))

    def finish(self):
        bad = [(n, e, t) for n, e, t in self.rows if not e < t]
        assert not bad, "%s: %s" % (self.prefix, ", ".join("%s %.3e >= %.1e" % r for r in bad))


@pytest.fixture(scope="module")
def tf32_oracle(fx, models):
    """The reference arithmetic as the reference itself runs it on a GPU: the oracle's torch ops on CUDA with TF32
    convolutions / matmuls allowed.  Its deviation from the fp32 CPU fixtures is the yardstick for quantities that
    are ill-conditioned under ANY TF32 path (a LeakyReLU pre-activation within TF32 rounding of zero takes the
    other slope: individual gradient elements move by a factor 5 while norms stay put)."""
    from oracle import stylegan2_oracle as SO
    G, D = models
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out = {}
    try:
        size = fx["size"]
        leaf = lambda m: {k: (v.detach().clone().requires_grad_(True) if not k.endswith(".kernel") else v.detach().clone())
                          for k, v in m.state_dict().items()}
        # discriminator case
        c = fx["d_case"]
        sd = leaf(D)
        x = c["x"].cuda().requires_grad_(True)
        d, p1, p2 = SO.d_forward(sd, x, size)
        ((d * c["c_d"].cuda()).sum() + (p1 * c["c1"].cuda()).sum() + (p2 * c["c2"].cuda()).sum()).backward()
        out["D.d"] = _rel(d, c["d"])
        out["D.dx_l2"] = _l2(x.grad, c["dx"])
        out["D.grad_from_rgb_l2"] = _l2(sd["layers.0.0.weight"].grad, c["grad_from_rgb"])
        out["D.grad_last_bias_l2"] = _l2(sd["last_conv.1.bias"].grad, c["grad_last_bias"])
        out["D.norms"] = _norm_errs({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"])
        # R1 case
        c = fx["r1_case"]
        sd = leaf(D)
        r1 = SO.r1_penalty(sd, c["x"].cuda(), size)
        r1.mean().backward()
        out["R1.per_sample"] = _rel(r1, c["per_sample"])
        out["R1.grad_from_rgb_l2"] = _l2(sd["layers.0.0.weight"].grad, c["grad_from_rgb"])
        out["R1.grad_conv1_bias_l2"] = _l2(sd["layers.1.conv1.1.bias"].grad, c["grad_conv1_bias"])
        out["R1.norms"] = _norm_errs({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"])
        # generator case
        c = fx["g_case"]
        sd = leaf(G)
        img = SO.g_forward(sd, c["z"].cuda(), size, [n.cuda() for n in c["noises"]], z_mix=c["z_mix"].cuda(),
                           mix_layer=c["mix_layer"])
        (img * c["c_img"].cuda()).sum().backward()
        out["G.image"] = _rel(img, c["image"])
        out["G.grad_const_l2"] = _l2(sd["input.const"].grad, c["grad_const"])
        nwk = ["conv1.noise.weight"] + ["layers.%d.noise.weight" % i for i in range(len(G.layers))]
        out["G.grad_noise_w"] = _rel(torch.stack([sd[k].grad for k in nwk]), c["grad_noise_w"])
        out["G.norms"] = _norm_errs({k: v.grad for k, v in sd.items() if v.requires_grad}, _no_noise(c["grad_norms"]),
                                    "torch_tf32.G")
        # D-step case
        c = fx["dstep_case"]
        sd = leaf(D)
        d_loss, pen, _, _ = SO.gd_losses(sd, size, c["real_aug2"].cuda(), c["fake_aug"].cuda())
        (d_loss + pen).backward()
        out["dstep.norms"] = _norm_errs({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"], "torch_tf32.dstep")
        # D-step objective at n = 16 (+ 0.05 * R1)
        c = fx["dstep16_case"]
        sd = leaf(D)
        real2, fake = c["real_aug2"].float().cuda(), c["fake_aug"].float().cuda()
        d_loss, pen, _, _ = SO.gd_losses(sd, size, real2, fake)
        r1 = SO.r1_penalty(sd, real2[:fake.shape[0]], size).mean()
        (d_loss + pen + 0.05 * r1).backward()
        grads = {k: v.grad for k, v in sd.items() if v.requires_grad}
        out["dstep16.norms"] = _norm_errs(grads, c["grad_norms"], "torch_tf32.dstep16")
        out["dstep16.total_norm"] = abs(_total_norm(grads.values()) - c["total_grad_norm"]) / c["total_grad_norm"]
        out["dstep16.d_loss"] = abs(float(d_loss) - c["d_loss"]) / abs(c["d_loss"])
        out["dstep16.r1"] = abs(float(r1) - c["r1"]) / abs(c["r1"])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    for k, v in out.items():
        REPORT["torch_tf32." + k] = v
    return out


def _tol(base, yard):
    """Tolerance for an ill-conditioned quantity: the documented base, or three times what torch's own TF32 GPU
    arithmetic shows on the same fixture, whichever is larger (at these tiny batches cuBLAS / cuDNN fall back to fp32
    SIMT kernels for several of the small GEMMs, so the yardstick is partly fp32 while every contraction here is
    TF32)."""
    return max(base, 3.0 * yard)


def test_discriminator_matches_reference(fx, models, tf32_oracle):
    G, D = models
    c = fx["d_case"]
    D.zero_grad()
    x = c["x"].cuda().requires_grad_(True)
    d, aux = D(x, projection=True, projection2=True, penultimate=True)
    ck = _Checks("D")
    ck.add("d", _rel(d, c["d"]), 1e-2)
    ck.add("projection", _rel(aux["projection"], c["projection"]), 1e-2)
    ck.add("projection2", _rel(aux["projection2"], c["projection2"]), 1e-2)
    ck.add("penultimate", _rel(aux["penultimate"], c["penultimate"]), 1e-2)
    ((d * c["c_d"].cuda()).sum() + (aux["projection"] * c["c1"].cuda()).sum() + (aux["projection2"] * c["c2"].cuda()).sum()).backward()
    ck.add("dx_l2", _l2(x.grad, c["dx"]), _tol(5e-2, tf32_oracle["D.dx_l2"]))
    ck.add("grad_from_rgb_l2", _l2(D.layers[0][0].weight.grad, c["grad_from_rgb"]), _tol(5e-2, tf32_oracle["D.grad_from_rgb_l2"]))
    ck.add("grad_last_bias_l2", _l2(D.last_conv[1].bias.grad, c["grad_last_bias"]), _tol(5e-2, tf32_oracle["D.grad_last_bias_l2"]))
    ck.add("norms", _norm_errs({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"]), _tol(2e-2, tf32_oracle["D.norms"]))
    ck.finish()


def test_r1_double_backward_matches_reference(fx, models, tf32_oracle):
    from contrad_b200.training.gan import stylegan2 as T
    G, D = models
    c = fx["r1_case"]
    D.zero_grad()
    per_sample = T.r1_per_sample(D, c["x"].cuda(), lambda t: t)
    ck = _Checks("R1")
    ck.add("per_sample", _rel(per_sample, c["per_sample"]), _tol(1e-2, tf32_oracle["R1.per_sample"]))
    per_sample.mean().backward()
    ck.add("grad_from_rgb_l2", _l2(D.layers[0][0].weight.grad, c["grad_from_rgb"]), _tol(5e-2, tf32_oracle["R1.grad_from_rgb_l2"]))
    ck.add("grad_conv1_bias_l2", _l2(D.layers[1].conv1[1].bias.grad, c["grad_conv1_bias"]),
           _tol(1e-1, tf32_oracle["R1.grad_conv1_bias_l2"]))
    ck.add("norms", _norm_errs({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"]), _tol(5e-2, tf32_oracle["R1.norms"]))
    ck.finish()


def test_generator_matches_reference(fx, models, tf32_oracle):
    G, D = models
    c = fx["g_case"]
    G.zero_grad()
    noises = [n.cuda() for n in c["noises"]]
    z_mix = c["z_mix"].cuda()
    orig = G.sample_latent
    G.sample_latent = lambda n: z_mix                       # the device generator differs from the CPU reference run
    try:
        torch.set_rng_state(c["rng_state_after_zmix"])      # CPU draws of the mixing mask (generator.py:257-258)
        img, latents = G(c["z"].cuda(), return_latents=True, style_mix=0.9, noise=noises)
    finally:
        G.sample_latent = orig
    ck = _Checks("G")
    ck.add("latents", _rel(latents, c["latents"]), 1e-2)
    ck.add("image", _rel(img, c["image"]), 1e-2)
    (img * c["c_img"].cuda()).sum().backward()
    ck.add("grad_const_l2", _l2(G.input.const.grad, c["grad_const"]), _tol(5e-2, tf32_oracle["G.grad_const_l2"]))
    nw = torch.stack([G.conv1.noise.weight.grad] + [l.noise.weight.grad for l in G.layers])
    ck.add("grad_noise_w", _rel(nw, c["grad_noise_w"]), _tol(1e-1, tf32_oracle["G.grad_noise_w"]))
    ck.add("grad_rgb_bias", _rel(G.to_rgbs[-1].bias.grad, c["grad_rgb_bias"]), 2e-2)
    ck.add("norms", _norm_errs({k: p.grad for k, p in G.named_parameters()}, _no_noise(c["grad_norms"]), "G"), _tol(2e-2, tf32_oracle["G.norms"]))
    ck.add("image_nomix", _rel(G(c["z"].cuda(), style_mix=0.0, noise=noises), c["image_nomix"]), 1e-2)
    ck.finish()


def test_dstep_losses_match_reference(fx, models, tf32_oracle):
    from contrad_b200.training.gan import stylegan2 as T
    G, D = models
    c = fx["dstep_case"]
    D.zero_grad()
    d_all, view_r, view_f = T.discriminate(D, c["real_aug2"].cuda(), c["fake_aug"].cuda())
    P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
    d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
    ck = _Checks("dstep")
    for name, got, want in (("d_loss", d_loss, c["d_loss"]), ("penalty", aux["penalty"], c["penalty"]),
                            ("d_real", aux["d_real"], c["d_real"]), ("d_gen", aux["d_gen"], c["d_gen"])):
        scale = max(abs(want), 1.0) if name in ("d_real", "d_gen") else abs(want)
        ck.add(name, abs(float(got.detach()) - want) / scale, 1e-3)        # north_star: 1e-3 relative on the losses
    (d_loss + aux["penalty"]).backward()
    ck.add("norms", _norm_errs({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], "dstep"), _tol(5e-2, tf32_oracle["dstep.norms"]))
    g_l = T.loss_G_fn(T.discriminate(D, None, c["fake_aug"].cuda(), train_G=True))
    ck.add("g_loss", abs(float(g_l.detach()) - c["g_loss"]) / abs(c["g_loss"]), 1e-3)
    ck.finish()


def test_dstep16_objective_matches_reference(fx, models, tf32_oracle):
    """contrastive + L_dis + 0.05 * R1 at n = 16: loss scalars to 1e-3 (north_star), the total D gradient norm to 1e-2."""
    from contrad_b200.training.gan import stylegan2 as T
    G, D = models
    c = fx["dstep16_case"]
    real2, fake = c["real_aug2"].float().cuda(), c["fake_aug"].float().cuda()
    n = fake.shape[0]
    D.zero_grad()
    d_all, view_r, view_f = T.discriminate(D, real2, fake)
    P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
    d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
    r1 = T.r1_loss(D, real2[:n], lambda t: t)
    ck = _Checks("dstep16")
    ck.add("d_loss", abs(float(d_loss.detach()) - c["d_loss"]) / abs(c["d_loss"]), 1e-3)
    ck.add("penalty", abs(float(aux["penalty"].detach()) - c["penalty"]) / abs(c["penalty"]), 1e-3)
    ck.add("r1", abs(float(r1.detach()) - c["r1"]) / abs(c["r1"]), _tol(1e-2, tf32_oracle["dstep16.r1"]))
    (d_loss + aux["penalty"] + 0.05 * r1).backward()
    grads = {k: p.grad for k, p in D.named_parameters()}
    ck.add("norms", _norm_errs(grads, c["grad_norms"], "dstep16"), _tol(3e-2, tf32_oracle["dstep16.norms"]))
    ck.add("total_norm", abs(_total_norm(grads.values()) - c["total_grad_norm"]) / c["total_grad_norm"],
           _tol(1e-2, tf32_oracle["dstep16.total_norm"]))
    ck.finish()


def test_strict_precision_meets_fixed_tolerances(fx, models):
    """The full strict precision mode (contrad_b200/precision.py: every GEMM Function of sg2_functional evaluates hi/lo-split
    "3xTF32" operands, producers stop rounding) against the reference fixtures with FIXED bars - no yardstick: outputs, R1
    per sample, loss scalars and the total D gradient norm of the n = 16 objective 1e-3 (north_star), worst per-parameter
    gradient norm 2e-3 for D / 4e-3 for the R1 double backward, the n = 16 objective and G (measured 3.0e-3 on the
    smallest tensors)."""
    from contrad_b200 import precision
    from contrad_b200.training.gan import stylegan2 as T
    G, D = models
    ck = _Checks("strict")
    with precision.strict("full"):
        c = fx["d_case"]
        D.zero_grad()
        x = c["x"].cuda().requires_grad_(True)
        d, aux = D(x, projection=True, projection2=True, penultimate=True)
        ck.add("D.d", _rel(d, c["d"]), 1e-3)
        ck.add("D.projection", _rel(aux["projection"], c["projection"]), 1e-3)
        ck.add("D.penultimate", _rel(aux["penultimate"], c["penultimate"]), 1e-3)
        ((d * c["c_d"].cuda()).sum() + (aux["projection"] * c["c1"].cuda()).sum() + (aux["projection2"] * c["c2"].cuda()).sum()).backward()
        ck.add("D.dx_l2", _l2(x.grad, c["dx"]), 2e-3)
        ck.add("D.norms", _norm_errs({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], "strict.D"), 2e-3)

        c = fx["r1_case"]
        D.zero_grad()
        per_sample = T.r1_per_sample(D, c["x"].cuda(), lambda t: t)
        ck.add("R1.per_sample", _rel(per_sample, c["per_sample"]), 1e-3)
        per_sample.mean().backward()
        ck.add("R1.norms", _norm_errs({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], "strict.R1"), 4e-3)

        c = fx["dstep16_case"]
        real2, fake = c["real_aug2"].float().cuda(), c["fake_aug"].float().cuda()
        n = fake.shape[0]
        D.zero_grad()
        d_all, view_r, view_f = T.discriminate(D, real2, fake)
        P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
        d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
        r1 = T.r1_loss(D, real2[:n], lambda t: t)
        ck.add("dstep16.d_loss", abs(float(d_loss.detach()) - c["d_loss"]) / abs(c["d_loss"]), 1e-3)
        ck.add("dstep16.r1", abs(float(r1.detach()) - c["r1"]) / abs(c["r1"]), 1e-3)
        (d_loss + aux["penalty"] + 0.05 * r1).backward()
        grads = {k: p.grad for k, p in D.named_parameters()}
        ck.add("dstep16.norms", _norm_errs(grads, c["grad_norms"], "strict.dstep16"), 4e-3)
        ck.add("dstep16.total_norm", abs(_total_norm(grads.values()) - c["total_grad_norm"]) / c["total_grad_norm"], 1e-3)

        c = fx["g_case"]
        G.zero_grad()
        noises = [t.cuda() for t in c["noises"]]
        z_mix = c["z_mix"].cuda()
        orig = G.sample_latent
        G.sample_latent = lambda k: z_mix
        try:
            torch.set_rng_state(c["rng_state_after_zmix"])
            img = G(c["z"].cuda(), style_mix=0.9, noise=noises)
        finally:
            G.sample_latent = orig
        ck.add("G.image", _rel(img, c["image"]), 1e-3)
        (img * c["c_img"].cuda()).sum().backward()
        ck.add("G.norms", _norm_errs({k: p.grad for k, p in G.named_parameters()}, _no_noise(c["grad_norms"]), "strict.G"), 4e-3)
    ck.finish()


def _seeded_states(size, small32, cm, seeds):
    from oracle import stylegan2_oracle as SO
    kw = dict(small32=True) if small32 else dict(small32=False, channel_multiplier=cm)
    sd_d = SO.make_d_state(size, d_hidden=512, generator=torch.Generator().manual_seed(seeds[0]), **kw)
    sd_g = SO.make_g_state(size, generator=torch.Generator().manual_seed(seeds[1]), **kw)
    gen = torch.Generator().manual_seed(seeds[2])
    for sd in (sd_d, sd_g):
        for k in sd:
            if k.endswith(".bias") and sd[k].abs().sum() == 0:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
            if k.endswith("noise.weight"):
                sd[k] = 0.1 * torch.randn(1, generator=gen)
    return sd_d, sd_g


@pytest.mark.parametrize("mode", [0, "full"])
def test_config4_dstep_at_batch_64_vs_oracle(mode):
    """BASELINE config 4 at its REAL batch (c10_style64.gin: n = 64, D sees 64 fakes and 128 real views, R1 on 64 images,
    --lbd_r1 0.1 --no_lazy): the D-step objective of train_stylegan2_contraD.py:207-236 on SimCLR views (oracle chain) of
    diverse synthetic images against
    the fp32 CPU oracle (oracle/stylegan2_oracle.py, pinned on the reference fixtures at n = 4 / 16).  Loss scalars
    1e-3 (north_star) in both modes; total D gradient norm 2e-3 (default) / 1e-3 (full); R1 1e-2 / 2e-3."""
    from oracle import stylegan2_oracle as SO
    from contrad_b200 import precision
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import stylegan2 as T
    size, n = 32, 64
    sd_d, _ = _seeded_states(size, True, None, (41, 42, 43))
    from oracle import contrad_oracle as CO
    torch.manual_seed(44); np.random.seed(44)

    def fields(m):          # diverse low-frequency images (white noise collapses every embedding onto one point: the
        base = F.interpolate(torch.rand(m, 3, 4, 4), size=(size, size), mode="bilinear", align_corners=False)   # contrastive
        return (0.15 + 0.7 * base + 0.1 * torch.rand(m, 3, size, size)).clamp(0, 1)      # gradient is then a 70x smaller residual)

    imgs = fields(n)
    p_r, o_r = CO.sample_simclr_params(2 * n, size, size)
    real2 = CO.augment_simclr(torch.cat([imgs, imgs]), p_r, o_r)          # two SimCLR views of every real image
    p_f, o_f = CO.sample_simclr_params(n, size, size)
    fake = CO.augment_simclr(fields(n), p_f, o_f)
    leaf = {k: (v.clone().requires_grad_(True) if not k.endswith(".kernel") else v) for k, v in sd_d.items()}
    d_loss_o, pen_o, _, _ = SO.gd_losses(leaf, size, real2, fake)
    r1_o = SO.r1_penalty(leaf, real2[:n], size).mean()
    (d_loss_o + pen_o + 0.05 * r1_o).backward()                       # 0.5 * lbd_r1 * r1 * d_reg_every with lbd_r1 = 0.1
    want = _total_norm([v.grad for v in leaf.values() if v.requires_grad and v.grad is not None])
    _, D = get_architecture("stylegan2", (size, size, 3))
    D.load_state_dict(sd_d, strict=True)
    D.cuda().train()
    ck = _Checks("config4_b64.%s" % mode)
    with precision.strict(mode):
        d_all, view_r, view_f = T.discriminate(D, real2.cuda(), fake.cuda())
        P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
        d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
        r1 = T.r1_loss(D, real2[:n].cuda(), lambda t: t)
        (d_loss + aux["penalty"] + 0.05 * r1).backward()
    # measured on the B200 (profiles/sg2_parity_r2.json): d_loss 3e-6 / 1e-6, penalty 4e-6 / 3e-7, r1 1e-4 / 6e-4, total
    # gradient norm 7e-5 / 2e-5 (default / full); R1 is a ~1e-3-sized double-backward quantity, hence its wider bar
    ck.add("d_loss", abs(float(d_loss.detach()) - float(d_loss_o)) / abs(float(d_loss_o)), 1e-3)
    ck.add("penalty", abs(float(aux["penalty"].detach()) - float(pen_o)) / abs(float(pen_o)), 1e-3)
    ck.add("r1", abs(float(r1.detach()) - float(r1_o)) / abs(float(r1_o)), 2e-3 if mode == "full" else 1e-2)
    ck.add("total_norm", abs(_total_norm([p.grad for p in D.parameters()]) - want) / want, 1e-3 if mode == "full" else 2e-3)
    ck.finish()


def test_config5_per_replica_batch_vs_oracle():
    """BASELINE config 5 per replica (b64 over 8 GPUs = 8 images of 512x512 per replica, `stylegan2_512`): what ONE
    DataParallel replica evaluates in the D step - D on 8 fakes and on 16 real views (two minibatch-stddev groupings) and
    the `simclr` augmentation at 512x512 - against the fp32 CPU oracle.  D outputs and embeddings 1e-2 of their scale
    (single-pass TF32 through 7 ResBlocks), L_dis 1e-3; the augmentation against the oracle chain on the same draws 2e-5."""
    from oracle import contrad_oracle as CO
    from oracle import stylegan2_oracle as SO
    from contrad_b200.functional import AugmentSimCLRFn
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import stylegan2 as T
    size, n = 512, 8
    sd_d, _ = _seeded_states(size, False, 1.0, (51, 52, 53))
    torch.manual_seed(54); np.random.seed(54)
    real, fake = torch.rand(n, 3, size, size), torch.rand(n, 3, size, size)
    # the large-image augmentation on explicit draws (RRC scale (0.08, 1), jitter 0.8 / 0.8 / 0.8 / 0.2: afhq_dog_style64.gin)
    params, order = CO.sample_simclr_params(n, size, size, scale=(0.08, 1.0), brightness=0.8, contrast=0.8, saturation=0.8,
                                            hue=0.2)
    aug_o = CO.augment_simclr(real, params, order)
    aug = AugmentSimCLRFn.apply(real.cuda(), CO.pack_params(params).cuda(), order)
    ck = _Checks("config5_replica")
    ck.add("augment_512", float((aug.cpu() - aug_o).abs().max()), 2e-5)
    real2 = torch.cat([aug_o, aug_o.flip(3)])
    with torch.no_grad():
        d_gen_o, others_o, fakes_o = SO.d_forward(sd_d, fake, size, sg_linear=True)
        d_rs_o, views_o, reals_o = SO.d_forward(sd_d, real2, size, sg_linear=True)
    pen_o = F.softplus(d_gen_o).mean() + F.softplus(-d_rs_o[:n]).mean()
    _, D = get_architecture("stylegan2_512", (size, size, 3))
    D.load_state_dict(sd_d, strict=True)
    D.cuda().train()
    with torch.no_grad():
        d_all, view_r, view_f = T.discriminate(D, real2.cuda(), fake.cuda())
    ck.add("d_real", _rel(d_all[0], d_rs_o[:n]), 1e-2)
    ck.add("d_gen", _rel(d_all[1], d_gen_o), 1e-2)
    ck.add("views_real", _rel(view_r[0], F.normalize(views_o)[:n]), 1e-2)
    ck.add("fakes", _rel(view_f[-1], F.normalize(fakes_o)), 1e-2)
    P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
    _, aux = T.loss_D_fn(P, d_all, view_r, view_f)
    ck.add("penalty", abs(float(aux["penalty"]) - float(pen_o)) / abs(float(pen_o)), 1e-3)
    ck.finish()


def test_eager_gpu_yardstick_timing(models):
    """Not a parity check: times the D-step objective (n = 64, contrastive + L_dis + R1, forward + double backward)
    through this library and through the oracle's torch ops on the same GPU (cuDNN / cuBLAS, TF32 allowed = what the
    reference executes on a GPU) and records both in the report."""
    from oracle import stylegan2_oracle as SO
    from contrad_b200.training.gan import stylegan2 as T
    G, D = models
    n = 64
    torch.manual_seed(0)
    real2, fake = torch.rand(2 * n, 3, 32, 32, device="cuda"), torch.rand(n, 3, 32, 32, device="cuda")
    P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)

    def ours():
        D.zero_grad(set_to_none=True)
        d_all, vr, vf = T.discriminate(D, real2, fake)
        d_loss, aux = T.loss_D_fn(P, d_all, vr, vf)
        (d_loss + aux["penalty"] + 0.05 * T.r1_loss(D, real2[:n], lambda t: t)).backward()

    leaf = {k: (v.detach().clone().requires_grad_(True) if not k.endswith(".kernel") else v.detach().clone())
            for k, v in D.state_dict().items()}

    def torch_eager():
        for v in leaf.values():
            v.grad = None
        d_loss, pen, _, _ = SO.gd_losses(leaf, 32, real2, fake)
        (d_loss + pen + 0.05 * SO.r1_penalty(leaf, real2[:n], 32).mean()).backward()

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for name, fn in (("timing.ours_dstep_ms", ours), ("timing.torch_eager_dstep_ms", torch_eager)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            REPORT[name] = e0.elapsed_time(e1) / 5
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert REPORT["timing.ours_dstep_ms"] > 0


def test_full_stylegan2_step_runs():
    """train_step_stylegan2: G step, D step with R1 (step % d_reg_every == 0) and EMA, on the fused augmentation."""
    import copy
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
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
    G, D = get_architecture("stylegan2", (32, 32, 3))
    G.cuda(); D.cuda()
    g_ema = copy.deepcopy(G)
    aug = get_augment(mode="simclr").cuda()
    GD = T.G_D(G, D, aug)
    P = SimpleNamespace(use_warmup=True, halflife_lr=0, ema_start_k=0, accum=0.5 ** (64 / 20000.0), d_reg_every=1,
                        lbd_r1=0.1, style_mix=0.9, temp=0.1, lbd_a=1.0, distributed=False)
    opt = {"warmup": 3000, "lr": 2e-3, "lr_d": 2e-3, "batch_size": 8}
    opt_G = FusedAdam(G.parameters(), lr=2e-3, betas=(0.0, 0.99))
    opt_D = FusedAdam(D.parameters(), lr=2e-3, betas=(0.0, 0.99))
    before_d = {k: v.detach().clone() for k, v in D.named_parameters()}
    ema_before = g_ema.input.const.detach().clone()
    launches = _capi.launch_count()
    for step in (1, 2):
        out = engine.train_step_stylegan2(P, opt, GD, g_ema, (opt_G, opt_D), torch.rand(8, 3, 32, 32, device="cuda"), step,
                                          record_grad_norms=True)
    torch.cuda.synchronize()
    for k, v in out.items():
        assert torch.isfinite(v).all(), (k, v)
    assert "d_r1" in out and float(out["d_r1"]) > 0
    assert _capi.launch_count() - launches > 100
    changed = [k for k, v in D.named_parameters() if not torch.equal(v.detach(), before_d[k])]
    assert len(changed) >= 0.9 * len(before_d), set(before_d) - set(changed)
    REPORT["step.unchanged_params"] = sorted(set(before_d) - set(changed))
    assert not torch.equal(g_ema.input.const, ema_before)
    REPORT["step.values"] = {k: float(v) for k, v in out.items()}


def test_stylegan2_512_matches_oracle():
    """BASELINE config 5 architecture (`stylegan2_512`, channel_multiplier 1.0, 512x512): thin 32/64-channel layers at
    512^2 / 256^2 (zero-extended weight-gradient operands, BN=32/64 tap-GEMM tiles), 7 ResBlocks, 15 style layers.
    D (B=4) and G (B=2) forward + backward against the fp32 CPU oracle; the R1 double backward and the large-image
    augmentation run on the same inputs."""
    from oracle import stylegan2_oracle as SO
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import stylegan2 as T
    size = 512
    sd_d = SO.make_d_state(size, small32=False, channel_multiplier=1.0, d_hidden=512, generator=torch.Generator().manual_seed(7))
    sd_g = SO.make_g_state(size, small32=False, channel_multiplier=1.0, generator=torch.Generator().manual_seed(8))
    gen = torch.Generator().manual_seed(9)
    for sd in (sd_d, sd_g):
        for k in sd:
            if k.endswith(".bias") and sd[k].abs().sum() == 0:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
            if k.endswith("noise.weight"):
                sd[k] = 0.1 * torch.randn(1, generator=gen)
    G, D = get_architecture("stylegan2_512", (size, size, 3))
    D.load_state_dict(sd_d, strict=True); G.load_state_dict(sd_g, strict=True)
    G.cuda().train(); D.cuda().train()
    ck = _Checks("sg2_512")
    # ---- discriminator
    torch.manual_seed(3)
    x = torch.rand(4, 3, size, size)
    c1 = torch.randn(4, 128)
    leaf = {k: (v.clone().requires_grad_(True) if not k.endswith(".kernel") else v) for k, v in sd_d.items()}
    d_o, p1_o, _ = SO.d_forward(leaf, x, size)
    (d_o.sum() + (p1_o * c1).sum()).backward()
    d, aux = D(x.cuda(), projection=True)
    (d.sum() + (aux["projection"] * c1.cuda()).sum()).backward()
    ck.add("D.d", _rel(d, d_o), 1e-2)
    ck.add("D.projection", _rel(aux["projection"], p1_o), 1e-2)
    ck.add("D.norms", _norm_errs({k: p.grad for k, p in D.named_parameters()},
                                 {k: float(v.grad.norm()) for k, v in leaf.items() if v.requires_grad and v.grad is not None},
                                 "sg2_512.D"), 5e-2)
    # ---- R1 at 512x512 (double backward through every operator) and the large-image augmentation in front of it
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    import gin
    from contrad_b200.augment import get_augment
    gin.clear_config()
    gin.parse_config("RandomResizeCropLayer.scale = (0.08, 1.0)\nColorJitterLayer.brightness = 0.8\n"
                     "ColorJitterLayer.contrast = 0.8\nColorJitterLayer.saturation = 0.8\nColorJitterLayer.hue = 0.2\n")
    D.zero_grad()
    r1 = T.r1_per_sample(D, x.cuda(), get_augment(mode="simclr").cuda())
    r1.mean().backward()
    assert torch.isfinite(r1).all() and float(r1.min()) > 0
    assert all(torch.isfinite(p.grad).all() for p in D.parameters() if p.grad is not None)
    assert float(D.layers[1].conv1[0].weight.grad.abs().sum()) > 0
    REPORT["sg2_512.r1"] = [float(v) for v in r1]
    # ---- generator
    torch.manual_seed(4)
    z = torch.randn(2, 512)
    noises = [torch.randn(*s_) for s_ in SO.noise_shapes(size, 2)]
    c_img = torch.randn(2, 3, size, size)
    leaf = {k: (v.clone().requires_grad_(True) if not k.endswith(".kernel") else v) for k, v in sd_g.items()}
    img_o = SO.g_forward(leaf, z, size, noises)
    (img_o * c_img).sum().backward()
    img = G(z.cuda(), style_mix=0.0, noise=[n_.cuda() for n_ in noises])
    (img * c_img.cuda()).sum().backward()
    ck.add("G.image", _rel(img, img_o), 1e-2)
    ck.add("G.norms", _norm_errs({k: p.grad for k, p in G.named_parameters()},
                                 _no_noise({k: float(v.grad.norm()) for k, v in leaf.items() if v.requires_grad and v.grad is not None}),
                                 "sg2_512.G"), 5e-2)
    ck.finish()


def test_gd_under_data_parallel():
    """BASELINE config 5 drives G_D through nn.DataParallel (train_stylegan2_contraD.py:300-306): one host thread per
    GPU in ONE process.  Needs >= 2 GPUs (`gpurun --gpus 2`); checks that the replicas run concurrently on their own
    devices / streams and that the gathered outputs back-propagate into the master parameters."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    import gin
    from contrad_b200.augment import get_augment
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import stylegan2 as T
    from contrad_b200 import engine
    gin.clear_config()
    gin.parse_config("RandomResizeCropLayer.scale = (0.2, 1.0)\nColorJitterLayer.brightness = 0.4\n"
                     "ColorJitterLayer.contrast = 0.4\nColorJitterLayer.saturation = 0.4\nColorJitterLayer.hue = 0.1\n")
    torch.manual_seed(0); np.random.seed(0)
    G, D = get_architecture("stylegan2", (32, 32, 3))
    G.cuda(0); D.cuda(0)
    GD = torch.nn.DataParallel(T.G_D(G, D, get_augment(mode="simclr").cuda(0)), device_ids=[0, 1])
    P = SimpleNamespace(temp=0.1, lbd_a=1.0, distributed=False)
    images = torch.rand(16, 3, 32, 32, device="cuda:0")
    # G step
    engine.set_grad(G, True); engine.set_grad(D, False)
    d_gen = GD(P, images, train_G=True)
    assert d_gen.shape == (16, 1) and d_gen.device.index == 0
    T.loss_G_fn(d_gen).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in G.parameters())
    assert float(G.input.const.grad.abs().sum()) > 0
    # D step + R1
    engine.set_grad(G, False); engine.set_grad(D, True)
    d_all, view_r, view_f = GD(P, images)
    assert d_all[0].shape == (16, 1) and view_r[0].shape == (16, 128) and view_f[2].shape == (16, 128)
    d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
    r1 = GD(P, images, return_r1_loss=True)
    assert r1.shape == (16,)
    (d_loss + aux["penalty"] + 0.05 * r1.mean()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in D.parameters())
    REPORT["dp2.values"] = {"d_loss": float(d_loss.detach()), "penalty": float(aux["penalty"].detach()), "r1": float(r1.mean().detach())}


def test_graphed_stylegan2_step():
    """GraphedStyleGAN2Step: 3 eager steps, capture, replays.  The replays must keep training (parameters move, losses
    stay finite and in the range of the eager steps) and consume fresh host draws (staged style-mixing indices)."""
    import copy
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    import gin
    from contrad_b200 import engine
    from contrad_b200.augment import get_augment
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.optim import FusedAdam
    from contrad_b200.training.gan import stylegan2 as T
    gin.clear_config()
    gin.parse_config("RandomResizeCropLayer.scale = (0.2, 1.0)\nColorJitterLayer.brightness = 0.4\n"
                     "ColorJitterLayer.contrast = 0.4\nColorJitterLayer.saturation = 0.4\nColorJitterLayer.hue = 0.1\n")
    torch.manual_seed(1); np.random.seed(1)
    n = 16
    G, D = get_architecture("stylegan2", (32, 32, 3))
    G.cuda(); D.cuda()
    g_ema = copy.deepcopy(G)
    GD = T.G_D(G, D, get_augment(mode="simclr").cuda())
    P = SimpleNamespace(use_warmup=True, halflife_lr=0, ema_start_k=0, accum=0.99, d_reg_every=1, lbd_r1=0.1, style_mix=0.9,
                        temp=0.1, lbd_a=1.0, distributed=False)
    opt = {"warmup": 3000, "lr": 2e-3, "lr_d": 2e-3, "batch_size": n}
    opts = (FusedAdam(G.parameters(), lr=2e-3, betas=(0.0, 0.99)), FusedAdam(D.parameters(), lr=2e-3, betas=(0.0, 0.99)))
    step_fn = engine.GraphedStyleGAN2Step(P, opt, GD, g_ema, opts)
    logs, snaps = [], []
    for s in range(1, 8):
        out = step_fn(torch.rand(n, 3, 32, 32, device="cuda"), s)
        torch.cuda.synchronize()
        logs.append({k: float(v) for k, v in out.items()})
        snaps.append((D.last_conv[0].weight.detach().clone(), G.input.const.detach().clone(), g_ema.input.const.detach().clone()))
    (variant,) = step_fn.variants.values()
    assert variant.graph is not None and variant.launches_per_replay > 300
    for log in logs:
        assert all(np.isfinite(v) for v in log.values()), log
    # steps 5..7 are replays: every one of them moved D, G and the EMA copy
    for a, b in zip(snaps[3:-1], snaps[4:]):
        assert all(not torch.equal(x, y) for x, y in zip(a, b))
    eager_d = [l["d_loss"] for l in logs[:3]]
    assert all(0.3 * min(eager_d) < l["d_loss"] < 3 * max(eager_d) for l in logs[3:]), logs
    assert int(opts[1].state[D.last_conv[0].weight]["step"]) == 7
    REPORT["graphed.launches_per_replay"] = variant.launches_per_replay
    REPORT["graphed.logs"] = logs
    step_fn.release()


def test_zz_write_report():
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", "sg2_parity.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass
