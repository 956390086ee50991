"""CPU tests of the StyleGAN2 side of the path.

1. The oracle (oracle/stylegan2_oracle.py) replays the fixtures produced by the unmodified reference
   (tests/golden/make_golden_sg2.py -> stylegan2_small.pt): forward values, first-order gradients, the R1
   double backward, the generator with style mixing and the D-step losses of train_stylegan2_contraD.py.
2. The product's HOST logic - the autograd Functions of contrad_b200.sg2_functional and the module mirrors under
   contrad_b200/models/gan/stylegan2 - is run with the kernel bindings replaced by the torch stand-ins of
   tests/cpu_kernels.py and must reproduce the same fixtures (operator wiring, weight re-layouts, NHWC/NCHW
   boundaries, double-backward structure).  The CUDA kernels themselves are checked against the same stand-ins by
   the `-m gpu` tests."""
import json
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import stylegan2_oracle as SO
from tests import cpu_kernels as CK

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN, "stylegan2_small.pt"), weights_only=False)


@pytest.fixture(scope="module")
def states(fx):
    sd_d = SO.make_d_state(fx["size"], small32=True, d_hidden=512, generator=torch.Generator().manual_seed(fx["w_seed_d"]))
    sd_g = SO.make_g_state(fx["size"], small32=True, generator=torch.Generator().manual_seed(fx["w_seed_g"]))
    gen = torch.Generator().manual_seed(fx["bias_seed"])
    for sd in (sd_d, sd_g):
        for k in sd:
            if k.endswith(".bias") and sd[k].abs().sum() == 0:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
            if k.endswith("noise.weight"):
                sd[k] = 0.1 * torch.randn(1, generator=gen)
    return sd_d, sd_g


def _close(a, b, rel=1e-4, what=""):
    a, b = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)
    err = float((a - b).abs().max())
    ref = float(b.abs().max())
    assert err <= rel * max(ref, 1e-6), "%s: max err %.3e vs scale %.3e" % (what, err, ref)


def _leafs(sd):
    return {k: (v.clone().requires_grad_(True) if not k.endswith(".kernel") else v) for k, v in sd.items()}


def _check_norms(named_grads, want, rel, what):
    for k, n in want.items():
        g = named_grads[k]
        assert g is not None, "%s: no gradient for %s" % (what, k)
        got = float(g.norm())
        assert abs(got - n) <= rel * max(n, 1e-6), "%s: |grad %s| = %.6e, reference %.6e" % (what, k, got, n)


# ------------------------------------------------------------------------------------------------ 1. oracle vs reference
def test_oracle_upfirdn2d(fx):
    for c in fx["upfirdn2d"]:
        x = c["x"].clone().requires_grad_(True)
        y = SO.upfirdn2d(x, c["kernel"], up=c["up"], down=c["down"], pad=c["pad"])
        _close(y, c["y"], 1e-6, "upfirdn2d fwd")
        (dx,) = torch.autograd.grad(y, x, c["dy"])
        _close(dx, c["dx"], 1e-6, "upfirdn2d bwd")


def test_oracle_discriminator(fx, states):
    sd = _leafs(states[0])
    c = fx["d_case"]
    x = c["x"].clone().requires_grad_(True)
    d, p1, p2 = SO.d_forward(sd, x, fx["size"])
    _close(d, c["d"], 1e-4, "d"); _close(p1, c["projection"], 1e-4, "projection"); _close(p2, c["projection2"], 1e-4, "projection2")
    _close(SO.d_penultimate(sd, x, fx["size"]), c["penultimate"], 1e-4, "penultimate")
    ((d * c["c_d"]).sum() + (p1 * c["c1"]).sum() + (p2 * c["c2"]).sum()).backward()
    _close(x.grad, c["dx"], 1e-3, "dx")
    _check_norms({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"], 1e-3, "oracle D")


def test_oracle_r1(fx, states):
    sd = _leafs(states[0])
    c = fx["r1_case"]
    r1 = SO.r1_penalty(sd, c["x"], fx["size"])
    _close(r1, c["per_sample"], 1e-3, "r1 per sample")
    r1.mean().backward()
    _check_norms({k: v.grad for k, v in sd.items() if v.requires_grad},
                 {k: n for k, n in c["grad_norms"].items() if n > 0}, 2e-3, "oracle R1")
    _close(sd["layers.0.0.weight"].grad, c["grad_from_rgb"], 2e-3, "r1 grad FromRGB")


def test_oracle_generator(fx, states):
    sd = _leafs(states[1])
    c = fx["g_case"]
    img = SO.g_forward(sd, c["z"], fx["size"], c["noises"], z_mix=c["z_mix"], mix_layer=c["mix_layer"])
    _close(img, c["image"], 1e-4, "image")
    (img * c["c_img"]).sum().backward()
    _check_norms({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"], 1e-3, "oracle G")
    _close(sd["input.const"].grad, c["grad_const"], 1e-3, "grad const")
    _close(SO.g_forward(states[1], c["z"], fx["size"], c["noises"]), c["image_nomix"], 1e-4, "image (no mixing)")


def test_oracle_dstep_losses(fx, states):
    sd = _leafs(states[0])
    c = fx["dstep_case"]
    d_loss, penalty, d_real, d_gen = SO.gd_losses(sd, fx["size"], c["real_aug2"], c["fake_aug"])
    assert abs(float(d_loss) - c["d_loss"]) < 1e-4 * abs(c["d_loss"])
    assert abs(float(penalty) - c["penalty"]) < 1e-4 * abs(c["penalty"])
    assert abs(float(d_real) - c["d_real"]) < 1e-4 * max(1.0, abs(c["d_real"]))
    assert abs(float(d_gen) - c["d_gen"]) < 1e-4 * max(1.0, abs(c["d_gen"]))
    (d_loss + penalty).backward()
    _check_norms({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"], 2e-3, "oracle D-step")
    assert abs(float(SO.g_loss(states[0], fx["size"], c["fake_aug"])) - c["g_loss"]) < 1e-4 * abs(c["g_loss"])


# ------------------------------------------------------------------------------------------------ 2. product host logic
def _product_models(states, size):
    from contrad_b200.models.gan import get_architecture
    G, D = get_architecture("stylegan2", (size, size, 3))
    D.load_state_dict(states[0], strict=True)
    G.load_state_dict(states[1], strict=True)
    return G.train(), D.train()


def test_state_dict_layout_matches_reference():
    from contrad_b200.models.gan import get_architecture
    with open(os.path.join(GOLDEN, "stylegan2_keys.json")) as f:
        keys = json.load(f)
    for arch, size, suffix in (("stylegan2", 32, ""), ("stylegan2_512", 512, "512")):
        G, D = get_architecture(arch, (size, size, 3))
        assert {k: list(v.shape) for k, v in D.state_dict().items()} == keys["D" + suffix]
        assert {k: list(v.shape) for k, v in G.state_dict().items()} == keys["G" + suffix]


def test_cpu_standins_upfirdn2d(fx):
    """The stand-in used as the kernel reference reproduces the reference op, incl. its backward via UpFirDn."""
    from contrad_b200.models.gan.stylegan2.op import upfirdn2d
    with CK.patched():
        for c in fx["upfirdn2d"]:
            x = c["x"].clone().requires_grad_(True)
            y = upfirdn2d(x, c["kernel"], up=c["up"], down=c["down"], pad=c["pad"])
            _close(y, c["y"], 1e-6, "upfirdn2d fwd")
            dy = c["dy"].clone().requires_grad_(True)
            (dx,) = torch.autograd.grad(y, x, dy, create_graph=True)
            _close(dx, c["dx"], 1e-6, "upfirdn2d bwd")
            # the backward of the backward (w.r.t. the cotangent) is the forward operator again
            (ddy,) = torch.autograd.grad(dx, dy, c["x"])
            _close(ddy, c["y"], 1e-6, "upfirdn2d bwd-of-bwd")


def test_product_discriminator_host_logic(fx, states):
    with CK.patched():
        G, D = _product_models(states, fx["size"])
        c = fx["d_case"]
        x = c["x"].clone().requires_grad_(True)
        d, aux = D(x, projection=True, projection2=True, penultimate=True)
        _close(d, c["d"], 1e-4, "d"); _close(aux["projection"], c["projection"], 1e-4, "projection")
        _close(aux["projection2"], c["projection2"], 1e-4, "projection2")
        _close(aux["penultimate"], c["penultimate"], 1e-4, "penultimate")
        ((d * c["c_d"]).sum() + (aux["projection"] * c["c1"]).sum() + (aux["projection2"] * c["c2"]).sum()).backward()
        _close(x.grad, c["dx"], 1e-3, "dx")
        _check_norms({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], 1e-3, "product D")
        _close(D.layers[0][0].weight.grad, c["grad_from_rgb"], 1e-3, "FromRGB grad")
        _close(D.last_conv[1].bias.grad, c["grad_last_bias"], 1e-3, "last bias grad")
        # sg_linear detaches the `linear` head from the backbone (models/gan/base.py:123-126)
        D.zero_grad()
        d = D(c["x"], sg_linear=True)
        (d * c["c_d"]).sum().backward()
        assert D.layers[0][0].weight.grad is None or float(D.layers[0][0].weight.grad.abs().sum()) == 0.0
        assert float(D.linear.l1.weight.grad.abs().sum()) > 0


def test_product_r1_double_backward(fx, states):
    from contrad_b200.training.gan import stylegan2 as T
    with CK.patched():
        G, D = _product_models(states, fx["size"])
        c = fx["r1_case"]
        per_sample = T.r1_per_sample(D, c["x"], lambda t: t)
        _close(per_sample, c["per_sample"], 1e-3, "r1 per sample")
        per_sample.mean().backward()
        _check_norms({k: p.grad for k, p in D.named_parameters()}, {k: n for k, n in c["grad_norms"].items() if n > 0},
                     2e-3, "product R1")
        _close(D.layers[0][0].weight.grad, c["grad_from_rgb"], 2e-3, "r1 grad FromRGB")
        _close(D.layers[1].conv1[1].bias.grad, c["grad_conv1_bias"], 2e-3, "r1 grad conv1 bias")


def test_product_generator_host_logic(fx, states):
    with CK.patched():
        G, D = _product_models(states, fx["size"])
        c = fx["g_case"]
        torch.manual_seed(c["mix_seed"])
        img, latents = G(c["z"], return_latents=True, style_mix=0.9, noise=c["noises"])
        _close(latents, c["latents"], 1e-4, "latents")
        _close(img, c["image"], 1e-4, "image")
        (img * c["c_img"]).sum().backward()
        _check_norms({k: p.grad for k, p in G.named_parameters()}, c["grad_norms"], 1e-3, "product G")
        _close(G.input.const.grad, c["grad_const"], 1e-3, "grad const")
        nw = torch.stack([G.conv1.noise.weight.grad] + [l.noise.weight.grad for l in G.layers])
        _close(nw, c["grad_noise_w"], 1e-3, "noise weight grads")
        _close(G.to_rgbs[-1].bias.grad, c["grad_rgb_bias"], 1e-3, "ToRGB bias grad")
        _close(G(c["z"], style_mix=0.0, noise=c["noises"]), c["image_nomix"], 1e-4, "image (no mixing)")


@pytest.mark.parametrize("exact_weights", [True, False])
def test_product_strict_precision_host_logic(fx, states, exact_weights):
    """The full strict precision mode (contrad_b200/precision.py): every GEMM Function evaluates hi/lo-split operands over
    a concatenated reduction axis.  With the kernel stand-ins (fp32 torch) the discriminator, the R1 double backward and the
    generator must reproduce the reference fixtures as in the default mode; exact_weights=False keeps the real TF32
    rounding inside the split (hi + lo of genuinely rounded parts), i.e. the arithmetic the tensor core will see."""
    from contrad_b200 import precision
    from contrad_b200.training.gan import stylegan2 as T
    with CK.patched(exact_weights=exact_weights), precision.strict("full"):
        G, D = _product_models(states, fx["size"])
        c = fx["d_case"]
        x = c["x"].clone().requires_grad_(True)
        d, aux = D(x, projection=True, projection2=True, penultimate=True)
        _close(d, c["d"], 1e-4, "strict d"); _close(aux["projection"], c["projection"], 1e-4, "strict projection")
        ((d * c["c_d"]).sum() + (aux["projection"] * c["c1"]).sum() + (aux["projection2"] * c["c2"]).sum()).backward()
        _check_norms({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], 1e-3, "strict D")
        D.zero_grad()
        c = fx["r1_case"]
        per_sample = T.r1_per_sample(D, c["x"].clone(), lambda t: t)
        _close(per_sample, c["per_sample"], 1e-3, "strict R1 per sample")
        per_sample.mean().backward()
        _check_norms({k: p.grad for k, p in D.named_parameters()}, {k: n for k, n in c["grad_norms"].items() if n > 0},
                     2e-3, "strict R1")
        c = fx["g_case"]
        torch.manual_seed(c["mix_seed"])
        img = G(c["z"], style_mix=0.9, noise=c["noises"])
        _close(img, c["image"], 1e-4, "strict image")
        (img * c["c_img"]).sum().backward()
        # NoiseInjection weights are left out: d/d(noise weight) = sum(dpre * noise) is a sum of ~1e5 signed terms in which
        # single LeakyReLU sign flips show (a 1e-6 perturbation of one layer's output moves these norms by +-30 %), so they
        # are only comparable between bit-identical forward passes - the GPU tests exclude them for the same reason
        want = {k: n for k, n in c["grad_norms"].items() if not k.endswith("noise.weight")}
        _check_norms({k: p.grad for k, p in G.named_parameters()}, want, 1e-3, "strict G")


def test_ema_generator_is_not_served_stale_weights(fx, states):
    """the reference's `utils.accumulate` (utils.py:130-143) updates g_ema through
    `param.data.mul_().add_()`, which leaves `param._version` untouched; a no-grad forward of g_ema after such an
    update must see the new weights (FixedSampleGeneration / FID, evaluate/gan.py:57-58)."""
    import copy

    def ref_accumulate(model_dst, model_src, decay=0.999):          # utils.py:130-143, restated
        params_dst, params_src = dict(model_dst.named_parameters()), dict(model_src.named_parameters())
        for k in params_dst:
            params_dst[k].data.mul_(decay).add_(params_src[k].data, alpha=1 - decay)

    with CK.patched():
        G, _ = _product_models(states, fx["size"])
        g_ema = copy.deepcopy(G).eval()
        for p in g_ema.parameters():
            p.requires_grad_(False)
        c = fx["g_case"]
        with torch.no_grad():
            img0 = g_ema(c["z"], style_mix=0.0, noise=c["noises"]).clone()
            versions = [p._version for p in g_ema.parameters()]
            src = copy.deepcopy(G)
            for p in src.parameters():
                p.data.add_(0.05 * torch.randn_like(p))
            ref_accumulate(g_ema, src, decay=0.5)
            assert versions == [p._version for p in g_ema.parameters()]        # the update is invisible to torch
            img1 = g_ema(c["z"], style_mix=0.0, noise=c["noises"])
            fresh = copy.deepcopy(g_ema)
            want = fresh(c["z"], style_mix=0.0, noise=c["noises"])
        assert float((img1 - img0).abs().max()) > 1e-3, "g_ema output did not move after accumulate()"
        _close(img1, want, 1e-6, "g_ema after accumulate vs a fresh copy of the same weights")


def test_product_dstep_losses(fx, states):
    from contrad_b200.training.gan import stylegan2 as T
    import tests.cpu_loss_standins as LS
    with CK.patched(), LS.patched():
        G, D = _product_models(states, fx["size"])
        c = fx["dstep_case"]
        d_all, view_r, view_f = T.discriminate(D, c["real_aug2"], c["fake_aug"])
        P = type("P", (), {"temp": 0.1, "lbd_a": 1.0, "distributed": False})()
        d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
        assert abs(float(d_loss) - c["d_loss"]) < 1e-4 * abs(c["d_loss"])
        assert abs(float(aux["penalty"]) - c["penalty"]) < 1e-4 * abs(c["penalty"])
        assert abs(float(aux["d_real"]) - c["d_real"]) < 1e-4 * max(1.0, abs(c["d_real"]))
        assert abs(float(aux["d_gen"]) - c["d_gen"]) < 1e-4 * max(1.0, abs(c["d_gen"]))
        (d_loss + aux["penalty"]).backward()
        _check_norms({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], 2e-3, "product D-step")


def test_dstep16_oracle_and_product(fx, states):
    """The full D-step objective (contrastive + L_dis + 0.05 * R1) at n = 16: scalars, per-parameter gradient norms and
    the total gradient norm, oracle and product host logic against the reference."""
    from contrad_b200.training.gan import stylegan2 as T
    import tests.cpu_loss_standins as LS
    c = fx["dstep16_case"]
    real2, fake = c["real_aug2"].float(), c["fake_aug"].float()
    n = fake.shape[0]
    sd = _leafs(states[0])
    d_loss, penalty, _, _ = SO.gd_losses(sd, fx["size"], real2, fake)
    r1 = SO.r1_penalty(sd, real2[:n], fx["size"]).mean()
    assert abs(float(d_loss) - c["d_loss"]) < 1e-4 * abs(c["d_loss"]) and abs(float(r1) - c["r1"]) < 1e-3 * abs(c["r1"])
    (d_loss + penalty + 0.05 * r1).backward()
    _check_norms({k: v.grad for k, v in sd.items() if v.requires_grad}, c["grad_norms"], 2e-3, "oracle D-step n=16")
    with CK.patched(), LS.patched():
        G, D = _product_models(states, fx["size"])
        d_all, view_r, view_f = T.discriminate(D, real2, fake)
        P = type("P", (), {"temp": 0.1, "lbd_a": 1.0, "distributed": False})()
        d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
        r1 = T.r1_loss(D, real2[:n], lambda t: t)
        assert abs(float(d_loss) - c["d_loss"]) < 1e-4 * abs(c["d_loss"]) and abs(float(r1) - c["r1"]) < 1e-3 * abs(c["r1"])
        (d_loss + aux["penalty"] + 0.05 * r1).backward()
        _check_norms({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], 2e-3, "product D-step n=16")
        total = math.sqrt(sum(float(p.grad.double().pow(2).sum()) for p in D.parameters()))
        assert abs(total - c["total_grad_norm"]) < 1e-3 * c["total_grad_norm"]


# ------------------------------------------------------------------------------------------------ 3. real SIMT kernels on CPU
def test_product_stylegan2_on_emulated_kernels(fx, states):
    """The same fixtures with the REAL csrc/sg2_ops.cu kernels (upfirdn2d, patch gather / scatter, bias_act, modulation,
    minibatch stddev incl. its double backward, layout kernels ...) executed on the CPU by the CUDA emulator (tests/emu);
    only the tcgen05 entry points are torch stand-ins (tests/cpu_tc_standins.py, TF32 rounding of the outputs included).
    Tolerances are the TF32 bars of tests/test_gpu_sg2.py: D forward, the R1 double backward, G forward / backward."""
    import tests.cpu_tc_standins as TC
    from tests.emu import emulated
    from contrad_b200.training.gan import stylegan2 as T
    with emulated(), TC.patched():
        G, D = _product_models(states, fx["size"])
        c = fx["d_case"]
        x = c["x"].clone().requires_grad_(True)
        d, aux = D(x, projection=True, projection2=True, penultimate=True)
        _close(d, c["d"], 1e-2, "d"); _close(aux["projection"], c["projection"], 1e-2, "projection")
        _close(aux["penultimate"], c["penultimate"], 1e-2, "penultimate")
        ((d * c["c_d"]).sum() + (aux["projection"] * c["c1"]).sum() + (aux["projection2"] * c["c2"]).sum()).backward()
        l2 = float((x.grad - c["dx"]).norm() / c["dx"].norm())      # a few LeakyReLU kinks flip under TF32 rounding: L2 bar
        assert l2 < 5e-2, l2
        _check_norms({k: p.grad for k, p in D.named_parameters()}, c["grad_norms"], 2e-2, "product D (emulated kernels)")
        D.zero_grad()
        c = fx["r1_case"]
        per_sample = T.r1_per_sample(D, c["x"], lambda t: t)
        _close(per_sample, c["per_sample"], 2e-2, "r1 per sample")
        per_sample.mean().backward()
        total = math.sqrt(sum(float(p.grad.double().pow(2).sum()) for p in D.parameters() if p.grad is not None))
        want = math.sqrt(sum(n * n for n in c["grad_norms"].values()))
        assert abs(total - want) < 2e-2 * want, (total, want)
        c = fx["g_case"]
        torch.manual_seed(c["mix_seed"])
        img, latents = G(c["z"], return_latents=True, style_mix=0.9, noise=c["noises"])
        _close(latents, c["latents"], 5e-3, "latents"); _close(img, c["image"], 1e-2, "image")
        (img * c["c_img"]).sum().backward()
        total = math.sqrt(sum(float(p.grad.double().pow(2).sum()) for p in G.parameters() if p.grad is not None))
        want = math.sqrt(sum(n * n for n in c["grad_norms"].values()))
        assert abs(total - want) < 2e-2 * want, (total, want)


# ------------------------------------------------------------------ the stand-ins themselves against the reference's native ops
def _reference_function(rel_path, name):
    """Compile ONE pure-Python function of the unmodified reference without importing its module (importing
    models/gan/stylegan2/op JIT-builds two CUDA extensions).  Test infrastructure; skipped without the reference."""
    import ast
    from oracle import ref_import
    path = os.path.join(ref_import.REFERENCE_ROOT, rel_path)
    if not os.path.exists(path):
        pytest.skip("reference sources not available")
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def test_standins_equal_reference_native_ops():
    """The kernel-level GPU tests of csrc/sg2_ops.cu compare with tests/cpu_kernels.py; this ties those stand-ins to the
    reference's own CPU code: `upfirdn2d_native` (models/gan/stylegan2/op/upfirdn2d.py:159-200: zero-stuffing, signed
    padding, flipped-kernel correlation, decimation) for every (up, down, pad) the two networks use and a few they do
    not, and `fused_leaky_relu` (op/fused_act.py:86-92)."""
    native = _reference_function("models/gan/stylegan2/op/upfirdn2d.py", "upfirdn2d_native")
    flr = _reference_function("models/gan/stylegan2/op/fused_act.py", "fused_leaky_relu")
    torch.manual_seed(0)
    k1 = torch.tensor([1., 3., 3., 1.])
    k2 = k1[None] * k1[:, None]
    k2 = k2 / k2.sum()
    asym = torch.randn(3, 4)                                    # not symmetric: catches a missing kernel flip
    for H, W in ((9, 11), (8, 8)):
        x = torch.randn(2, 5, H, W)
        for fir in (k2, asym):
            for up, down, pad in ((1, 1, (2, 1)), (1, 1, (2, 2)), (1, 1, (1, 1)), (2, 1, (2, 1)), (1, 2, (1, 1)), (1, 2, (2, 2)),
                                  (2, 2, (0, 3)), (1, 1, (-1, 2)), (3, 2, (1, -1))):
                ref = native(x, fir, up, up, down, down, pad[0], pad[1], pad[0], pad[1])          # NCHW in, NCHW out
                got = CK.upfirdn2d(x.permute(0, 2, 3, 1).contiguous(), fir, up, down, (pad[0], pad[1], pad[0], pad[1]))
                assert got.shape[1:3] == ref.shape[2:], (up, down, pad)
                assert torch.allclose(got.permute(0, 3, 1, 2), ref, atol=1e-6, rtol=1e-6), (up, down, pad)
    x, bias = torch.randn(3, 7, 6, 6), torch.randn(7)
    ref = flr(x, bias, 0.2, 2 ** 0.5)
    got = CK.bias_act(x.permute(0, 2, 3, 1).contiguous(), bias, 0.2, 2 ** 0.5)
    assert torch.allclose(got.permute(0, 3, 1, 2), ref, atol=1e-6, rtol=1e-6)
