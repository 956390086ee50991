"""End-to-end parity of the B200 path (contrad_b200 modules + kernels) against the CPU oracle and the
reference-generated golden scalars.  Run with -m gpu."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import contrad_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    compat = os.path.join(repo, "contrad_b200", "compat")
    if compat not in sys.path:
        sys.path.append(compat)
    import gin
    gin.clear_config()
    gin.parse_config("""
ColorJitterLayer.brightness = 0.4
ColorJitterLayer.contrast = 0.4
ColorJitterLayer.saturation = 0.4
ColorJitterLayer.hue = 0.1
RandomResizeCropLayer.scale = (0.2, 1.0)
""")
    from contrad_b200.augment import get_augment
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import contrad
    from contrad_b200 import engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return SimpleNamespace(get_augment=get_augment, get_architecture=get_architecture, contrad=contrad, engine=engine)


def _unpack(packed):
    return {k: packed[i] for i, k in enumerate(O.PARAM_FIELDS)}


class _FixedAugment(torch.nn.Module):
    """Augment module fed with pre-drawn parameter blocks (so GPU and oracle see identical draws)."""

    def __init__(self, blocks):
        super().__init__()
        self.blocks = list(blocks)

    def forward(self, x):
        from contrad_b200.functional import AugmentSimCLRFn
        packed, order = self.blocks.pop(0)
        return AugmentSimCLRFn.apply(x, packed.to(x.device), order)


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


@pytest.mark.parametrize("loss_kind,n", [("nonsat", 4), ("hinge", 6)])
def test_d_and_g_losses_and_grads_vs_oracle(env, loss_kind, n):
    """Full-width SNDCGAN D (ndf=64) at a small batch: every loss, the D outputs and every parameter
    gradient against the fp32 CPU oracle on identical weights, latents and augmentation draws."""
    gen_w = torch.Generator().manual_seed(99)
    sd_d = O.make_d_state(generator=gen_w)
    sd_g = O.make_g_state(generator=gen_w)
    G, D = env.get_architecture("sndcgan", (32, 32, 3))
    D.load_state_dict(sd_d); G.load_state_dict(sd_g)
    G.cuda(); D.cuda(); G.train(); D.train()
    np.random.seed(3); torch.manual_seed(3)
    images = torch.rand(n, 3, 32, 32)
    z_d = O.sample_latent(n)
    aug_d = O.sample_simclr_params(3 * n, 32, 32)
    z_g = O.sample_latent(n)
    aug_g = O.sample_simclr_params(n, 32, 32)

    # ---------------- oracle (CPU fp32)
    sd_d_o = {k: v.clone() for k, v in sd_d.items()}
    sd_g_o = {k: v.clone() for k, v in sd_g.items()}
    O.set_requires_grad(sd_g_o, False); O.set_requires_grad(sd_d_o, True)
    with torch.no_grad():
        gen_o = O.g_sndcgan_forward(sd_g_o, z_d)
    l_con_o, l_dis_o, ex = O.loss_d(sd_d_o, images, gen_o, aug_d[0], aug_d[1], loss=loss_kind)
    (l_con_o + l_dis_o).backward()
    grads_o = {k: v.grad.clone() for k, v in O.trainable(sd_d_o).items()}
    uv_o = {k: v.clone() for k, v in sd_d_o.items() if k.endswith(("weight_u", "weight_v"))}

    # ---------------- B200 path
    P = SimpleNamespace(augment_fn=_FixedAugment([(O.pack_params(aug_d[0]), aug_d[1]), (O.pack_params(aug_g[0]), aug_g[1])]),
                        temp=0.1, lbd_a=1.0, distributed=False)
    options = {"loss": loss_kind}
    env.engine.set_grad(G, False); env.engine.set_grad(D, True)
    with torch.no_grad():
        gen = G(z_d.cuda())
    assert torch.allclose(gen.cpu(), gen_o, atol=2e-3)        # 4 TF32 layers + tanh; images are O(0.5)
    d_loss, aux = env.contrad.loss_D_fn(P, D, options, images.cuda(), gen)
    (d_loss + aux["penalty"]).backward()
    assert _rel(float(d_loss), float(l_con_o)) < 1e-3, (float(d_loss), float(l_con_o))
    assert _rel(float(aux["penalty"]), float(l_dis_o)) < 1e-3
    assert abs(float(aux["d_real"]) - float(ex["d_real"])) < 1e-3 * max(1.0, abs(float(ex["d_real"])))
    sd_now = D.state_dict()
    for k, ref in uv_o.items():
        assert torch.allclose(sd_now[k].cpu(), ref, atol=2e-5), k
    tot_o = torch.sqrt(sum(g.double().pow(2).sum() for g in grads_o.values()))
    tot = float(env.engine.grad_norm(D))
    # Gradient NORMS are the parity quantity (north_star: 1e-3 at the benchmark batch, checked on the b64 golden
    # scalars below).  At this tiny batch a handful of LeakyReLU pre-activations within TF32 rounding of zero take
    # the other slope than in the fp32 oracle (the reference's own cuDNN-TF32 GPU path does the same), which moves
    # the norm by a few 1e-3 and individual gradient tensors by a few percent - see DESIGN.md "parity".
    assert _rel(tot, float(tot_o)) < 5e-3, (tot, float(tot_o))
    named = dict(D.named_parameters())
    # calibrate the per-tensor tolerance on the reference stack itself: the oracle run on cuda with TF32 enabled
    torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
    sd_t = {k: v.clone().cuda() for k, v in sd_d.items()}
    O.set_requires_grad(sd_t, True)
    aug_c = {k: v.cuda() for k, v in aug_d[0].items()}
    l_con_t, l_dis_t, _ = O.loss_d(sd_t, images.cuda(), gen_o.cuda(), aug_c, aug_d[1], loss=loss_kind)
    (l_con_t + l_dis_t).backward()
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    errs, worst = {}, 0.0
    for k, g_o in grads_o.items():
        g = named[k].grad
        assert g is not None, k
        mine = float((g.cpu() - g_o).double().norm() / g_o.double().norm().clamp_min(1e-30))
        ref_tf32 = float((sd_t[k].grad.cpu() - g_o).double().norm() / g_o.double().norm().clamp_min(1e-30))
        errs[k] = (mine, ref_tf32)
        worst = max(worst, mine)
    print("per-parameter grad rel err (this repo, reference-on-cuda-TF32):", {k: ("%.1e" % a, "%.1e" % b) for k, (a, b) in errs.items()})
    for k, (mine, ref_tf32) in errs.items():
        if float(grads_o[k].double().norm()) < 1e-6 * float(tot_o):
            continue                      # exactly-cancelling gradients (e.g. hinge bias): relative error is meaningless
        assert mine < max(0.12, 4 * ref_tf32), (k, mine, ref_tf32)
    cos = sum(float((named[k].grad.cpu().double() * g_o.double()).sum()) for k, g_o in grads_o.items()) / (tot * float(tot_o))
    assert cos > 0.995, cos

    # ---------------- G step through the frozen D
    O.set_requires_grad(sd_g_o, True); O.set_requires_grad(sd_d_o, False)
    gen2_o = O.g_sndcgan_forward(sd_g_o, z_g)
    l_gen_o = O.loss_g(sd_d_o, gen2_o, aug_g[0], aug_g[1], loss=loss_kind)
    l_gen_o.backward()
    env.engine.set_grad(G, True); env.engine.set_grad(D, False)
    gen2 = G(z_g.cuda())
    g_loss = env.contrad.loss_G_fn(P, D, options, images.cuda(), gen2)
    g_loss.backward()
    assert _rel(float(g_loss), float(l_gen_o)) < 1e-3 or abs(float(g_loss) - float(l_gen_o)) < 1e-6
    gn_o = O.grad_norm(sd_g_o)
    # the generator gradient at initialisation is a near-cancelling sum pushed through D's TF32 data-gradient
    # chain; the reference's own GPU path (cuDNN TF32) deviates from its fp32 CPU path by the same order.
    torch.backends.cudnn.allow_tf32 = True
    sd_d_c = {k: v.detach().clone().cuda() for k, v in sd_d_o.items()}
    sd_g_c = {k: v.detach().clone().cuda() for k, v in sd_g.items()}
    for k, v in sd_d_c.items():            # undo the G-step power iteration so the cuda oracle repeats it
        if k.endswith(("weight_u", "weight_v")):
            v.copy_(uv_o[k].cuda())
    O.set_requires_grad(sd_g_c, True)
    aug_c = {k: v.cuda() for k, v in aug_g[0].items()}
    l_c = O.loss_g(sd_d_c, O.g_sndcgan_forward(sd_g_c, z_g.cuda()), aug_c, aug_g[1], loss=loss_kind)
    l_c.backward()
    gn_ref_tf32 = O.grad_norm(sd_g_c)
    torch.backends.cudnn.allow_tf32 = False
    mine = float(env.engine.grad_norm(G))
    print("G grad-norm: oracle fp32 CPU %.6e | oracle on cuda with cuDNN TF32 %.6e (dev %.2e) | this repo %.6e (dev %.2e)"
          % (gn_o, gn_ref_tf32, _rel(gn_ref_tf32, gn_o), mine, _rel(mine, gn_o)))
    assert _rel(mine, gn_o) < 3e-2, (mine, gn_o)       # tiny batch: ~1 % run-to-run (atomics order x kink flips)


@pytest.mark.parametrize("n", [8, 64])
def test_generator_forward_backward_vs_oracle(env, n):
    """G_SNDCGAN on the tcgen05 / SIMT kernels vs the fp32 oracle: images, BN running statistics, and every
    parameter gradient for a random upstream gradient."""
    gen_w = torch.Generator().manual_seed(7)
    sd_g = O.make_g_state(generator=gen_w)
    # non-trivial BN affine parameters
    for k in sd_g:
        if k.endswith(".weight") and sd_g[k].dim() == 1:
            sd_g[k] = 1.0 + 0.1 * torch.randn(sd_g[k].shape, generator=gen_w)
        if k.endswith(".bias"):
            sd_g[k] = 0.05 * torch.randn(sd_g[k].shape, generator=gen_w)
    G, _ = env.get_architecture("sndcgan", (32, 32, 3))
    G.load_state_dict(sd_g); G.cuda().train()
    sd_o = {k: v.clone() for k, v in sd_g.items()}
    O.set_requires_grad(sd_o, True)
    torch.manual_seed(n)
    z = O.sample_latent(n)
    dout = torch.randn(n, 3, 32, 32)
    out_o = O.g_sndcgan_forward(sd_o, z)
    (out_o * dout).sum().backward()
    out = G(z.cuda())
    (out * dout.cuda()).sum().backward()
    assert torch.allclose(out.cpu(), out_o.detach(), atol=2e-3), (out.cpu() - out_o).abs().max()   # TF32, O(0.5) values
    sd_now = G.state_dict()
    for k in sd_o:
        if "running" in k:
            assert torch.allclose(sd_now[k].cpu(), sd_o[k], atol=1e-5, rtol=1e-3), k
    assert int(sd_now["norm_init.num_batches_tracked"]) == 1
    named = dict(G.named_parameters())
    errs = {}
    for k, v in O.trainable(sd_o).items():
        g = named[k].grad
        assert g is not None, k
        errs[k] = float((g.cpu() - v.grad).double().norm() / v.grad.double().norm().clamp_min(1e-30))
    print("G grad rel errs:", {k: "%.1e" % e for k, e in errs.items()})
    gn_o = O.grad_norm(sd_o)
    for k, e in errs.items():
        if float(O.trainable(sd_o)[k].grad.double().norm()) < 1e-4 * gn_o:
            continue                  # biases in front of a BatchNorm: the exact gradient is zero
        assert e < 8e-2, (k, e)       # ReLU kink flips under TF32; norms below are tight
    gn = float(env.engine.grad_norm(G))
    assert _rel(gn, gn_o) < 5e-3, (gn, gn_o)


@pytest.mark.parametrize("strict", [False, "full"])
def test_config1_two_steps_vs_reference_scalars(env, golden_dir, strict):
    """BASELINE config 1 shape (b64, c10_b512.gin hyper-parameters): two complete train steps (Adam incl.) on
    the B200 path reproduce the UNMODIFIED reference's scalars (tests/golden/config1_scalars.json) to 1e-3.

    strict = "full": the strict precision mode (contrad_b200/precision.py, error-compensated "3xTF32" operands in both
    steps; the generator step of iteration 2 runs through discriminator weights that two Adam updates produced, so the
    discriminator step has to be compensated as well) - EVERY scalar incl. the generator's gradient norm within
    north_star's 1e-3 at both steps.  strict = False (the
    default single-pass TF32): losses and the discriminator's gradient norm within 1e-3; the generator's gradient norm at
    initialisation is a small residual that any single TF32 rounding moves by 1e-3 .. 1e-2 (tools/tf32_sensitivity.py,
    profiles/tf32_sensitivity_r2.json) - bounded at 2e-2 here and stated as such in DESIGN.md."""
    from contrad_b200 import precision
    precision.set_strict(strict)
    try:
        _config1_two_steps(env, golden_dir, strict)
    finally:
        precision.set_strict(False)


def _config1_two_steps(env, golden_dir, strict):
    with open(os.path.join(golden_dir, "config1_scalars.json")) as f:
        fx = json.load(f)
    n = fx["batch"]
    gen_w = torch.Generator().manual_seed(fx["weights_seed"])
    sd_d = O.make_d_state(generator=gen_w)
    sd_g = O.make_g_state(generator=gen_w)
    G, D = env.get_architecture("sndcgan", (32, 32, 3))
    D.load_state_dict(sd_d); G.load_state_dict(sd_g)
    G.cuda(); D.cuda()
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    options = {"loss": "nonsat", "warmup": 3000, "lr": 2e-4, "lr_d": 2e-4}
    np.random.seed(fx["data_seed"]); torch.manual_seed(fx["data_seed"])
    train_fn = {"D": env.contrad.loss_D_fn, "G": env.contrad.loss_G_fn}

    class _G(torch.nn.Module):       # latents drawn on the CPU generator in the reference's order
        def __init__(self, g, zs):
            super().__init__(); self.g, self.zs = g, zs
        def sample_latent(self, k):
            return self.zs.pop(0).cuda()
        def forward(self, z):
            return self.g(z)
        def parameters(self, recurse=True):
            return self.g.parameters(recurse)
        def train(self, mode=True):
            self.g.train(mode); return self

    for ref in fx["steps"]:
        images = torch.rand(n, 3, 32, 32)
        z_d = O.sample_latent(n)
        aug_d = O.sample_simclr_params(3 * n, 32, 32)
        z_g = O.sample_latent(n)
        aug_g = O.sample_simclr_params(n, 32, 32)
        P = SimpleNamespace(augment_fn=_FixedAugment([(O.pack_params(aug_d[0]), aug_d[1]),
                                                      (O.pack_params(aug_g[0]), aug_g[1])]),
                            temp=0.1, lbd_a=1.0, distributed=False)
        got = env.engine.train_step(P, options, train_fn, (_G(G, [z_d, z_g]), D), (opt_G, opt_D), images.cuda(),
                                    ref["step"], record_grad_norms=True)
        print("step %d:" % ref["step"], {k: (float(got[m]), ref[k]) for k, m in
              (("l_con", "d_loss"), ("l_dis", "d_penalty"), ("l_gen", "g_loss"), ("d_grad_norm", "d_grad_norm"),
               ("g_grad_norm", "g_grad_norm"))})
        for key, mine in (("l_con", "d_loss"), ("l_dis", "d_penalty"), ("l_gen", "g_loss"),
                          ("d_grad_norm", "d_grad_norm")):
            assert _rel(float(got[mine]), ref[key]) < 1e-3, (ref["step"], key, float(got[mine]), ref[key])
        g_tol = 1e-3 if strict else 2e-2
        assert _rel(float(got["g_grad_norm"]), ref["g_grad_norm"]) < g_tol, (ref["step"], strict, float(got["g_grad_norm"]),
                                                                            ref["g_grad_norm"])


def test_full_batch_step_runs_and_is_finite(env):
    """BASELINE config 2 shape (N=512, D-step batch 1536): one step; losses finite and at their
    initialisation values (L_dis ~ 2 ln 2, L_gen ~ ln 2, L_con+ ~ ln(2N-1) scale)."""
    torch.manual_seed(0); np.random.seed(0)
    G, D = env.get_architecture("sndcgan", (32, 32, 3))
    G.cuda(); D.cuda()
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    options = {"loss": "nonsat", "warmup": 3000, "lr": 2e-4}
    P = SimpleNamespace(augment_fn=env.get_augment("simclr").cuda(), temp=0.1, lbd_a=1.0, distributed=False)
    train_fn = {"D": env.contrad.loss_D_fn, "G": env.contrad.loss_G_fn}
    images = torch.rand(512, 3, 32, 32, device="cuda")
    out = env.engine.train_step(P, options, train_fn, (G, D), (opt_G, opt_D), images, 1)
    vals = {k: float(v) for k, v in out.items()}
    assert all(np.isfinite(v) for v in vals.values()), vals
    assert abs(vals["d_penalty"] - 2 * np.log(2)) < 0.05 and abs(vals["g_loss"] - np.log(2)) < 0.05, vals
    assert 8.0 < vals["d_loss"] < 16.0, vals


class _HostDrawnAugment(torch.nn.Module):
    """SimCLR parameters drawn entirely on the host (the oracle's sampler, numpy RNG) and staged to the device:
    deterministic for a seed in BOTH the eager loop and the CUDA-graph loop, and exercises the per-image jitter
    order (kernel order = -1, row 11)."""

    def forward(self, x):
        from contrad_b200 import staging
        from contrad_b200.functional import AugmentSimCLRFn
        n = x.shape[0]

        def draw():
            params, order = O.sample_simclr_params(n, 32, 32)
            return torch.cat([O.pack_params(params), torch.full((1, n), float(order))])

        return AugmentSimCLRFn.apply(x, staging.stage(draw, x.device, shape=(12, n)), -1)


def test_cuda_graph_step_matches_eager_step(env):
    """engine.GraphedTrainStep (3 eager steps, capture, replays) against the plain eager loop: same seeds -> same
    host draws -> the same losses, the same Adam step counters and the same weight UPDATES after 7 steps.
    The learning rate is tiny on purpose: early Adam steps are sign-like (m/sqrt(v) = +-1), so fp32-atomic
    reordering of near-zero gradients flips O(lr) updates and, at the reference lr, two EAGER runs already differ by
    0.3 % in the loss after four steps."""
    from contrad_b200.optim import FusedAdam
    n, steps, lr = 32, 7, 2e-6
    options = {"loss": "nonsat", "warmup": 5, "lr": lr, "lr_d": lr}
    train_fn = {"D": env.contrad.loss_D_fn, "G": env.contrad.loss_G_fn}
    runs = []
    for graphed in (False, True):
        gen_w = torch.Generator().manual_seed(77)
        G, D = env.get_architecture("sndcgan", (32, 32, 3))
        D.load_state_dict(O.make_d_state(generator=gen_w)); G.load_state_dict(O.make_g_state(generator=gen_w))
        G.cuda(); D.cuda()
        params = [p for p in D.parameters()] + [p for p in G.parameters()]
        w0 = [p.detach().clone() for p in params]
        opt_G = FusedAdam(G.parameters(), lr=lr, betas=(0.5, 0.999))
        opt_D = FusedAdam(D.parameters(), lr=lr, betas=(0.5, 0.999))
        P = SimpleNamespace(augment_fn=_HostDrawnAugment(), temp=0.1, lbd_a=1.0, distributed=False)
        np.random.seed(21); torch.manual_seed(21)
        data = torch.rand(steps, n, 3, 32, 32).cuda()
        step_fn = (env.engine.GraphedTrainStep(P, options, train_fn, (G, D), (opt_G, opt_D)) if graphed else
                   (lambda im, s: env.engine.train_step(P, options, train_fn, (G, D), (opt_G, opt_D), im, s)))
        log = []
        for s in range(steps):
            out = step_fn(data[s], s + 1)
            log.append({k: float(v) for k, v in out.items()})
        torch.cuda.synchronize()
        if graphed:
            assert step_fn.graph is not None and step_fn.launches_per_replay > 100
        runs.append((log, [p.detach() - a for p, a in zip(params, w0)],
                     [float(opt_D.state[p]["step"]) for p in D.parameters()]))
    (log_e, d_e, t_e), (log_g, d_g, t_g) = runs
    assert t_e == t_g == [float(steps)] * len(t_e)
    for s, (a, b) in enumerate(zip(log_e, log_g)):
        for k in a:
            assert _rel(b[k], a[k]) < 1e-3 or abs(a[k] - b[k]) < 1e-4, (s, k, a[k], b[k])
    num = torch.stack([(a - b).norm() for a, b in zip(d_e, d_g)]).norm()
    den = torch.stack([a.norm() for a in d_e]).norm()
    print("relative difference of the 7-step weight updates, graph vs eager: %.3g" % float(num / den))
    assert float(num / den) < 5e-2
    # every tensor moved, and moved alike - except the four G biases that feed a BatchNorm (exactly zero true
    # gradient: their Adam updates are rounding noise in both runs)
    odd = [i for i, (a, b) in enumerate(zip(d_e, d_g))
           if not (float(a.norm()) > 0 and float((a - b).norm()) < 0.25 * float(a.norm()) + 1e-9)]
    assert len(odd) <= 4, odd




def test_snresnet18_matches_reference(env, golden_dir):
    """D_SNResNet18 against the fixture produced by the unmodified reference module (CPU fp32):
    outputs to 1e-2 of their max (TF32), input gradient / first-layer weight gradient in L2, per-parameter gradient
    norms, and the in-place power-iteration update of u / v."""
    fx = torch.load(os.path.join(golden_dir, "snresnet18.pt"), weights_only=False)
    G, D = env.get_architecture("snresnet18", (32, 32, 3))
    sd = O.make_d_resnet18_state(generator=torch.Generator().manual_seed(fx["w_seed"]))
    D.load_state_dict(sd, strict=True)
    D.cuda().train()
    x = fx["x"].cuda().requires_grad_(True)
    d, aux = D(x, projection=True, projection2=True, penultimate=True)
    ((d * fx["c_d"].cuda()).sum() + (aux["projection"] * fx["c1"].cuda()).sum() + (aux["projection2"] * fx["c2"].cuda()).sum()).backward()
    rel = lambda a, b: float((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max())
    l2 = lambda a, b: float((a.detach().cpu().double() - b.double()).norm() / b.double().norm())
    errs = {"d": rel(d, fx["d"]), "projection": rel(aux["projection"], fx["projection"]),
            "penultimate": rel(aux["penultimate"], fx["penultimate"]), "dx_l2": l2(x.grad, fx["dx"]),
            "grad_conv1_l2": l2(D.conv1.weight_orig.grad, fx["grad_conv1"])}
    grads = dict(D.named_parameters())
    errs["worst_norm"] = max(abs(float(grads[k].grad.norm()) - n) / n for k, n in fx["grad_norms"].items() if n > 0)
    for k, v in fx["uv_after"].items():
        errs["uv:" + k] = rel(D.state_dict()[k], v)
    # yardstick for the ill-conditioned element-wise gradients (17 LeakyReLU(0.1) layers: a pre-activation within TF32
    # rounding of zero flips its slope by a factor 10): the same fixture through torch's own TF32 GPU arithmetic
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        sd_c = {k: v.clone().cuda() for k, v in O.make_d_resnet18_state(generator=torch.Generator().manual_seed(fx["w_seed"])).items()}
        O.set_requires_grad(sd_c, True)
        xc = fx["x"].cuda().requires_grad_(True)
        dc, auxc = O.d_snresnet18_forward(sd_c, xc)
        ((dc * fx["c_d"].cuda()).sum() + (auxc["projection"] * fx["c1"].cuda()).sum() + (auxc["projection2"] * fx["c2"].cuda()).sum()).backward()
        errs["torch_tf32:dx_l2"] = l2(xc.grad, fx["dx"])
        errs["torch_tf32:grad_conv1_l2"] = l2(sd_c["conv1.weight_orig"].grad, fx["grad_conv1"])
        errs["torch_tf32:d"] = rel(dc, fx["d"])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    print("snresnet18 parity:", {k: "%.2e" % v for k, v in errs.items()})
    assert errs["d"] < 1e-2 and errs["projection"] < 1e-2 and errs["penultimate"] < 1e-2
    assert errs["dx_l2"] < max(5e-2, 3 * errs["torch_tf32:dx_l2"]) + 5e-2
    assert errs["grad_conv1_l2"] < max(5e-2, 3 * errs["torch_tf32:grad_conv1_l2"]) + 5e-2 and errs["worst_norm"] < 5e-2
    assert all(v < 1e-4 for k, v in errs.items() if k.startswith("uv:"))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", "snresnet18_parity.json"), "w") as f:
            json.dump(errs, f, indent=1)
    except OSError:
        pass


def test_snresnet18_train_step_runs(env):
    """`train_gan.py ... snresnet18 --mode=contrad --aug=simclr`: one full step with D_SNResNet18 + G_SNDCGAN."""
    from contrad_b200.optim import FusedAdam
    np.random.seed(3); torch.manual_seed(3)
    G, D = env.get_architecture("snresnet18", (32, 32, 3))
    G.cuda(); D.cuda()
    P = SimpleNamespace(augment_fn=env.get_augment(mode="simclr").cuda(), temp=0.1, lbd_a=1.0, distributed=False)
    options = {"loss": "hinge", "warmup": 3000, "lr": 2e-4, "lr_d": 2e-4}
    train_fn = {"D": env.contrad.loss_D_fn, "G": env.contrad.loss_G_fn}
    opt_G = FusedAdam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    opt_D = FusedAdam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    w0 = {k: v.detach().clone() for k, v in D.named_parameters()}
    for step in (1, 2):
        out = env.engine.train_step(P, options, train_fn, (G, D), (opt_G, opt_D), torch.rand(32, 3, 32, 32, device="cuda"), step,
                                    record_grad_norms=True)
    torch.cuda.synchronize()
    assert all(torch.isfinite(v).all() for v in out.values()), out
    moved = sum(1 for k, v in D.named_parameters() if not torch.equal(v.detach(), w0[k]))
    # hinge loss at initialisation: every d_real / d_gen is inside the margin, so the gradient of `linear.l2.bias`
    # (+1/n per fake, -1/n per real) cancels exactly and Adam leaves that single scalar untouched
    assert moved >= len(w0) - 1, (moved, len(w0))


def test_zz_eager_gpu_yardstick_sndcgan(env):
    """Not a parity check: the denominator of north_star's ">= 10x the reference single-GPU PyTorch-eager images/sec".
    The reference cannot travel to the GPU box, so its per-step arithmetic is executed through the oracle's torch ops ON
    THE GPU (cuDNN / cuBLAS with TF32 allowed, torch.optim.Adam, the reference's 5 `.item()` reads per step,
    host-drawn augmentation parameters): BASELINE config 2, b512.  Written to gpurun_out/eager_gpu_sndcgan.json."""
    import json
    n, steps, warm = 512, 6, 3
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        np.random.seed(0); torch.manual_seed(0)
        sd_d = {k: v.cuda() for k, v in O.make_d_state().items()}
        sd_g = {k: v.cuda() for k, v in O.make_g_state().items()}
        for sd in (sd_d, sd_g):
            O.set_requires_grad(sd, True)
        opt_d = torch.optim.Adam(list(O.trainable(sd_d).values()), lr=2e-4, betas=(0.5, 0.999))
        opt_g = torch.optim.Adam(list(O.trainable(sd_g).values()), lr=2e-4, betas=(0.5, 0.999))
        images = torch.rand(n, 3, 32, 32, device="cuda")

        def step():
            O.set_requires_grad(sd_g, False); O.set_requires_grad(sd_d, True)
            with torch.no_grad():
                gen = O.g_sndcgan_forward(sd_g, O.sample_latent(n).cuda())
            p, order = O.sample_simclr_params(3 * n, 32, 32, device="cuda")
            l_con, l_dis, ex = O.loss_d(sd_d, images, gen, p, order)
            opt_d.zero_grad()
            (l_con + l_dis).backward()
            opt_d.step()
            reads = [l_con.item(), l_dis.item(), ex["d_real"].item(), ex["d_gen"].item()]
            O.set_requires_grad(sd_g, True); O.set_requires_grad(sd_d, False)
            gen = O.g_sndcgan_forward(sd_g, O.sample_latent(n).cuda())
            p, order = O.sample_simclr_params(n, 32, 32, device="cuda")
            l_gen = O.loss_g(sd_d, gen, p, order)
            opt_g.zero_grad()
            l_gen.backward()
            opt_g.step()
            reads.append(l_gen.item())
            return reads

        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            reads = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert all(np.isfinite(r) for r in reads)
    res = {"workload": "SNDCGAN+ContraD b512 32x32, oracle torch ops on the GPU (cuDNN/cuBLAS TF32, torch.optim.Adam)",
           "ms_per_step": ms, "images_per_s": n / ms * 1e3}
    print(res)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", "eager_gpu_sndcgan.json"), "w") as f:
            json.dump(res, f)
    except OSError:
        pass
