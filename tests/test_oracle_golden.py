"""Pin the CPU oracle (oracle/contrad_oracle.py) against the fixtures produced by running the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import contrad_oracle as O


def _unpack(packed):
    return {k: packed[i] for i, k in enumerate(O.PARAM_FIELDS)}


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_augment_forward_backward_matches_reference(golden_dir):
    fx = _load(golden_dir, "augment_simclr.pt")
    assert fx["fields"] == list(O.PARAM_FIELDS)
    orders = set()
    for case in fx["cases"]:
        x = case["x"].clone().requires_grad_(True)
        y = O.augment_simclr(x, _unpack(case["params"]), case["order"])
        (y * case["dy"]).sum().backward()
        orders.add(case["order"])
        assert torch.allclose(y, case["y"], atol=5e-6, rtol=0), (y - case["y"]).abs().max()
        assert torch.allclose(x.grad, case["dx"], atol=2e-5, rtol=1e-5), (x.grad - case["dx"]).abs().max()
    assert orders == {0, 1}


def test_sampler_replays_reference_rng_stream(golden_dir):
    """Same seeds -> same explicit draws as stored (and those reproduced the reference output)."""
    fx = _load(golden_dir, "augment_simclr.pt")
    for case in fx["cases"]:
        b, _, h, w = case["x"].shape
        np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
        x = torch.rand(b, 3, h, w); _ = torch.randn(b, 3, h, w)
        assert torch.equal(x, case["x"])
        params, order = O.sample_simclr_params(b, h, w)
        assert order == case["order"]
        assert torch.equal(O.pack_params(params), case["params"])


def test_contrastive_losses_match_reference(golden_dir):
    fx = _load(golden_dir, "contrastive.pt")
    for case in fx["cases"]:
        a, b, c = (case[k].clone().requires_grad_(True) for k in "abc")
        l1 = O.nt_xent(a, b, 0.1)
        g1 = torch.autograd.grad(l1, [a, b])
        assert abs(float(l1) - case["nt_xent"]) < 1e-5 * max(1, abs(case["nt_xent"]))
        for g, r in zip(g1, case["nt_xent_grads"]):
            assert torch.allclose(g, r, atol=1e-6, rtol=1e-4)
        l2 = O.supcon_fake(a, b, c, 0.1)
        g2 = torch.autograd.grad(l2, [a, b, c])
        assert abs(float(l2) - case["supcon"]) < 1e-5 * max(1, abs(case["supcon"]))
        for g, r in zip(g2, case["supcon_grads"]):
            assert torch.allclose(g, r, atol=1e-6, rtol=1e-4)
        assert abs(float(O.nt_xent(a, b, 0.5)) - case["nt_xent_t05"]) < 1e-5


@pytest.mark.parametrize("n", [1, 2, 33])
def test_contrastive_oracle_equals_reference_on_degenerate_batches(n):
    """The oracle against the UNMODIFIED reference functions (training/criterion.py:24-45, training/gan/contrad.py:8-32;
    oracle/_ref or /root/reference, skipped when neither exists) where the fixtures do not reach: a single pair, duplicated
    and antipodal embeddings, tau = 0.01.  supcon_fake with one fake row is NaN in the reference (its mask row sums to
    zero) - the oracle must say the same.  tests/test_emulated_kernels.py holds the kernels to the oracle on these inputs."""
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference sources not available")
    ref_import.activate()
    try:
        from training.criterion import nt_xent as ref_nt_xent
        from training.gan.contrad import supcon_fake as ref_supcon
        torch.manual_seed(100 + n)
        base = F.normalize(torch.randn(n, 128))
        cases = [tuple(F.normalize(torch.randn(n, 128)) for _ in range(3)),
                 (base[:1].expand(n, 128).contiguous(),) * 3,
                 (base, base.clone(), F.normalize(torch.randn(n, 128))),
                 (base, -base, F.normalize(torch.randn(n, 128)))]
        for a, b, c in cases:
            for temp in (0.1, 0.01):
                r1, o1 = ref_nt_xent(a, b, temperature=temp), O.nt_xent(a, b, temp)
                assert torch.allclose(o1, r1, atol=1e-6, rtol=1e-5), (n, temp, float(o1), float(r1))
                r2, o2 = ref_supcon(a, b, c, temp), O.supcon_fake(a, b, c, temp)
                if n == 1:
                    assert torch.isnan(r2) and torch.isnan(o2)
                else:
                    assert torch.allclose(o2, r2, atol=1e-6, rtol=1e-5), (n, temp, float(o2), float(r2))
    finally:
        ref_import.deactivate()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_augment_oracle_equals_reference_on_degenerate_images(seed):
    """The oracle's SimCLR chain against the UNMODIFIED reference's `get_augment('simclr')` (augment/__init__.py:95-112,
    skipped when the reference is absent) on images the random fixtures never contain (tests/edge_inputs.py).  The
    oracle's sampler replays the reference's draws after the same seed (as in tests/golden/make_golden.py)."""
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference sources not available")
    from tests.edge_inputs import degenerate_images
    x = degenerate_images()
    B, size = x.shape[0], x.shape[-1]
    ref_import.activate()
    try:
        from augment import get_augment
        aug = get_augment(mode="simclr")
        np.random.seed(seed); torch.manual_seed(seed)
        params, order = O.sample_simclr_params(B, size, size)
        np.random.seed(seed); torch.manual_seed(seed)
        xr = x.clone().requires_grad_(True)
        y_ref = aug(xr)
        xo = x.clone().requires_grad_(True)
        y = O.augment_simclr(xo, params, order)
        assert torch.allclose(y, y_ref, atol=5e-6, rtol=0), float((y - y_ref).abs().max())
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed))
        (y_ref * dy).sum().backward(); (y * dy).sum().backward()
        assert torch.allclose(xo.grad, xr.grad, atol=1e-4, rtol=1e-4), float((xo.grad - xr.grad).abs().max())
    finally:
        ref_import.deactivate()


def test_spectral_norm_matches_torch_hook(golden_dir):
    fx = _load(golden_dir, "spectral_norm.pt")
    for name, rec in fx.items():
        w = rec["weight_orig"].clone().requires_grad_(True)
        u, v = rec["u0"].clone(), rec["v0"].clone()
        w_hat = O.spectral_normalize(w, u, v, training=True)
        assert torch.allclose(u, rec["u1"], atol=1e-6) and torch.allclose(v, rec["v1"], atol=1e-6)
        assert torch.allclose(w_hat, rec["w_hat"], atol=1e-6, rtol=1e-5)
        if name == "conv":
            y = F.conv2d(rec["x"], w_hat, rec["bias"], padding=1)
        else:
            y = F.linear(rec["x"], w_hat, rec["bias"])
        assert torch.allclose(y, rec["y"], atol=1e-5, rtol=1e-5)
        y.pow(2).sum().backward()
        assert torch.allclose(w.grad, rec["grad_weight_orig"], atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("loss_kind", ["nonsat", "hinge"])
def test_small_sndcgan_step_matches_reference(golden_dir, loss_kind):
    fx = _load(golden_dir, "sndcgan_small.pt")[loss_kind]
    sd_d = {k: v.clone() for k, v in fx["sd_d"].items()}
    sd_g = {k: v.clone() for k, v in fx["sd_g"].items()}
    n = fx["images"].shape[0]
    O.set_requires_grad(sd_g, False); O.set_requires_grad(sd_d, True)
    with torch.no_grad():
        gen = O.g_sndcgan_forward(sd_g, fx["z_d"], ngf=4)
    assert torch.allclose(gen, fx["d_step"]["gen"], atol=1e-5)
    l_con, l_dis, ex = O.loss_d(sd_d, fx["images"], gen, _unpack(fx["aug_d"]), fx["order_d"], loss=loss_kind)
    (l_con + l_dis).backward()
    ref = fx["d_step"]
    assert abs(float(l_con) - ref["l_con"]) < 1e-4 * abs(ref["l_con"])
    assert abs(float(l_dis) - ref["l_dis"]) < 1e-4 * abs(ref["l_dis"])
    assert abs(float(ex["d_real"]) - ref["d_real"]) < 1e-5 + 1e-3 * abs(ref["d_real"])
    assert abs(float(ex["d_gen"]) - ref["d_gen"]) < 1e-5 + 1e-3 * abs(ref["d_gen"])
    for k, gn in ref["grad_norms"].items():
        mine = float(sd_d[k].grad.double().norm()) if sd_d[k].grad is not None else 0.0
        assert abs(mine - gn) <= 1e-3 * gn + 1e-7, (k, mine, gn)
    for k, buf in fx["uv_after_d_step"].items():
        assert torch.allclose(sd_d[k], buf, atol=1e-5), k
    # G step through the frozen D (second power iteration happens here as in the reference)
    O.set_requires_grad(sd_g, True); O.set_requires_grad(sd_d, False)
    for v in sd_d.values():
        v.grad = None
    gen2 = O.g_sndcgan_forward(sd_g, fx["z_g"], ngf=4)
    l_gen = O.loss_g(sd_d, gen2, _unpack(fx["aug_g"]), fx["order_g"], loss=loss_kind)
    l_gen.backward()
    assert abs(float(l_gen) - fx["g_step"]["l_gen"]) < 1e-4 * abs(fx["g_step"]["l_gen"]) + 1e-7
    for k, gn in fx["g_step"]["grad_norms"].items():
        mine = float(sd_g[k].grad.double().norm())
        assert abs(mine - gn) <= 2e-3 * gn + 1e-8, (k, mine, gn)


def test_config1_two_full_steps_match_reference(golden_dir):
    """BASELINE config 1 (SNDCGAN+ContraD, c10_b512.gin hyper-parameters, b64, CPU): two complete
    train steps incl. Adam agree with the reference's scalars within the north_star tolerance (1e-3)."""
    with open(os.path.join(golden_dir, "config1_scalars.json")) as f:
        fx = json.load(f)
    n = fx["batch"]
    gen_w = torch.Generator().manual_seed(fx["weights_seed"])
    sd_d = O.make_d_state(generator=gen_w)
    sd_g = O.make_g_state(generator=gen_w)
    opt_g = O.Adam(O.trainable(sd_g).values(), 2e-4)
    opt_d = O.Adam(O.trainable(sd_d).values(), 2e-4)
    np.random.seed(fx["data_seed"]); torch.manual_seed(fx["data_seed"])
    for ref in fx["steps"]:
        images = torch.rand(n, 3, 32, 32)
        z_d = O.sample_latent(n)
        aug_d = O.sample_simclr_params(3 * n, 32, 32)
        z_g = O.sample_latent(n)
        aug_g = O.sample_simclr_params(n, 32, 32)
        got = O.train_step(sd_g, sd_d, opt_g, opt_d, images, z_d, z_g, aug_d, aug_g, step=ref["step"])
        l_con = got["l_con_pos"] + got["l_con_neg"]
        assert abs(l_con - ref["l_con"]) < 1e-3 * abs(ref["l_con"])
        for key in ("l_dis", "l_gen", "d_grad_norm", "g_grad_norm"):
            assert abs(got[key] - ref[key]) < 1e-3 * abs(ref[key]), (key, got[key], ref[key])


def test_oracle_augment_hq_matches_reference(golden_dir):
    """simclr_hq / simclr_hq_cutout: the oracle's sampler + arithmetic reproduce the reference chain (forward 5e-6,
    backward through autograd 1e-5)."""
    fx = torch.load(os.path.join(golden_dir, "augment_hq.pt"), weights_only=False)
    for c in fx["cases"]:
        np.random.seed(c["seed"]); torch.manual_seed(c["seed"])
        b, _, h, w = c["x"].shape
        x = torch.rand(b, 3, h, w); _ = torch.randn(b, 3, h, w)
        params, order = O.sample_simclr_params(b, h, w)
        hq = O.sample_hq_params(b, h, w, cutout=c["mode"].endswith("cutout"))
        assert torch.equal(x, c["x"]) and order == c["order"] and torch.equal(O.pack_params(params), c["params"])
        assert hq["sigma"] == c["hq"]["sigma"] and torch.equal(hq["blur_on"], c["hq"]["blur_on"])
        xr = x.clone().requires_grad_(True)
        y = O.augment_simclr_hq(xr, params, order, hq, cutout_length=c["length"])
        assert (y.detach() - c["y"]).abs().max() < 5e-6
        (y * c["dy"]).sum().backward()
        assert (xr.grad - c["dx"]).abs().max() < 1e-5 * max(1.0, float(c["dx"].abs().max()))


def test_oracle_snresnet18_matches_reference(golden_dir):
    """D_SNResNet18 restatement vs the unmodified reference module (forward, input gradient, gradient norms, u/v)."""
    fx = torch.load(os.path.join(golden_dir, "snresnet18.pt"), weights_only=False)
    sd = O.make_d_resnet18_state(generator=torch.Generator().manual_seed(fx["w_seed"]))
    assert {k: list(v.shape) for k, v in sd.items()} == fx["keys"]
    O.set_requires_grad(sd, True)
    x = fx["x"].clone().requires_grad_(True)
    d, aux = O.d_snresnet18_forward(sd, x)
    assert (d - fx["d"]).abs().max() < 1e-6 and (aux["penultimate"] - fx["penultimate"]).abs().max() < 1e-6
    ((d * fx["c_d"]).sum() + (aux["projection"] * fx["c1"]).sum() + (aux["projection2"] * fx["c2"]).sum()).backward()
    assert (x.grad - fx["dx"]).abs().max() < 1e-6 * max(1.0, float(fx["dx"].abs().max()))
    for k, n in fx["grad_norms"].items():
        assert abs(float(sd[k].grad.norm()) - n) <= 1e-5 * max(n, 1e-6), k
    for k, v in fx["uv_after"].items():
        assert (sd[k] - v).abs().max() < 1e-6


# ---- rows f3 / f4: uint8 input, hfrt / RandomCrop, Gaussian noise, baseline training modes --------------------

def test_shift_flip_matches_reference(golden_dir):
    """HorizontalFlipRandomCrop / RandomCrop through the reference's affine_grid + nearest grid_sample vs the oracle's
    index-gather restatement: bit-exact forward, backward to fp32 summation order."""
    fx = _load(golden_dir, "augment_aux.pt")
    pads = set()
    for case in fx["shift_flip"]:
        b = case["x"].shape[0]
        np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
        _ = torch.rand_like(case["x"]); _ = torch.randn_like(case["x"])
        params = O.sample_shift_flip(b, case["max_pixels"], case["width"], flip=(case["kind"] == "hfrt"))
        assert torch.equal(params, case["params"])
        x = case["x"].clone().requires_grad_(True)
        y = O.shift_flip(x, case["params"], case["padding_mode"])
        (y * case["dy"]).sum().backward()
        assert torch.equal(y, case["y"]), (case["kind"], case["padding_mode"], (y - case["y"]).abs().max())
        assert torch.allclose(x.grad, case["dx"], atol=1e-6, rtol=1e-6)
        pads.add(case["padding_mode"])
    assert pads == {"zeros", "border", "reflection"}


def test_gaussian_noise_matches_reference(golden_dir):
    fx = _load(golden_dir, "augment_aux.pt")
    for case in fx["noise"]:
        x = case["x"].clone().requires_grad_(True)
        y = O.gaussian_noise(x, case["noise"], case["sigma"])
        (y * case["dy"]).sum().backward()
        assert torch.equal(y, case["y"])
        assert torch.equal(x.grad, case["dx"])


def test_uint8_views_match_reference(golden_dir):
    """Row f3: augment_simclr(cat[ToTensor(bytes), ToTensor(bytes), fakes]) on the stored draws equals the reference
    chain on the converted batch."""
    fx = _load(golden_dir, "augment_aux.pt")
    for case in fx["uint8"]:
        fakes = case["fakes"].clone().requires_grad_(True)
        x_f = O.to_tensor_u8(case["x_u8"])
        y = O.augment_simclr(torch.cat([x_f, x_f, fakes], dim=0), _unpack(case["params"]), case["order"])
        (y * case["dy"]).sum().backward()
        # quantised inputs put more pixels near the hue wheel's kinks than U[0,1) floats do: 7e-6 worst case
        assert torch.allclose(y, case["y"], atol=2e-5, rtol=0), (y - case["y"]).abs().max()
        assert torch.allclose(fakes.grad, case["d_fakes"], atol=5e-5, rtol=1e-5)      # scatter-add order; |grad| up to 7


def test_baseline_modes_match_reference(golden_dir):
    """std / aug / aug_both with penalty none / cr / bcr under `--aug hfrt`: oracle composition vs the reference modules."""
    fx = _load(golden_dir, "baseline_modes.pt")
    seen = set()
    for case in fx["cases"]:
        sd = {k: v.clone() for k, v in case["sd_d"].items()}
        for k in O.trainable(sd):
            sd[k].requires_grad_(True)
        augs = [(lambda p: (lambda x: O.shift_flip(x, p, "reflection")))(p) for p in case["aug_params"]]
        d_loss, pen, d_real, d_gen = O.loss_d_baseline(sd, case["mode"], case["images"], case["gen"], augs, loss=case["loss"],
                                                       penalty=case["penalty"], lbd=case["lbd"], lbd2=case["lbd2"])
        assert abs(float(d_loss) - case["d_loss"]) < 1e-5 * max(1.0, abs(case["d_loss"]))
        assert abs(float(pen) - case["pen"]) < 1e-5 * max(1e-2, abs(case["pen"])), (float(pen), case["pen"])
        assert abs(float(d_real) - case["d_real"]) < 1e-5 and abs(float(d_gen) - case["d_gen"]) < 1e-5
        (d_loss + pen.sum()).backward()
        for k, ref in case["grad_norms"].items():
            got = float(sd[k].grad.double().norm())
            assert abs(got - ref) <= 1e-4 * max(ref, 1e-6) + 1e-9, (case["mode"], k, got, ref)
        seen.add((case["mode"], case["penalty"]))
    assert {m for m, _ in seen} == {"std", "aug", "aug_both"} and {p for _, p in seen} == {"none", "cr", "bcr"}


def test_diffaug_matches_reference(golden_dir):
    """third_party/diffaug.DiffAugment vs the oracle's slicing / masking restatement on replayed draws."""
    fx = _load(golden_dir, "diffaug.pt")
    for case in fx["cases"]:
        stages = tuple(case["policy"].split(","))
        b, _, h, w = case["x"].shape
        np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
        _ = torch.rand_like(case["x"]); _ = torch.randn_like(case["x"])
        assert torch.equal(O.sample_diffaug(b, h, w, stages), case["params"])
        x = case["x"].clone().requires_grad_(True)
        y = O.diffaug(x, case["params"], stages)
        (y * case["dy"]).sum().backward()
        assert torch.allclose(y, case["y"], atol=1e-6, rtol=0), (case["policy"], (y - case["y"]).abs().max())
        assert torch.allclose(x.grad, case["dx"], atol=2e-6, rtol=1e-5), case["policy"]


def test_reference_copy_is_unmodified_and_runs_one_cpu_step():
    """oracle/_ref (the copy of the reference that bench.py's `--impl reference`, `cpu_baseline` and `eager_gpu_baseline`
    legs execute; made by oracle/make_ref.py) must be byte-identical to its manifest, and the CPU driver must complete a
    step of the unmodified modules with finite losses.  Skipped when neither the copy nor /root/reference exists."""
    import os
    import numpy as np
    import pytest
    from oracle import make_ref, ref_import, ref_runner
    if not os.path.isdir(os.path.join(make_ref.DST, "augment")):
        if not os.path.isdir(os.path.join(make_ref.SRC, "augment")):
            pytest.skip("reference sources not available")
        make_ref.make(quiet=True)
    assert make_ref.verify()
    r = ref_runner.run_cpu(1, 0, batch=8, threads=2)
    assert r["steps"] == 1 and r["batch"] == 8 and all(np.isfinite(v) for v in r["last_losses"]), r
    assert abs(r["last_losses"][1] - 2 * np.log(2)) < 0.05          # L_dis at initialisation
    ref_import.deactivate()
