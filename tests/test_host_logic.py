"""CPU tests of the host side: the C ABI library loads and exports every declared symbol, the drop-in
mirrors keep the reference surfaces, the product's parameter sampler replays the reference RNG stream,
mini-gin semantics, and GatherLayer semantics under gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import contrad_oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(REPO, "contrad_b200", "compat")
if COMPAT not in sys.path:
    sys.path.append(COMPAT)


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from contrad_b200 import _capi
    lib = _capi.lib()
    names = _capi.declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert lib.cb200_version() >= 100


def test_product_refuses_cpu_tensors():
    from contrad_b200 import _capi, kernels
    with pytest.raises(_capi.CB200Error):
        kernels.augment_simclr_fwd(torch.rand(2, 3, 32, 32), torch.zeros(11, 2), 0)


def _gin_defaults():
    import gin
    gin.clear_config()
    gin.parse_config("""
# comment line
ColorJitterLayer.brightness = 0.4
ColorJitterLayer.contrast = 0.4
ColorJitterLayer.saturation = 0.4
ColorJitterLayer.hue = 0.1   # trailing comment
RandomResizeCropLayer.scale = (0.2,
                               1.0)
""")
    return gin


def test_fused_simclr_sampler_replays_reference_stream(golden_dir):
    """contrad_b200.augment draws its parameters in the reference's numpy/torch order: with the seeds of
    the golden fixtures it reproduces the parameter blocks that reproduced the reference outputs."""
    _gin_defaults()
    from contrad_b200.augment import get_augment
    aug = get_augment("simclr")
    fx = torch.load(os.path.join(golden_dir, "augment_simclr.pt"), weights_only=False)
    for case in fx["cases"]:
        b, _, h, w = case["x"].shape
        np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
        x = torch.rand(b, 3, h, w); _ = torch.randn(b, 3, h, w)
        params, order = aug.sample_params(x)
        assert order == case["order"]
        assert torch.equal(params, case["params"])


def test_gin_shim_semantics():
    gin = _gin_defaults()
    from contrad_b200.augment.layers import ColorJitterLayer, RandomResizeCropLayer
    assert RandomResizeCropLayer().scale == (0.2, 1.0)
    assert RandomResizeCropLayer(scale=(0.5, 1.0)).scale == (0.5, 1.0)       # explicit argument wins
    cj = ColorJitterLayer()
    assert cj.hue == [-0.1, 0.1] and cj.contrast == [0.6, 1.4]

    @gin.configurable("options")
    def get_options(dataset=gin.REQUIRED, loss=gin.REQUIRED, batch_size=64):
        return dataset, loss, batch_size

    with pytest.raises(RuntimeError):
        get_options()
    gin.parse_config(['options.dataset = "cifar10"', 'options.loss = "nonsat"', "options.batch_size = 512"])
    assert get_options() == ("cifar10", "nonsat", 512)
    assert gin.query_parameter("options.batch_size") == 512
    gin.clear_config()


def test_module_surfaces_match_reference():
    _gin_defaults()
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import setup
    from types import SimpleNamespace
    G, D = get_architecture("sndcgan", (32, 32, 3))
    ref_d, ref_g = O.make_d_state(), O.make_g_state()
    sd = D.state_dict()
    assert set(sd) == set(ref_d) and all(sd[k].shape == ref_d[k].shape for k in sd)
    assert set(G.state_dict()) == set(ref_g)
    D.load_state_dict(ref_d); G.load_state_dict(ref_g)
    assert D.d_penul == 8192 and hasattr(D, "penultimate") and hasattr(D, "reset_parameters")
    P = setup(SimpleNamespace(mode="contrad", aug="simclr", lbd_a=1.0, temp=0.1, penalty="none"))
    assert P.filename == "contrad_simclr_L1.0_T0.1" and set(P.train_fn) == {"G", "D"}
    for mode, pen, aug, name in (("std", "none", "none", "std_none"), ("std", "bcr", "hfrt", "std_bcr_hfrt"),
                                 ("aug", "cr", "hfrt", "aug_hfrt_cr"), ("aug_both", "none", "simclr", "aug_both_simclr_none"),
                                 ("simclr_only", "none", "simclr", "simclr_only_simclr_T0.1")):
        Pb = setup(SimpleNamespace(mode=mode, aug=aug, penalty=pen, temp=0.1, lbd_a=1.0))
        assert Pb.filename == name and set(Pb.train_fn) == {"G", "D"}          # training/gan/__init__.py:9-20
    with pytest.raises(NotImplementedError):
        setup(SimpleNamespace(mode="no_such_mode", aug="none", penalty="none", temp=0.1, lbd_a=1.0))
    with pytest.raises(NotImplementedError):
        get_architecture("biggan", (32, 32, 3))
    G3, D3 = get_architecture("snresnet18", (32, 32, 3))
    ref_r = O.make_d_resnet18_state()
    assert {k: tuple(v.shape) for k, v in D3.state_dict().items()} == {k: tuple(v.shape) for k, v in ref_r.items()}
    assert D3.d_penul == 512 and isinstance(G3, type(G))
    # the StyleGAN2 networks also run on the kernels only (no CPU fallback)
    G2, D2 = get_architecture("stylegan2", (32, 32, 3))
    from contrad_b200._capi import CB200Error as _E
    with pytest.raises(_E):
        D2(torch.rand(4, 3, 32, 32))
    with pytest.raises(_E):
        G2(torch.randn(4, 512))
    # the generator runs on the sm_100a kernels only: on CPU tensors it must fail loudly, not fall back
    from contrad_b200._capi import CB200Error
    with pytest.raises(CB200Error):
        G(O.sample_latent(4))


def _gather_worker(rank, world, port, results):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from contrad_b200.third_party.gather_layer import GatherLayer, gather_rows
    torch.manual_seed(0)
    full = torch.randn(world * 3, 8)
    mine = full[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    cat = torch.cat(GatherLayer.apply(mine), dim=0)
    ok = torch.equal(cat, full)
    w = torch.arange(float(world * 3 * 8)).view(world * 3, 8)
    (cat * w).sum().backward()
    ok = ok and torch.equal(mine.grad, w[rank * 3:(rank + 1) * 3])
    mine2 = full[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    g = gather_rows(mine2)
    ok = ok and torch.equal(g.reshape(world * 3, 8), full)
    (g.reshape(world * 3, 8) * w).sum().backward()
    ok = ok and torch.equal(mine2.grad, w[rank * 3:(rank + 1) * 3])
    results[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gather_layer_semantics_gloo_world2(world):
    """third_party/gather_layer.py:8-23: forward = rank-major all-gather, backward = own slice (world sizes 2 and 4: the
    scaling bench runs 1 / 2 / 4 / 8 ranks)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_gather_worker, args=(world, 29611 + world, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_distributed_loss_equals_single_process_oracle():
    """Rank-major packing used by loss_D_fn(distributed=True): the re-ordered gathered rows equal
    [out1_all; out2_all; others_all] (checked on CPU against the oracle's loss on the full batch)."""
    from contrad_b200.training.gan.contrad import _rank_major
    torch.manual_seed(1)
    for world, n, d in ((2, 3, 4), (8, 64, 128)):
        per_rank = [torch.randn(3 * n, d) for _ in range(world)]
        got = _rank_major(torch.stack(per_rank), n)
        want = torch.cat([torch.cat([p[i * n:(i + 1) * n] for p in per_rank]) for i in range(3)])
        assert torch.equal(got, want)


def test_staging_eager_and_recorder_on_cpu():
    """staging.stage: eager = the producer's value on the device; Recorder.plan() = an eager step that lays out
    static buffers; Recorder.capture() hands the same buffers out again without calling the producers; refresh()
    refills them by calling the producers in the recorded order (host RNG advances like the eager loop)."""
    from contrad_b200 import staging
    calls = []

    def producer_a():
        calls.append("a")
        return torch.full((2, 3), float(len(calls)))

    def producer_b():
        calls.append("b")
        return torch.full((1,), -float(len(calls)))

    def filler_c(out):
        calls.append("c")
        out.fill_(100.0 + len(calls))

    assert torch.equal(staging.stage(producer_a, "cpu"), torch.full((2, 3), 1.0)) and calls == ["a"]
    assert float(staging.stage(filler_c, "cpu", shape=(2,), fill=True)[0]) == 102.0 and calls == ["a", "c"]
    del calls[:]
    rec = staging.Recorder(slots=2)
    with rec.plan():
        assert staging.recording()
        buf_a = staging.stage(producer_a, "cpu", shape=(2, 3))
        buf_b = staging.stage(producer_b, "cpu")
        buf_c = staging.stage(filler_c, "cpu", shape=(2,), fill=True)
    # the planned step is a real step: every producer ran exactly once and its value is in the static buffer
    assert not staging.recording() and calls == ["a", "b", "c"]
    assert float(buf_a[0, 0]) == 1.0 and float(buf_b[0]) == -2.0 and float(buf_c[0]) == 103.0
    with rec.capture():
        assert staging.stage(producer_a, "cpu", shape=(2, 3)) is buf_a
        assert staging.stage(producer_b, "cpu") is buf_b
        assert staging.stage(filler_c, "cpu", shape=(2,), fill=True) is buf_c
    assert calls == ["a", "b", "c"]                                 # capture draws nothing
    for k in range(3):                                              # more refreshes than pinned slots
        rec.refresh()
        assert calls[-3:] == ["a", "b", "c"]
        assert float(buf_a[0, 0]) == len(calls) - 2 and float(buf_b[0]) == -(len(calls) - 1)
        assert float(buf_c[1]) == 100.0 + len(calls)
    n = len(calls)
    rec.produce()                                                   # host half only: buffers unchanged until upload()
    assert len(calls) == n + 3 and float(buf_a[0, 0]) == n - 2
    rec.upload()
    assert float(buf_a[0, 0]) == n + 1
    with pytest.raises(RuntimeError):                               # a captured step that stages fewer inputs
        with rec.capture():
            staging.stage(producer_a, "cpu", shape=(2, 3))
    with pytest.raises(RuntimeError):
        with rec.capture():
            staging.stage(producer_a, "cpu", shape=(5, 3))


def test_fused_simclr_host_draw_order_matches_reference_in_recording_mode():
    """Under a Recorder the numpy draws (crop boxes, then the jitter order) happen in the eager order, so a graphed
    loop consumes the host RNG stream exactly like the reference loop."""
    _gin_defaults()
    from contrad_b200 import staging
    from contrad_b200.augment.layers import ColorJitterLayer, RandomResizeCropLayer, _ShapeOnly
    rrc, cj = RandomResizeCropLayer(scale=(0.2, 1.0)), ColorJitterLayer(0.4, 0.4, 0.4, 0.1)
    shape = _ShapeOnly((6, 3, 32, 32))
    np.random.seed(11)
    want = [(rrc.sample(shape), cj.draw_order()) for _ in range(4)]
    np.random.seed(11)
    rec = staging.Recorder()

    def one_step():
        box = staging.stage(lambda: rrc.sample(shape), "cpu", shape=(4, 6))
        order = staging.stage(lambda: torch.tensor([float(cj.draw_order())]), "cpu", shape=(1,))
        return box, order

    with rec.plan():
        box, order = one_step()
    assert torch.equal(box, want[0][0]) and float(order) == want[0][1]
    with rec.capture():
        box2, order2 = one_step()
    assert box2 is box and order2 is order
    for w_box, w_order in want[1:]:
        rec.refresh()
        assert torch.equal(box, w_box) and float(order) == w_order


def _grad_sync_worker(rank, world, port, results):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from contrad_b200 import engine
    torch.manual_seed(rank)                      # different init per rank -> broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3), torch.nn.Linear(3, 1, bias=False))
    engine.broadcast_parameters(model, src=0)
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3), torch.nn.Linear(3, 1, bias=False))
    ok = all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), ref.state_dict().values()))
    torch.manual_seed(5)
    data = torch.randn(world, 8, 4)
    model.eval()
    model(data[rank]).sum().backward()
    engine.allreduce_gradients(model)
    ref.eval()
    (sum(ref(data[r]).sum() for r in range(world)) / world).backward()
    ok = ok and all(torch.allclose(a.grad, b.grad, atol=1e-6) for a, b in zip(model.parameters(), ref.parameters()))
    results[rank] = bool(ok)
    dist.destroy_process_group()


def test_engine_gradient_sync_gloo_world2():
    """engine.broadcast_parameters + engine.allreduce_gradients = what DDP does around backward
    (train_gan.py:311-313): same start everywhere, gradients averaged over ranks."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_grad_sync_worker, args=(world, 29613, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_staging_late_entries_are_produced_at_upload_time():
    """late=True (Adam's step scalars): produce() - which runs one step AHEAD, overlapping the GPU - must not touch
    them; upload() draws them right before the replay."""
    from contrad_b200 import staging
    state = {"t": 0}

    def hyper():
        state["t"] += 1
        return torch.tensor([float(state["t"])])

    rec = staging.Recorder()
    with rec.plan():
        buf = staging.stage(hyper, "cpu", shape=(1,), late=True)
    assert state["t"] == 1 and float(buf) == 1.0
    with rec.capture():
        staging.stage(hyper, "cpu", shape=(1,), late=True)
    rec.produce()
    assert state["t"] == 1
    rec.upload()
    assert state["t"] == 2 and float(buf) == 2.0


def test_shift_flip_index_function_compiled_for_host_matches_oracle(tmp_path):
    """The index arithmetic of cb200_shift_flip_* (csrc/augment_aux.cu: nearest_source) is plain C: compile that function
    with g++ and compare the source tables it produces with the oracle's (which is pinned bit-exactly on the reference's
    affine_grid + nearest grid_sample) for every padding mode, odd sizes and shifts beyond the image."""
    import ctypes
    import subprocess
    src = open(os.path.join(REPO, "contrad_b200", "csrc", "augment_aux.cu")).read()
    body = src[src.index("// [host-testable: nearest_source]"):src.index("// [host-testable: end]")]
    cpp = tmp_path / "nearest.cpp"
    cpp.write_text("#include <math.h>\n#define __device__\n#define __forceinline__ inline\n"
                   "enum PadMode { kZeros = 0, kBorder = 1, kReflection = 2 };\n" + body +
                   'extern "C" void table(float sign, float bias, int n, int pad, int* out) {\n'
                   "  for (int e = 0; e < n; ++e) { float base = (2.f * (float)e + 1.f) / (float)n - 1.f;\n"
                   "    out[e] = nearest_source(sign * base + bias, n, pad); } }\n")
    so = tmp_path / "nearest.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(cpp), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    rng = np.random.RandomState(0)
    checked = 0
    for pad_id, pad in enumerate(("zeros", "border", "reflection")):
        for n in (4, 20, 28, 32, 33, 64, 512):
            for _ in range(12):
                sign = float(rng.choice([-1.0, 1.0]))
                shift = int(rng.randint(-n - 3, n + 4))
                bias = np.float32(shift) / np.float32(n / 2)
                out = (ctypes.c_int * n)()
                lib.table(ctypes.c_float(sign), ctypes.c_float(bias), n, pad_id, out)
                g = torch.tensor(sign) * ((2.0 * torch.arange(n, dtype=torch.float32) + 1.0) / n - 1.0) + torch.tensor(bias)
                idx, valid = O._nearest_index(g, n, pad)
                want = torch.where(valid, idx, torch.full_like(idx, -1)).tolist()
                assert list(out) == want, (pad, n, sign, shift)
                checked += 1
    assert checked == 3 * 7 * 12


def _load_fx(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_forward_views_uint8_host_logic(golden_dir):
    """Row f3: FusedSimCLR.forward_views(bytes, 2, fakes) draws the parameters of all 2n + m views in the reference's
    order (same seed -> the views the reference produces for cat[ToTensor(x), ToTensor(x), fakes]) and routes the
    gradient to the fp32 images only.  Kernel bindings replaced by the oracle stand-ins."""
    import tests.cpu_augment_standins as AS
    _gin_defaults()
    from contrad_b200.augment import get_augment
    aug = get_augment("simclr")
    fx = _load_fx(golden_dir, "augment_aux.pt")
    with AS.patched():
        for case in fx["uint8"]:
            np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
            fakes = case["fakes"].clone().requires_grad_(True)
            y = aug.forward_views(case["x_u8"], 2, fakes)
            assert torch.allclose(y, case["y"], atol=2e-5, rtol=0), (y - case["y"]).abs().max()
            (y * case["dy"]).sum().backward()
            assert torch.allclose(fakes.grad, case["d_fakes"], atol=5e-5, rtol=1e-5)
        # only fp32 views: same as the ordinary call; only uint8 views: no gradient anywhere
        case = fx["uint8"][0]
        np.random.seed(1); torch.manual_seed(1)
        y1 = aug.forward_views(case["x_u8"], 1, None)
        np.random.seed(1); torch.manual_seed(1)
        y2 = aug(O.to_tensor_u8(case["x_u8"]))
        assert torch.equal(y1, y2) and not y1.requires_grad
    with pytest.raises(ValueError):
        aug.forward_views(case["fakes"], 2, None)            # not uint8
    with pytest.raises(ValueError):
        aug.forward_views(case["x_u8"], 2, case["fakes"][:, :, :16])


def test_loss_d_fn_accepts_uint8_images(golden_dir):
    """training/gan/contrad.py:35-70 with raw uint8 images == the same call on ToTensor(images) (same seed)."""
    import tests.cpu_augment_standins as AS
    import tests.cpu_loss_standins as LS
    from types import SimpleNamespace
    _gin_defaults()
    from contrad_b200.augment import get_augment
    from contrad_b200.training.gan import contrad
    fxs = _load_fx(golden_dir, "sndcgan_small.pt")["nonsat"]
    sd = {k: v.clone() for k, v in fxs["sd_d"].items()}

    class OracleD(torch.nn.Module):
        def forward(self, x, sg_linear=False, projection=False, projection2=False):
            local = {k: v.clone() for k, v in sd.items()}
            d, aux = O.d_sndcgan_forward(local, x, sg_linear=sg_linear)
            return d, aux

    x_u8 = torch.randint(0, 256, (4, 3, 32, 32), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    gen = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    P = SimpleNamespace(augment_fn=get_augment("simclr"), temp=0.1, lbd_a=1.0, distributed=False)
    out = []
    with AS.patched(), LS.patched():
        for images in (x_u8, O.to_tensor_u8(x_u8)):
            np.random.seed(9); torch.manual_seed(9)
            loss, aux = contrad.loss_D_fn(P, OracleD(), {"loss": "nonsat"}, images, gen)
            out.append((float(loss), float(aux["penalty"]), float(aux["d_real"]), float(aux["d_gen"])))
        assert out[0] == pytest.approx(out[1], rel=1e-6, abs=1e-7)
        P2 = SimpleNamespace(augment_fn=get_augment("hflip"), temp=0.1, lbd_a=1.0, distributed=False)
        with pytest.raises(TypeError):
            contrad.loss_D_fn(P2, OracleD(), {"loss": "nonsat"}, x_u8, gen)


def test_shift_flip_and_gaussian_layers_replay_reference_stream(golden_dir):
    """Row f4: the product's HorizontalFlipRandomCrop / RandomCrop / Gaussian layers draw in the reference order (same
    seed -> the stored reference outputs); kernels replaced by the oracle stand-ins."""
    import tests.cpu_augment_standins as AS
    from contrad_b200.augment.layers import Gaussian, HorizontalFlipRandomCrop, RandomCrop
    fx = _load_fx(golden_dir, "augment_aux.pt")
    with AS.patched():
        for case in fx["shift_flip"]:
            cls = HorizontalFlipRandomCrop if case["kind"] == "hfrt" else RandomCrop
            layer = cls(max_pixels=case["max_pixels"], width=case["width"], padding_mode=case["padding_mode"])
            assert set(layer.state_dict()) == {"_eye"}
            np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
            x = torch.rand_like(case["x"]); _ = torch.randn_like(case["x"])
            x.requires_grad_(True)
            y = layer(x)
            assert torch.equal(y, case["y"])
            (y * case["dy"]).sum().backward()
            assert torch.allclose(x.grad, case["dx"], atol=1e-6, rtol=1e-6)
        for case in fx["noise"]:
            np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
            x = torch.rand_like(case["x"]); _ = torch.randn_like(case["x"])
            x.requires_grad_(True)
            y = Gaussian(sigma=case["sigma"])(x)
            assert torch.equal(y, case["y"])
            (y * case["dy"]).sum().backward()
            assert torch.equal(x.grad, case["dx"])
    with pytest.raises(ValueError):
        RandomCrop(max_pixels=4, width=32, padding_mode="circular")


def test_baseline_training_modes_host_logic(golden_dir):
    """training/gan/{std,aug,aug_both}.py + penalty cr / bcr under --aug hfrt: the product's mode modules reproduce the
    reference's scalars and D gradient norms when the discriminator and the kernels are oracle stand-ins (checks the
    wiring: which tensors are augmented, what the penalty sees, RNG order, the generalised GAN-loss offsets)."""
    import tests.cpu_augment_standins as AS
    from types import SimpleNamespace
    gin = _gin_defaults()
    gin.parse_config("""
HorizontalFlipRandomCrop.max_pixels = 4
HorizontalFlipRandomCrop.width = 32
HorizontalFlipRandomCrop.padding_mode = "reflection"
""")
    from contrad_b200.augment import get_augment
    from contrad_b200.training.gan import setup
    fx = _load_fx(golden_dir, "baseline_modes.pt")

    class OracleD(torch.nn.Module):
        def __init__(self, sd):
            super().__init__()
            self.sd = {k: v.clone() for k, v in sd.items()}
            for k in O.trainable(self.sd):
                self.sd[k].requires_grad_(True)

        def forward(self, x):
            return O.d_sndcgan_forward(self.sd, x)[0]

    with AS.patched():
        for case in fx["cases"]:
            P = setup(SimpleNamespace(mode=case["mode"], aug="hfrt", penalty=case["penalty"], temp=0.1, lbd_a=1.0,
                                      distributed=False))
            assert P.filename.startswith(case["mode"])
            P.augment_fn = get_augment("hfrt")
            D = OracleD(case["sd_d"])
            # the generator seeds of make_golden.gen_baselines: images / gen first, then the mode's own draws
            seed = 400 + fx["cases"].index(case)
            np.random.seed(seed + 50); torch.manual_seed(seed + 50)
            images = torch.rand(4, 3, 32, 32); gen = torch.rand(4, 3, 32, 32)
            assert torch.equal(images, case["images"])
            options = {"loss": case["loss"], "lbd": case["lbd"], "lbd2": case["lbd2"]}
            d_loss, aux = P.train_fn["D"](P, D, options, images, gen)
            assert float(d_loss) == pytest.approx(case["d_loss"], rel=1e-5, abs=1e-6)
            assert float(aux["penalty"]) == pytest.approx(case["pen"], rel=1e-4, abs=1e-7)
            assert float(aux["d_real"]) == pytest.approx(case["d_real"], abs=1e-5)
            assert float(aux["d_gen"]) == pytest.approx(case["d_gen"], abs=1e-5)
            (d_loss + aux["penalty"]).sum().backward()
            for k, ref in case["grad_norms"].items():
                got = float(D.sd[k].grad.double().norm())
                assert abs(got - ref) <= 1e-4 * max(ref, 1e-6) + 1e-9, (case["mode"], case["penalty"], k, got, ref)
            g_loss = P.train_fn["G"](P, D, options, images, gen)
            assert float(g_loss) == pytest.approx(case["g_loss"], rel=1e-5, abs=1e-6)
    from contrad_b200.penalty import compute_penalty
    with pytest.raises(NotImplementedError):
        compute_penalty("gp", D=None, images=None, gen_images=None, lbd=1.0)


def _compile_host_section(tmp_path, cu_file, begin, end, extra):
    """g++-compile the part of a .cu file between two "[host-testable: ...]" markers under tests/cuda_host_shim.h."""
    import ctypes
    import subprocess
    src = open(os.path.join(REPO, "contrad_b200", "csrc", cu_file)).read()
    body = src[src.index(begin):src.index(end)]
    cpp = tmp_path / "section.cpp"
    cpp.write_text('#include "%s"\nnamespace {\n%s\n}\n%s' % (os.path.join(REPO, "tests", "cuda_host_shim.h"), body, extra))
    so = tmp_path / "section.so"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-w", str(cpp), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def test_diffaug_kernels_compiled_for_host_match_reference(golden_dir, tmp_path):
    """The four DiffAugment kernels of csrc/augment_aux.cu, run single-threaded on the host through tests/cuda_host_shim.h,
    reproduce the reference's DiffAugment outputs and input gradients on the stored draws (all canonical policies, odd
    sizes): checks the index arithmetic (translation, clamped cutout range), the adjoint and the scratch protocol."""
    import ctypes
    lib = _compile_host_section(tmp_path, "augment_aux.cu", "// [host-testable: diffaug]", "// [host-testable: end diffaug]", """
extern "C" void run_fwd(const float* x, float* y, const float* p, float* sums, int B, int H, int W, int flags) {
  gridDim.y = B;
  for (int b = 0; b < B; ++b) { blockIdx.y = b; sums[b] = 0.f; if (flags & 1) diffaug_mean_kernel(x, p, sums, B, H, W, flags); }
  for (int b = 0; b < B; ++b) { blockIdx.y = b; diffaug_apply_kernel(x, y, p, sums, B, H, W, flags); }
}
extern "C" void run_bwd(const float* dy, float* dx, const float* p, float* gs, int B, int H, int W, int flags) {
  gridDim.y = B;
  for (int b = 0; b < B; ++b) { blockIdx.y = b; gs[b] = 0.f; if (flags & 1) diffaug_bwd_sum_kernel(dy, p, gs, B, H, W, flags); }
  for (int b = 0; b < B; ++b) { blockIdx.y = b; diffaug_bwd_apply_kernel(dy, dx, p, gs, B, H, W, flags); }
}
""")
    fptr = lambda t: ctypes.c_void_p(t.data_ptr())
    fx = _load_fx(golden_dir, "diffaug.pt")
    for case in fx["cases"]:
        x, dy, p = case["x"].contiguous(), case["dy"].contiguous(), case["params"].contiguous()
        b, _, h, w = x.shape
        flags = sum({"color": 1, "translation": 2, "cutout": 4}[s] for s in case["policy"].split(","))
        y, dx, scratch = torch.empty_like(x), torch.empty_like(x), torch.empty(b)
        lib.run_fwd(fptr(x), fptr(y), fptr(p), fptr(scratch), b, h, w, flags)
        assert torch.allclose(y, case["y"], atol=2e-6, rtol=0), (case["policy"], float((y - case["y"]).abs().max()))
        lib.run_bwd(fptr(dy), fptr(dx), fptr(p), fptr(scratch), b, h, w, flags)
        assert torch.allclose(dx, case["dx"], atol=5e-6, rtol=1e-5), (case["policy"], float((dx - case["dx"]).abs().max()))


def test_diffaug_layer_replays_reference_stream(golden_dir):
    """DiffAugLayer draws in the reference order; with the kernel binding replaced by the oracle it reproduces the
    reference outputs for the same seed.  get_augment knows every mode of the reference's registry."""
    import tests.cpu_augment_standins as AS
    _gin_defaults()
    from contrad_b200.augment import get_augment
    from contrad_b200.augment.layers import DiffAugLayer
    fx = _load_fx(golden_dir, "diffaug.pt")
    with AS.patched():
        for case in fx["cases"]:
            layer = DiffAugLayer(policy=case["policy"])
            np.random.seed(case["seed"]); torch.manual_seed(case["seed"])
            x = torch.rand_like(case["x"]); _ = torch.randn_like(case["x"])
            x.requires_grad_(True)
            y = layer(x)
            assert torch.allclose(y, case["y"], atol=2e-6, rtol=0)
            (y * case["dy"]).sum().backward()
            assert torch.allclose(x.grad, case["dx"], atol=5e-6, rtol=1e-5)
    assert isinstance(get_augment("diffaug"), DiffAugLayer) and get_augment("diffaug").policy == "color,cutout"
    x = torch.rand(2, 3, 8, 8)
    assert DiffAugLayer(policy="")(x) is x
    with pytest.raises(NotImplementedError):
        DiffAugLayer(policy="cutout,color")
    with pytest.raises(KeyError):
        DiffAugLayer(policy="rotate")
    with pytest.raises(KeyError):
        get_augment("no_such_mode")
    for mode in ("none", "gaussian", "hflip", "hfrt", "color_jitter", "cutout", "simclr", "simclr_hq", "simclr_hq_cutout", "diffaug"):
        gin = _gin_defaults()
        gin.parse_config("""
Gaussian.sigma = 0.12
CutOut.length = 15
GaussianBlur.sigma_range = (0.1, 2.0)
HorizontalFlipRandomCrop.max_pixels = 4
HorizontalFlipRandomCrop.width = 32
HorizontalFlipRandomCrop.padding_mode = "reflection"
""")
        assert isinstance(get_augment(mode), torch.nn.Module), mode            # augment/__init__.py:14-25
