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
    with pytest.raises(NotImplementedError):
        setup(SimpleNamespace(mode="std", aug="none", penalty="none", temp=0.1, lbd_a=1.0))
    with pytest.raises(NotImplementedError):
        get_architecture("stylegan2", (32, 32, 3))
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


def test_gather_layer_semantics_gloo_world2():
    """third_party/gather_layer.py:8-23: forward = rank-major all-gather, backward = own slice."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_gather_worker, args=(world, 29611, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_distributed_loss_equals_single_process_oracle():
    """Rank-major packing used by loss_D_fn(distributed=True): the re-ordered gathered rows equal
    [out1_all; out2_all; others_all] (checked on CPU against the oracle's loss on the full batch)."""
    from contrad_b200.training.gan.contrad import _rank_major
    torch.manual_seed(1)
    world, n, d = 2, 3, 4
    per_rank = [torch.randn(3 * n, d) for _ in range(world)]
    got = _rank_major(torch.stack(per_rank), n)
    want = torch.cat([torch.cat([p[i * n:(i + 1) * n] for p in per_rank]) for i in range(3)])
    assert torch.equal(got, want)
