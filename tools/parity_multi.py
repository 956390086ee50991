"""W > 1 parity of the data-parallel D step against the SINGLE-PROCESS full-batch step (BASELINE config 3).

Called by bench.py at N > 1 before the timed region (and by tests/test_gpu_multi.py under torchrun).  Every rank
builds the same seeded global batch (images, latents, augmentation draws for the 3N concatenation), takes its slice,
and runs the product's distributed D step: SyncBN generator forward, local augment + D, ONE packed all-gather of
the embeddings (reference: five GatherLayer calls, third_party/gather_layer.py:8-23, training/criterion.py:30-32,
training/gan/contrad.py:9-12), replicated contrastive losses, backward, gradient averaging (DDP semantics,
train_gan.py:311-313).  Rank 0 additionally evaluates the same global batch in one process (fresh non-distributed
copies of G and D with identical weights, P.distributed = False) through the same kernels.  Checked:

* L_con (= L_con+ + lbd_a * L_con-) of every rank equals the full-batch value (the loss is replicated);
* L_dis averaged over ranks equals the full-batch value (a mean over local samples);
* gradient norms: the contrastive gradient reaching backbone / projection parameters after DDP averaging is 1/W x the
  full-batch gradient, the L_dis gradient (it only reaches `linear.*`, sg_linear) is unscaled - so
  W * ||g_avg[non-linear params]|| and ||g_avg[linear.*]|| must both equal the single-process norms.
"""
import copy
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist


class _FixedAug(torch.nn.Module):
    def __init__(self, params, order):
        super().__init__()
        self.params, self.order = params, order

    def forward(self, x):
        from contrad_b200.functional import AugmentSimCLRFn
        return AugmentSimCLRFn.apply(x, self.params, self.order)


def _norms(model):
    lin, rest = 0.0, 0.0
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        sq = float(p.grad.double().pow(2).sum())
        if name.startswith("linear."):
            lin += sq
        else:
            rest += sq
    return lin ** 0.5, rest ** 0.5


def check(P, G, D, options, n_global, device, seed=777):
    """G, D: the (bare, broadcast) modules of this rank; G converted to SyncBatchNorm.  Returns a dict on rank 0
    (None elsewhere).  Module state (BN running statistics, spectral-norm vectors, gradients) is restored."""
    from contrad_b200 import engine
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import contrad
    world, rank = dist.get_world_size(), dist.get_rank()
    n = n_global // world
    sd_g, sd_d = copy.deepcopy(G.state_dict()), copy.deepcopy(D.state_dict())
    torch.manual_seed(seed); np.random.seed(seed)
    images = torch.rand(n_global, 3, 32, 32, device=device)
    z = torch.empty(n_global, G.nz).uniform_(-1, 1).to(device)
    carrier = torch.empty(1, device=device).expand(3 * n_global, 3, 32, 32)
    params, order = P.augment_fn.sample_params(carrier)
    sl = torch.arange(rank * n, (rank + 1) * n, device=device)
    cols = torch.cat([sl, sl + n_global, sl + 2 * n_global])
    G.train(); D.train()
    engine.set_grad(G, False); engine.set_grad(D, True)

    # ---- distributed step on this rank's slice
    Pd = SimpleNamespace(**vars(P))
    Pd.augment_fn, Pd.distributed = _FixedAug(params[:, cols].contiguous(), order), True
    with torch.no_grad():
        gen = G(z[sl])
    l_con, aux = contrad.loss_D_fn(Pd, D, options, images[sl], gen)
    D.zero_grad(set_to_none=True)
    (l_con + aux["penalty"]).backward()
    engine.allreduce_gradients(D)
    lin_d, rest_d = _norms(D)
    stats = torch.tensor([float(l_con), float(aux["penalty"])], device=device, dtype=torch.float64)
    all_stats = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(all_stats, stats)
    D.zero_grad(set_to_none=True)
    G.load_state_dict(sd_g); D.load_state_dict(sd_d)

    out = None
    if rank == 0:
        # ---- the same global batch in ONE process (no collectives), same kernels
        G1, D1 = get_architecture("sndcgan", (32, 32, 3))
        G1.load_state_dict(sd_g); D1.load_state_dict(sd_d)
        G1.to(device).train(); D1.to(device).train()
        engine.set_grad(G1, False)
        P1 = SimpleNamespace(**vars(P))
        P1.augment_fn, P1.distributed = _FixedAug(params.contiguous(), order), False
        with torch.no_grad():
            gen1 = G1(z)
        l_con1, aux1 = contrad.loss_D_fn(P1, D1, options, images, gen1)
        (l_con1 + aux1["penalty"]).backward()
        lin_1, rest_1 = _norms(D1)
        rel = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
        l_con_ranks = [float(s[0]) for s in all_stats]
        l_dis_mean = float(np.mean([float(s[1]) for s in all_stats]))
        out = {
            "world": world, "n_global": n_global,
            "L_con_full_batch": float(l_con1), "L_con_ranks_max_rel": max(rel(v, float(l_con1)) for v in l_con_ranks),
            "L_dis_full_batch": float(aux1["penalty"]), "L_dis_rank_mean_rel": rel(l_dis_mean, float(aux1["penalty"])),
            "D_grad_norm_nonlinear_xW_rel": rel(rest_d * world, rest_1),
            "D_grad_norm_linear_head_rel": rel(lin_d, lin_1),
            "D_grad_norm_full_batch": (lin_1 ** 2 + rest_1 ** 2) ** 0.5,
            "tolerance": 1e-3,
        }
        out["ok"] = bool(max(out["L_con_ranks_max_rel"], out["L_dis_rank_mean_rel"], out["D_grad_norm_nonlinear_xW_rel"],
                             out["D_grad_norm_linear_head_rel"]) < out["tolerance"])
        del G1, D1
    dist.barrier()
    return out
