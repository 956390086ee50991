"""Mirror of the StyleGAN2 + ContraD step functions of train_stylegan2_contraD.py:95-164 (``_loss_D_fn``,
``_loss_G_fn``, ``G_D``) and train_stylegan2.py:95-113 (``_update_lr``, ``r1_loss``), plus ``accumulate``
(utils.py:130-143), on the sm_100a kernels.  The reference keeps these inside its train CLIs; here they are an
importable module so the step can be driven by ``engine.train_step_stylegan2`` (bench / tests) or by the CLIs.

Differences that do not change results: the per-sample |grad|^2 of the R1 penalty is one reduction kernel
(``RowSqSum``) instead of pow/reshape/sum; the losses are the fused contrastive / GAN-loss kernels."""
import torch
from torch import autograd, nn

from ... import sg2_functional as SF
from ... import sg2_kernels as S
from ...functional import GanDLossFn, GanGLossFn, RowNormalizeFn, contrastive_loss
from ..criterion import nt_xent
from .contrad import supcon_fake


def update_lr(optimizer, cur_step, batch_size, halflife_lr, lr, mult=1.0):
    """train_stylegan2.py:95-103."""
    if halflife_lr > 0 and (cur_step > 0) and (cur_step % 1000 == 0):
        ratio = (cur_step * batch_size) / halflife_lr
        lr_w = (0.5 ** ratio) * lr * mult
        for group in optimizer.param_groups:
            group["lr"] = lr_w
        return lr_w
    return None


def accumulate(model_dst, model_src, decay=0.999):
    """utils.py:130-143: EMA of the generator parameters (one multi-tensor kernel), buffers copied."""
    model_dst = getattr(model_dst, "module", model_dst)
    model_src = getattr(model_src, "module", model_src)
    src = dict(model_src.named_parameters())
    pairs = [(p.data, src[k].data) for k, p in model_dst.named_parameters()]
    if pairs:
        S.ema_lerp(pairs, decay)
        SF.bump_weight_epoch()
    buf_src = dict(model_src.named_buffers())
    for k, b in model_dst.named_buffers():
        b.data.copy_(buf_src[k].data)


def r1_per_sample(D, images, augment_fn):
    """Per-sample squared norm of dD/dx at augment_fn(images), differentiable w.r.t. D's parameters
    (train_stylegan2.py:106-112; train_stylegan2_contraD.py:129-136)."""
    images_aug = augment_fn(images).detach()
    images_aug.requires_grad = True
    d_real = D(images_aug)
    grad_real, = autograd.grad(outputs=d_real.sum(), inputs=images_aug, create_graph=True, retain_graph=True)
    return SF.RowSqSum.apply(grad_real)


def r1_loss(D, images, augment_fn):
    """train_stylegan2.py:106-113."""
    return r1_per_sample(D, images, augment_fn).mean()


def discriminate(D, real_aug2, fake_aug, train_G=False):
    """The discriminator half of G_D.forward (train_stylegan2_contraD.py:143-164) on already augmented batches:
    D on the fakes [n] and on cat(real, real) [2n] in two calls (two minibatch-stddev groupings)."""
    d_gen, aux_f = D(fake_aug, sg_linear=(not train_G), projection=True, projection2=True)
    if train_G:
        return d_gen
    d_rs, aux_r = D(real_aug2, sg_linear=True, projection=True, projection2=True)
    views_r = RowNormalizeFn.apply(aux_r["projection"])
    reals = RowNormalizeFn.apply(aux_r["projection2"])
    others = RowNormalizeFn.apply(aux_f["projection"])
    fakes = RowNormalizeFn.apply(aux_f["projection2"])
    n = fake_aug.size(0)
    return (d_rs[:n], d_gen), (views_r[:n], views_r[n:], others), (reals[:n], reals[n:], fakes)


def loss_D_fn(P, d_all, view_r, view_f):
    """`_loss_D_fn` (train_stylegan2_contraD.py:95-109)."""
    d_real, d_gen = d_all
    view1, view2, _others = view_r
    real1, real2, fakes = view_f
    simclr_loss = nt_xent(view1, view2, temperature=P.temp, distributed=P.distributed)
    sup_loss = supcon_fake(real1, real2, fakes, temperature=P.temp, distributed=P.distributed)
    n = d_real.size(0)
    d_loss, means = GanDLossFn.apply(torch.cat([d_real, d_real, d_gen], dim=0), n, "nonsat")
    return simclr_loss + P.lbd_a * sup_loss, {"penalty": d_loss, "d_real": means[0], "d_gen": means[1]}


def loss_G_fn(d_gen):
    """`_loss_G_fn` (train_stylegan2_contraD.py:112-114)."""
    return GanGLossFn.apply(d_gen, "nonsat")


class G_D(nn.Module):
    """train_stylegan2_contraD.py:117-164: the per-replica G -> augment -> D pipeline (under nn.DataParallel the
    contrastive losses then run on the gathered embeddings, SURVEY 8e: replicas only)."""

    def __init__(self, G, D, augment_fn):
        super().__init__()
        self.G, self.D, self.augment_fn = G, D, augment_fn

    def forward(self, P, real_images, style_mix=0.9, train_G=False, return_r1_loss=False):
        if return_r1_loss:
            return r1_per_sample(self.D, real_images, self.augment_fn)
        with torch.set_grad_enabled(train_G):
            latent_samples = self.G.sample_latent(real_images.size(0))
            gen_images = self.G(latent_samples, style_mix=style_mix)
        fake_aug = self.augment_fn(gen_images)
        if train_G:
            return discriminate(self.D, None, fake_aug, train_G=True)
        cat_images = torch.cat([real_images, real_images], dim=0)
        # the reference augments the fakes first, then the reals (RNG order of train_stylegan2_contraD.py:143-150)
        return discriminate(self.D, self.augment_fn(cat_images), fake_aug)
