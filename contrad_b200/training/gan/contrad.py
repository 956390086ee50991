"""Mirror of training/gan/contrad.py: ``supcon_fake``, ``loss_D_fn``, ``loss_G_fn`` with the reference
signatures; the arithmetic runs on the sm_100a kernels (one fused contrastive kernel per loss, one
fused GAN-loss kernel) and, when ``P.distributed``, on ONE packed all-gather of the embeddings."""
import torch

from ...functional import contrastive_loss, GanDLossFn, GanGLossFn, RowNormalizeFn
from ...third_party.gather_layer import GatherLayer, gather_rows

_D_LOSSES = ("nonsat", "wgan", "hinge", "lsgan")


def supcon_fake(out1, out2, others, temperature, distributed=False):
    """training/gan/contrad.py:8-32."""
    if distributed:
        out1 = torch.cat(GatherLayer.apply(out1), dim=0)
        out2 = torch.cat(GatherLayer.apply(out2), dim=0)
        others = torch.cat(GatherLayer.apply(others), dim=0)
    n = out1.size(0)
    return contrastive_loss(torch.cat([out1, out2, others], dim=0), n, 1, float(temperature))


def _rank_major(blocks, n):
    """[world, 3n, d] gathered rows -> [out1_all; out2_all; others_all], each rank-major (the order
    torch.cat(GatherLayer.apply(x)) produces in the reference, criterion.py:31-32)."""
    world, _, d = blocks.shape
    return blocks.view(world, 3, n, d).transpose(0, 1).reshape(3 * world * n, d)


def loss_D_fn(P, D, options, images, gen_images):
    """training/gan/contrad.py:35-70."""
    assert images.size(0) == gen_images.size(0)
    if options["loss"] not in _D_LOSSES:
        raise NotImplementedError()
    gen_images = gen_images.detach()
    n = images.size(0)

    if images.dtype == torch.uint8:
        # row f3: raw dataset bytes; ToTensor, the two-view duplication and the concatenation happen inside the kernel
        if not hasattr(P.augment_fn, "forward_views"):
            raise TypeError("uint8 images need a fused augmentation (simclr / simclr_hq / simclr_hq_cutout); "
                            "got %s - convert with images.float().div(255) first" % type(P.augment_fn).__name__)
        aug_images = P.augment_fn.forward_views(images, 2, gen_images)
    else:
        aug_images = P.augment_fn(torch.cat([images, images, gen_images], dim=0))
    d_all, aux = D(aug_images, sg_linear=True, projection=True, projection2=True)
    views = RowNormalizeFn.apply(aux["projection"])
    reals = RowNormalizeFn.apply(aux["projection2"])
    if P.distributed:
        both = gather_rows(torch.cat([views, reals], dim=1))          # one collective for all 5 blocks
        d = views.shape[1]
        views_all = _rank_major(both[:, :, :d].contiguous(), n)
        reals_all = _rank_major(both[:, :, d:].contiguous(), n)
        n_all = n * both.shape[0]
        simclr_loss = contrastive_loss(views_all[:2 * n_all], n_all, 0, float(P.temp))
        sup_loss = contrastive_loss(reals_all, n_all, 1, float(P.temp))
    else:
        simclr_loss = contrastive_loss(views[:2 * n], n, 0, float(P.temp))
        sup_loss = contrastive_loss(reals, n, 1, float(P.temp))

    d_loss, means = GanDLossFn.apply(d_all, n, options["loss"])
    return simclr_loss + P.lbd_a * sup_loss, {
        "penalty": d_loss,
        "d_real": means[0],
        "d_gen": means[1],
    }


def loss_G_fn(P, D, options, images, gen_images):
    """training/gan/contrad.py:73-82."""
    d_gen = D(P.augment_fn(gen_images))
    return GanGLossFn.apply(d_gen, options["loss"])
