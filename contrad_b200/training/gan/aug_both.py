"""Mirror of training/gan/aug_both.py (real and generated images are augmented, also in the G step)."""
import torch

from ._baselines import d_loss_with_penalty, g_loss


def loss_D_fn(P, D, options, images, gen_images):
    """training/gan/aug_both.py:7-37."""
    assert images.size(0) == gen_images.size(0)
    gen_images = gen_images.detach()
    all_images = torch.cat([images, gen_images], dim=0)
    return d_loss_with_penalty(P, D, options, images, gen_images, all_images, P.augment_fn(all_images))


def loss_G_fn(P, D, options, images, gen_images):
    """training/gan/aug_both.py:40-43."""
    return g_loss(D, options, P.augment_fn(gen_images))
