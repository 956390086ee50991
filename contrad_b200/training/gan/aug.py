"""Mirror of training/gan/aug.py (only the real images are augmented)."""
import torch

from ._baselines import d_loss_with_penalty, g_loss


def loss_D_fn(P, D, options, images, gen_images):
    """training/gan/aug.py:7-36."""
    gen_images = gen_images.detach()
    all_images = torch.cat([P.augment_fn(images), gen_images], dim=0)
    return d_loss_with_penalty(P, D, options, images, gen_images, all_images, all_images)


def loss_G_fn(P, D, options, images, gen_images):
    """training/gan/aug.py:39-47."""
    return g_loss(D, options, gen_images)
