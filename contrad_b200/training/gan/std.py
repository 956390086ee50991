"""Mirror of training/gan/std.py (plain GAN objective + `--penalty none|cr|bcr`)."""
import torch

from ._baselines import d_loss_with_penalty, g_loss


def loss_D_fn(P, D, options, images, gen_images):
    """training/gan/std.py:7-36."""
    gen_images = gen_images.detach()
    all_images = torch.cat([images, gen_images], dim=0)
    return d_loss_with_penalty(P, D, options, images, gen_images, all_images, all_images)


def loss_G_fn(P, D, options, images, gen_images):
    """training/gan/std.py:39-47."""
    return g_loss(D, options, gen_images)
