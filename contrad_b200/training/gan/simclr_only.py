"""Mirror of training/gan/simclr_only.py: the discriminator is trained by NT-Xent on two views of the real images only."""
import torch

from ...functional import GanGLossFn, RowNormalizeFn
from ...models.gan.base import projection
from ..criterion import nt_xent


def loss_D_fn(P, D, options, images, gen_images):
    """training/gan/simclr_only.py:9-21."""
    real_images = torch.cat([images, images], dim=0)
    views = RowNormalizeFn.apply(projection(D, P.augment_fn(real_images)))
    view1, view2 = torch.chunk(views, 2, dim=0)
    simclr_loss = nt_xent(view1, view2, temperature=P.temp, distributed=P.distributed)
    return simclr_loss, {
        "penalty": 0. * simclr_loss,
        "d_real": 0. * simclr_loss,
        "d_gen": 0. * simclr_loss,
    }


def loss_G_fn(P, D, options, images, gen_images):
    """training/gan/simclr_only.py:24-32."""
    return GanGLossFn.apply(D(P.augment_fn(gen_images)), options["loss"])
