"""Mirror of the reference's ``augment`` package surface (augment/__init__.py:13-28) for the ContraD hot
path: ``get_augment(mode)`` and the gin-configurable layer classes.  The ``simclr`` chain is ONE fused
sm_100a kernel (csrc/augment.cu) instead of ~120 ATen ops; per-sample parameters are still drawn on the
host side in the reference's exact numpy / torch RNG order, so identical seeds give
identical augmentations.

`simclr_hq` / `simclr_hq_cutout` (the README's StyleGAN2 recipe) append a separable Gaussian blur and a CutOut kernel.
`hfrt` / `gaussian` (row f4, the CR / bCR baselines' augmentations) are one gather kernel / one elementwise kernel,
`diffaug` (third_party/diffaug.py, policy 'color,cutout') a reduction + a pointwise launch.  Every mode of the reference's
registry (augment/__init__.py:14-25) is built; an unknown mode raises KeyError as in the reference."""
import gin
import torch.nn as nn

from .layers import (ColorJitterLayer, CutOut, DiffAugLayer, FusedSimCLR, FusedSimCLRHQ, Gaussian, GaussianBlur,  # noqa: F401
                     HorizontalFlipLayer, HorizontalFlipRandomCrop, NoAugment, RandomApply, RandomColorGrayLayer,
                     RandomCrop, RandomResizeCropLayer)


def simclr():
    """augment/__init__.py:106-112 - same four stages, same constructor plumbing (gin supplies the args)."""
    return FusedSimCLR(
        RandomResizeCropLayer(),
        HorizontalFlipLayer(),
        RandomApply(ColorJitterLayer(), p=0.8),
        RandomApply(RandomColorGrayLayer(), p=0.2),
    )


def simclr_hq():
    """augment/__init__.py:115-122."""
    return FusedSimCLRHQ(
        RandomResizeCropLayer(),
        HorizontalFlipLayer(),
        RandomApply(ColorJitterLayer(), p=0.8),
        RandomApply(RandomColorGrayLayer(), p=0.2),
        RandomApply(GaussianBlur(), p=0.5),
    )


def simclr_hq_cutout():
    """augment/__init__.py:125-133."""
    return FusedSimCLRHQ(
        RandomResizeCropLayer(),
        HorizontalFlipLayer(),
        RandomApply(ColorJitterLayer(), p=0.8),
        RandomApply(RandomColorGrayLayer(), p=0.2),
        RandomApply(GaussianBlur(), p=0.5),
        RandomApply(CutOut(), p=0.5),
    )


def diffaug():
    """augment/__init__.py:144-145."""
    return DiffAugLayer(policy="color,cutout")


_BUILT = {"none": NoAugment, "simclr": simclr, "simclr_hq": simclr_hq, "simclr_hq_cutout": simclr_hq_cutout,
          "cutout": CutOut, "hflip": HorizontalFlipLayer, "color_jitter": ColorJitterLayer,
          "gaussian": Gaussian, "hfrt": HorizontalFlipRandomCrop, "diffaug": diffaug}


@gin.configurable("augment", whitelist=["fn"])
def get_augment(mode="none", **kwargs):
    return _BUILT[mode]()
