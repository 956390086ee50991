"""Mirror of the reference's ``augment`` package surface (augment/__init__.py:13-28) for the ContraD hot
path: ``get_augment(mode)`` and the gin-configurable layer classes.  The ``simclr`` chain is ONE fused
sm_100a kernel (csrc/augment.cu) instead of ~120 ATen ops; per-sample parameters are still drawn on the
host side in the reference's exact numpy / torch RNG order, so identical seeds give
identical augmentations.

Out of scope for round 1: hfrt, gaussian, cutout, diffaug, simclr_hq (GaussianBlur) -
requesting them raises NotImplementedError rather than silently running something else."""
import gin
import torch.nn as nn

from .layers import (ColorJitterLayer, FusedSimCLR, HorizontalFlipLayer, NoAugment, RandomApply,  # noqa: F401
                     RandomColorGrayLayer, RandomResizeCropLayer)


def simclr():
    """augment/__init__.py:106-112 - same four stages, same constructor plumbing (gin supplies the args)."""
    return FusedSimCLR(
        RandomResizeCropLayer(),
        HorizontalFlipLayer(),
        RandomApply(ColorJitterLayer(), p=0.8),
        RandomApply(RandomColorGrayLayer(), p=0.2),
    )


_BUILT = {"none": NoAugment, "simclr": simclr}
_NEXT = ("gaussian", "hflip", "hfrt", "color_jitter", "cutout", "simclr_hq", "simclr_hq_cutout", "diffaug")


@gin.configurable("augment", whitelist=["fn"])
def get_augment(mode="none", **kwargs):
    if mode in _BUILT:
        return _BUILT[mode]()
    if mode in _NEXT:
        raise NotImplementedError(
            "augment mode %r is outside the round-1 hot path of contrad_b200 (SURVEY 8f); "
            "only 'simclr' and 'none' are built" % mode)
    raise KeyError(mode)
