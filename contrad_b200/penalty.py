"""Mirror of penalty.py's public surface.  ``compute_penalty`` (gp / cr / bcr) is used only by the
std/aug/aug_both baselines (training/gan/std.py:27-30), never by mode='contrad' whose "penalty" is
L_dis (SURVEY discrepancy 3); the module must stay importable because train_gan.py:34-35 imports it for
its gin side effects."""
import torch


def no_penalty(images):
    return torch.zeros(1, device=images.device)


def compute_penalty(mode="none", **kwargs):
    if mode == "none":
        return no_penalty(kwargs["images"])
    raise NotImplementedError(
        "penalty %r belongs to the std/aug baselines, outside the ContraD hot path (SURVEY 2.1)" % mode)
