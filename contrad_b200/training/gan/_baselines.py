"""Shared bodies of the baseline training modes (training/gan/{std,aug,aug_both}.py): same discriminator, same fused
GAN-loss kernel as mode='contrad'; they differ only in where `P.augment_fn` is applied."""
from ...functional import GanDLossFn, GanGLossFn
from ...penalty import compute_penalty

_D_LOSSES = ("nonsat", "wgan", "hinge", "lsgan")


def d_loss_with_penalty(P, D, options, images, gen_images, all_images, d_inputs):
    """std.py:7-36 / aug.py:7-36 / aug_both.py:7-37 once `all_images` (what the penalty sees) and `d_inputs` (what the
    discriminator sees) have been assembled."""
    if options["loss"] not in _D_LOSSES:
        raise NotImplementedError()
    n = images.size(0)
    d_all = D(d_inputs)
    d_real, d_gen = d_all[:n], d_all[n:]
    d_loss, means = GanDLossFn.apply(d_all, n, options["loss"], n)
    penalty = compute_penalty(P.penalty, P=P, D=D, all_images=all_images, images=images, gen_images=gen_images,
                              d_real=d_real, d_gen=d_gen, lbd=options["lbd"], lbd2=options["lbd2"])
    return d_loss, {"penalty": penalty, "d_real": means[0], "d_gen": means[1]}


def g_loss(D, options, d_inputs):
    """std.py:39-47."""
    return GanGLossFn.apply(D(d_inputs), options["loss"])
