"""Mirror of penalty.py: ``compute_penalty`` for the std / aug / aug_both baselines (training/gan/std.py:27-30).
mode='contrad' never calls it (its "penalty" is L_dis, SURVEY discrepancy 3) but train_gan.py:34-35 imports the module
for its gin side effects.

`cr` / `bcr` (row f4) are one more pass of the hot path's discriminator over `P.augment_fn(...)` (the `hfrt` gather kernel
for the paper's CR baselines) plus N-element arithmetic.  `gp` differentiates the discriminator's backward pass; the
SNDCGAN / SNResNet discriminators of this library are single autograd nodes (first order only), so it is refused loudly -
the second-order operator families exist for the StyleGAN2 discriminator only (R1, training/gan/stylegan2.py)."""
import torch


def no_penalty(images):
    """penalty.py:12-13."""
    return torch.zeros(1, device=images.device)


def gradient_penalty(D, images, gen_images, lbd):
    """penalty.py:16-44."""
    raise NotImplementedError(
        "penalty 'gp' needs the double backward of the discriminator; contrad_b200 builds second-order operators for "
        "the StyleGAN2 discriminator only (R1) - outside the ContraD hot path (SURVEY 2.1)")


def consistency(D, P, images, d_real, lbd):
    """penalty.py:47-49 (CR)."""
    d_aug = D(P.augment_fn(images))
    return lbd * ((d_real - d_aug) ** 2).mean()


def balanced_consistency(D, P, all_images, d_real, d_gen, lbd, lbd2):
    """penalty.py:52-60 (bCR)."""
    d_aug_all = D(P.augment_fn(all_images))
    n = all_images.size(0) // 2
    d_aug_real, d_aug_gen = d_aug_all[:n], d_aug_all[n:]
    d_reg_real = ((d_real - d_aug_real) ** 2).mean()
    d_reg_gen = ((d_gen - d_aug_gen) ** 2).mean()
    return lbd * d_reg_real + lbd2 * d_reg_gen


_ACCEPTED = {
    "none": (no_penalty, ("images",)),
    "gp": (gradient_penalty, ("D", "images", "gen_images", "lbd")),
    "cr": (consistency, ("D", "P", "images", "d_real", "lbd")),
    "bcr": (balanced_consistency, ("D", "P", "all_images", "d_real", "d_gen", "lbd", "lbd2")),
}


def compute_penalty(mode="none", **kwargs):
    """penalty.py:63-71: dispatch by name, passing only the keyword arguments the penalty accepts
    (utils.call_with_accepted_args)."""
    fn, names = _ACCEPTED[mode]
    return fn(**{k: kwargs[k] for k in names if k in kwargs})
