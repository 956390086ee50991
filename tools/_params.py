"""Random but valid per-sample parameter blocks for profiling / timing scripts (NOT the reference's sampling order - the
product's samplers live in contrad_b200.augment.layers; the test oracle is not imported outside tests/)."""
import torch


def random_simclr_params(batch, device="cuda", seed=0):
    """[11, B] block of cb200_augment_simclr_fwd (include/contrad_b200.h) and a jitter order."""
    g = torch.Generator().manual_seed(seed)
    u = lambda lo, hi: torch.rand(batch, generator=g) * (hi - lo) + lo
    sx, sy = u(0.45, 1.0), u(0.45, 1.0)
    bx, by = (torch.rand(batch, generator=g) * 2 - 1) * (1 - sx), (torch.rand(batch, generator=g) * 2 - 1) * (1 - sy)
    flip = (torch.rand(batch, generator=g) < 0.5).float() * 2 - 1
    cj_on = (torch.rand(batch, generator=g) < 0.8).float()
    gray_on = (torch.rand(batch, generator=g) < 0.2).float()
    p = torch.stack([sx, sy, bx, by, flip, cj_on, u(0.6, 1.4), u(-0.1, 0.1), u(0.6, 1.4), u(0.6, 1.4), gray_on])
    return p.contiguous().to(device), int(torch.rand((), generator=g) < 0.5)


def random_shift_flip_params(batch, max_pixels, width, device="cuda", seed=0):
    """[3, B] block of cb200_shift_flip_fwd."""
    g = torch.Generator().manual_seed(seed)
    sign = (torch.rand(batch, generator=g) < 0.5).float() * 2 - 1
    bias = torch.randint(-max_pixels, max_pixels + 1, (2, batch), generator=g).float() / (width / 2)
    return torch.cat([sign[None], bias]).contiguous().to(device)
