"""Restatement of the two kornia.filters entry points the reference imports
(augment/__init__.py:4,74-76): a normalised outer-product Gaussian and a depthwise
'reflect'-padded correlation.  kornia is unpinned in the reference (environment.yml:25) and
absent here, so this is *parity unpinned* (SURVEY 8c)."""
import torch
import torch.nn.functional as F


def _gaussian1d(ksize, sigma):
    x = torch.arange(ksize, dtype=torch.float32) - ksize // 2
    if ksize % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2.0 * float(sigma) ** 2))
    return g / g.sum()


def get_gaussian_kernel2d(kernel_size, sigma, force_even=False):
    ky, kx = kernel_size
    sy, sx = sigma
    return torch.outer(_gaussian1d(ky, sy), _gaussian1d(kx, sx))


def filter2D(input, kernel, border_type="reflect", normalized=False):
    b, c, h, w = input.shape
    k = kernel.to(input)
    if k.dim() == 2:
        k = k.unsqueeze(0)
    kh, kw = k.shape[-2:]
    if normalized:
        k = k / k.sum(dim=(-2, -1), keepdim=True)
    pad = (kw // 2, kw - 1 - kw // 2, kh // 2, kh - 1 - kh // 2)
    x = F.pad(input, pad, mode=border_type)
    weight = k.expand(c, 1, kh, kw) if k.shape[0] == 1 else k.view(-1, 1, kh, kw)
    return F.conv2d(x, weight.contiguous(), groups=c)


filter2d = filter2D
