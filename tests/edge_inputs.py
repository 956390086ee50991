"""TEST INFRASTRUCTURE: inputs at the edges of the augmentation chain's domain, shared by the oracle-vs-reference and the
kernel-vs-oracle tests."""
import torch


def degenerate_images(size=32, reps=4):
    """Constant images at the corners of the colour cube (hue breakpoints 0, 1/6, ... of the piecewise-linear colour
    wheel; saturation 0 for black / white / gray, where the hue is undefined), a near-gray, a 0/1 checkerboard and
    red / blue stripes (every bilinear tap straddles a jump)."""
    def const(rgb):
        return torch.tensor(rgb, dtype=torch.float32).view(1, 3, 1, 1).expand(1, 3, size, size)
    imgs = [const(c) for c in ([0, 0, 0], [1, 1, 1], [.5, .5, .5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [0, 1, 1],
                               [1, 0, 1], [.5, .5 + 1e-6, .5])]
    cb = torch.zeros(1, 3, size, size); cb[:, :, ::2, ::2] = 1; cb[:, :, 1::2, 1::2] = 1
    stripes = torch.zeros(1, 3, size, size); stripes[:, 0, :, ::2] = 1; stripes[:, 2, :, 1::2] = 1
    return torch.cat((imgs + [cb, stripes]) * reps).contiguous()
