"""Mirror of models/gan/stylegan2/op/upfirdn2d.py:145-156 (`upfirdn2d(input, kernel, up, down, pad)` on NCHW
tensors) on cb200_upfirdn2d; differentiable to any order like the reference's UpFirDn2d / UpFirDn2dBackward pair
(:19-142)."""
from ..... import sg2_functional as SF


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return SF.UpFirDn.apply(input, kernel, up, down, (pad[0], pad[1], pad[0], pad[1]), None, False, False, 1.0, False)
