"""Mirror of models/gan/stylegan2/op/__init__.py:1-2 - the two native ops of the reference, API compatible."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d
