"""Mirror of models/gan/stylegan2/op/fused_act.py:74-94 (`FusedLeakyReLU`, `fused_leaky_relu`) on
cb200_bias_act.  The public functions keep the reference's channel-at-dim-1 convention (NCHW / [B, C] inputs);
the discriminator / generator call the NHWC primitive (sg2_functional.BiasAct) directly."""
import torch
from torch import nn

from ..... import sg2_functional as SF


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """leaky_relu(input + bias.view(1, C, 1, ...), negative_slope) * scale with C = input.shape[1]."""
    if input.dim() == 2:
        return SF.BiasAct.apply(input, bias, None, negative_slope, scale, False)
    perm = (0,) + tuple(range(2, input.dim())) + (1,)
    inv = (0, input.dim() - 1) + tuple(range(1, input.dim() - 1))
    out = SF.BiasAct.apply(input.permute(*perm).contiguous(), bias, None, negative_slope, scale, False)
    return out.permute(*inv).contiguous()
