"""Mirror of models/gan/stylegan2/layers.py on the sm_100a kernels.

Same classes, constructor arguments, parameter / buffer names and initialisation as the reference, so that
state_dicts interchange; the arithmetic goes through ``contrad_b200.sg2_functional``.  Unlike the reference the
feature-map modules here take and return **NHWC** tensors (the layout of the tensor-core kernels); the NCHW
boundary is handled once by the discriminator / generator (``Rgb2Nhwc`` / ``Nhwc2Rgb``).  ``Upsample`` /
``Downsample`` act on the 3-channel NCHW image skip, as in the reference."""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .... import sg2_functional as SF
from .op import FusedLeakyReLU


class PixelNorm(nn.Module):
    """layers.py:15-20 (latent input, forward only on the training path)."""

    def forward(self, input):
        return SF.pixelnorm(input)


def make_kernel(k):
    """layers.py:23-31."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """layers.py:34-52; NCHW image skip."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        pad = self.pad
        return SF.UpFirDn.apply(input, self.kernel, self.factor, 1, (pad[0], pad[1], pad[0], pad[1]), None, False, False,
                                1.0, False)


class Downsample(nn.Module):
    """layers.py:55-73; NCHW image skip."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel)
        self.register_buffer("kernel", kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        pad = self.pad
        return SF.UpFirDn.apply(input, self.kernel, 1, self.factor, (pad[0], pad[1], pad[0], pad[1]), None, False, False,
                                1.0, False)


class Blur(nn.Module):
    """layers.py:76-93; NHWC feature maps.  `down` > 1 fuses the decimation of a following stride-2 1x1 convolution."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input, down=1, round_out=True):
        return SF.upfirdn2d_nhwc(input, self.kernel, down=down, pad=self.pad, round_out=round_out)


class EqualConv2d(nn.Module):
    """layers.py:96-129 (parameter container + the runtime weight scale; evaluated by ConvLayer)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def scaled_weight(self, cin_pad=None, mul=1.0, gemm=False):
        """weight * scale (* mul), input channels zero-padded to cin_pad; gemm=True: as the [Cout, k*k*Cin] matrix of the
        patch / 1x1 GEMMs.  Memoised per optimiser step (sg2_functional.weight_memo)."""
        cout, cin, k, _ = self.weight.shape
        alpha = self.scale * mul
        cp = cin if cin_pad is None else max(cin, cin_pad)

        def fwd(w):
            w = w * alpha
            if cp > cin:
                w = F.pad(w, (0, 0, 0, 0, 0, cp - cin))                  # zero weights for zero-padded input channels
            return w.permute(0, 2, 3, 1).reshape(cout, -1) if gemm else w

        def bwd(g):
            if gemm:
                g = g.reshape(cout, k, k, cp).permute(0, 3, 1, 2)
            return g[:, :cin] * alpha

        return SF.weight_memo(self.weight, ("scaled", cp, mul, gemm), lambda: SF.LinearMap.apply(self.weight, fwd, bwd))

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},"
                f" {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})")


class EqualLinear(nn.Module):
    """layers.py:132-160; [B, in_dim] inputs."""

    def __init__(self, in_dim, out_dim, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim))
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul
        self.bias_init = bias_init

    def forward(self, input):
        lr_mul, b0, scale = self.lr_mul, self.bias_init, self.scale
        bias = SF.weight_memo(self.bias, "scaled", lambda: SF.LinearMap.apply(
            self.bias, lambda b: b * lr_mul + b0, lambda g: g * lr_mul))
        w = SF.weight_memo(self.weight, "scaled", lambda: SF.LinearMap.apply(
            self.weight, lambda t: t * scale, lambda g: g * scale))
        if self.activation:
            return SF.BiasAct.apply(SF.MmNT.apply(input, w), bias, None, 0.2, 2 ** 0.5, True)
        return SF.MmNT.apply(input, w, bias)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ConvLayer(nn.Sequential):
    """layers.py:174-198: [Blur] + EqualConv2d(bias=False) + [FusedLeakyReLU], same child indices.

    forward(x NHWC, res=None, mul=1.0): `res` is added after the activation and `mul` scales the layer's output
    (the `(out + skip) / sqrt(2)` of ResBlock, discriminator.py:70-76, folded into the epilogue / the weights).
      3x3 stride 1      tcgen05 implicit-GEMM convolution
      Blur + 3x3 s2     FIR (cb200_upfirdn2d) -> 3x3 stride-2 patch matrix -> tcgen05 GEMM with K = 9*Cin
      Blur + 1x1 s2     FIR with decimation 2 -> GEMM
      1x1               GEMM (FromRGB: the 3 image channels arrive zero-padded to 32)"""

    def __init__(self, in_channel, out_channel, kernel_size, blur_kernel=[1, 3, 3, 1], downsample=False, activate=True):
        layers = []
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride, bias=False))
        if activate:
            layers.append(FusedLeakyReLU(out_channel))
        super().__init__(*layers)
        self.kernel_size, self.downsample, self.activate = kernel_size, downsample, activate
        if kernel_size not in (1, 3):
            raise NotImplementedError("ConvLayer: kernel sizes 1 and 3 (every StyleGAN2 discriminator layer)")

    def forward(self, x, res=None, mul=1.0):
        blur = self[0] if self.downsample else None
        conv = self[1] if self.downsample else self[0]
        act = self[len(self) - 1] if self.activate else None
        B, H, W, C = x.shape
        cout = conv.weight.shape[0]
        wmul = 1.0 if act is not None else mul
        if self.kernel_size == 3 and not self.downsample:
            y = SF.Conv3x3.apply(x, conv.scaled_weight(cin_pad=C, mul=wmul))
        elif self.kernel_size == 3:
            t = blur(x)                                            # [B, H+1, W+1, C]
            u = SF.PatchS2.apply(t, False)                         # [B, H/2, W/2, 9, C]
            y = SF.MmNT.apply(u.view(-1, 9 * C), conv.scaled_weight(cin_pad=C, mul=wmul, gemm=True))
            y = y.view(B, u.shape[1], u.shape[2], cout)
        else:
            t = blur(x, down=2) if self.downsample else x
            y = SF.MmNT.apply(t.reshape(-1, C), conv.scaled_weight(cin_pad=C, mul=wmul, gemm=True))
            y = y.view(t.shape[0], t.shape[1], t.shape[2], cout)
        if act is not None:
            return SF.BiasAct.apply(y, act.bias, res, act.negative_slope, act.scale * mul, True)
        if res is not None:
            return SF.Axpby.apply(y, res, 1.0, 1.0, 0.0)
        return y
