"""Mirror of models/gan/stylegan2/generator.py (``Generator`` and its parts) on the sm_100a kernels; same module
tree / parameter names / initialisation / RNG consumption order as the reference.

ModulatedConv2d is evaluated in its algebraically identical "modulate the activations" form instead of the
reference's per-sample grouped convolution (generator.py:52-82):
    conv(x, w * s[b,ci] * d[b,co]) = d[b,co] * conv(x * s[b,ci], w),   d = rsqrt(sum_ci s^2 * sum_k w^2 + eps)
so every sample shares ONE weight matrix and the convolution is a single tcgen05 GEMM for the whole batch:
    3x3            implicit-GEMM convolution (cb200_conv2d_nhwc_*)
    3x3 upsample   GEMM [B*H*W, Cin] x [Cin, 9*Cout] -> stride-2 scatter (= conv_transpose2d) -> FIR blur
    1x1 (ToRGB)    GEMM with the 3 output channels padded to 32
followed by one fused epilogue (demodulation, noise, bias, leaky ReLU).  Feature maps are NHWC; the image skip
is NCHW like the reference's output."""
import math

import torch
from torch import nn
from torch.nn import functional as F

from .... import sg2_functional as SF
from .... import staging
from .layers import Blur, EqualLinear, PixelNorm, Upsample
from .op import FusedLeakyReLU


class ModulatedConv2d(nn.Module):
    """generator.py:17-82."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            pad0 = (p + 1) // 2 + factor - 1
            pad1 = p // 2 + 1
            self.blur = Blur(blur_kernel, pad=(pad0, pad1), upsample_factor=factor)
        fan_in = in_channel * kernel_size ** 2
        self.scale = 1 / math.sqrt(fan_in)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        if kernel_size not in (1, 3) or (upsample and kernel_size != 3):
            raise NotImplementedError("ModulatedConv2d: 3x3 (optionally upsampling) and 1x1 kernels")

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample})")

    def forward(self, input, style, bias=None):
        """input NHWC [B or 1, H, W, Cin] -> (pre-demodulation output NHWC, demod [B, Cout] or None).
        The demodulation is applied by the caller's epilogue (ModEpilogue) together with noise / bias / activation;
        `bias` (1x1 kernels only: ToRGB) is added in the GEMM epilogue."""
        s = self.modulation(style)                                       # [B, Cin]
        scale = self.scale
        w = SF.weight_memo(self.weight, "scaled", lambda: SF.LinearMap.apply(        # [Cout, Cin, k, k]
            self.weight, lambda t: t[0] * scale, lambda g: (g * scale).unsqueeze(0)))
        cout, cin = w.shape[0], w.shape[1]
        demod = None
        if self.demodulate:
            # not memoised: `pow` keeps its input for the backward pass, and memo entries must survive a second
            # backward in the same optimiser step (weight_memo only holds scale / pad / permute results)
            wsq = w.pow(2).sum([2, 3])                                   # [Cout, Cin]
            demod = torch.rsqrt(SF.MmNT.apply(s * s, wsq) + self.eps)    # generator.py:58-60
        xm = SF.Modulate.apply(input, s, True)
        B, H, W, _ = xm.shape
        if self.kernel_size == 1:
            rows = (cout + 31) // 32 * 32
            w2 = SF.weight_memo(self.weight, "rows", lambda: SF.LinearMap.apply(
                self.weight, lambda t: F.pad(t[0].reshape(cout, cin) * scale, (0, 0, 0, rows - cout)),
                lambda g: (g[:cout] * scale).reshape(1, cout, cin, 1, 1)))
            b2 = None if bias is None else SF.weight_memo(bias, "rows", lambda: SF.LinearMap.apply(
                bias, lambda t: F.pad(t.reshape(-1), (0, rows - cout)), lambda g: g[:cout].reshape(bias.shape)))
            out = SF.MmNT.apply(xm.view(-1, cin), w2, b2).view(B, H, W, rows)
        elif self.upsample:
            # rows (kh, kw, co): conv_transpose2d taps
            wt = SF.weight_memo(self.weight, "taps", lambda: SF.LinearMap.apply(
                self.weight, lambda t: (t[0] * scale).permute(2, 3, 0, 1).reshape(9 * cout, cin),
                lambda g: (g.reshape(3, 3, cout, cin).permute(2, 3, 0, 1) * scale).unsqueeze(0)))
            v = SF.MmNT.apply(xm.view(-1, cin), wt).view(B, H, W, 9, cout)
            out = SF.PatchS2T.apply(v, False)                            # [B, 2H+1, 2W+1, Cout]
            out = self.blur(out, round_out=False)                        # [B, 2H, 2W, Cout]
        else:
            out = SF.Conv3x3.apply(xm, w)
        return out, demod


class NoiseInjection(nn.Module):
    """generator.py:85-94 (parameter container; the sum is fused into ModEpilogue)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class ConstantInput(nn.Module):
    """generator.py:97-105; returns the constant as a broadcastable NHWC tensor [1, size, size, C]."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.const = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.const.permute(0, 2, 3, 1).contiguous()


class StyleLayer(nn.Module):
    """generator.py:108-124."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        out, demod = self.conv(input, style)
        B, H, W, _ = out.shape
        if noise is None:
            noise = out.new_empty(B, 1, H, W).normal_()                 # generator.py:91-93 (device generator)
        elif noise.shape[0] != B:
            noise = noise.expand(B, -1, -1, -1)
        return SF.ModEpilogue.apply(out, demod, noise.contiguous(), self.noise.weight, self.activate.bias, True)


class ToRGB(nn.Module):
    """generator.py:127-149; returns the NCHW image skip."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out, _ = self.conv(input, style, bias=self.bias)                 # [B, H, W, 32], channels 3.. are zero
        res = self.upsample(skip) if skip is not None else None
        return SF.Nhwc2Rgb.apply(out, res, 1.0)


class Generator(nn.Module):
    """generator.py:152-290."""

    def __init__(self, size, style_dim=512, n_mlp=8, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 small32=False):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for i in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        if small32:
            self.channels = {4: 512, 8: 512, 16: 256, 32: 128}
        else:
            self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: int(256 * channel_multiplier),
                             128: int(128 * channel_multiplier), 256: int(64 * channel_multiplier),
                             512: int(32 * channel_multiplier), 1024: int(16 * channel_multiplier)}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyleLayer(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.layers = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.layers.append(StyleLayer(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.layers.append(StyleLayer(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    @property
    def device(self):
        return self.input.const.device

    def make_noise(self):
        noises = []
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            noises.append(torch.randn(1, 1, 2 ** res, 2 ** res, device=self.device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def sample_latent(self, num_samples):
        return torch.randn(num_samples, self.style_dim, device=self.device)

    def forward(self, input, return_latents=False, style_mix=0.9, input_is_latent=False, noise=None):
        """generator.py:233-290.  Style mixing picks, per sample and per layer, one of two mapped latents: instead of
        materialising `latents * mask + latent_mix * (1 - mask)` the per-layer rows are selected by index."""
        latent = self.style(input) if not input_is_latent else input
        if noise is None:
            noise = [None] * self.num_layers
        batch = input.size(0)
        per_layer = None                                 # None: every layer uses `latent`
        if latent.ndim >= 3:
            per_layer = [latent[:, i] for i in range(self.n_latent)]
        if self.training and (style_mix > 0):
            latent_mix = self.style(self.sample_latent(batch))
            n_latent = self.n_latent

            def draw():
                """generator.py:257-264 on the CPU generator, as in the reference: row l holds, per sample, the row of
                cat([latents_l, latent_mix]) that layer l uses (mask = layer_idx < mix_layer keeps `latents`)."""
                nomix_mask = torch.rand(batch) >= style_mix
                mix_layer = torch.randint(n_latent, (batch,))
                mix_layer = mix_layer.masked_fill(nomix_mask, n_latent)
                use_mix = torch.arange(n_latent)[:, None] >= mix_layer[None, :]
                return torch.arange(batch)[None, :] + batch * use_mix.long()

            # staged: under CUDA-graph capture the indices live in a static buffer that is refreshed per replay
            idx_all = staging.stage(draw, latent_mix.device, shape=(n_latent, batch), dtype=torch.int64)
            base = per_layer
            shared = torch.cat([latent, latent_mix], 0) if base is None else None
            per_layer = []
            for i in range(n_latent):
                both = shared if shared is not None else torch.cat([base[i], latent_mix], 0)
                per_layer.append(both.index_select(0, idx_all[i]))
        lat = (lambda i: latent) if per_layer is None else (lambda i: per_layer[i])

        out = self.input(latent)
        out = self.conv1(out, lat(0), noise=noise[0])
        skip = self.to_rgb1(out, lat(1))
        idx = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.layers[::2], self.layers[1::2], noise[1::2], noise[2::2],
                                                        self.to_rgbs):
            out = conv1(out, lat(idx), noise=noise1)
            out = conv2(out, lat(idx + 1), noise=noise2)
            skip = to_rgb(out, lat(idx + 2), skip)
            idx += 2
        image = SF.Axpby.apply(skip, None, 0.5, 0.0, 0.5)
        if not self.training:
            image = image.clamp(0, 1)
        if return_latents:
            if per_layer is None:
                latents = latent.unsqueeze(1).repeat(1, self.n_latent, 1)
            else:
                latents = torch.stack(per_layer, 1)
            return image, latents
        return image
