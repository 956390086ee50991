"""Mirror of models/gan/stylegan2/discriminator.py:191-235 (``ResidualDiscriminatorP``, the StyleGAN2 discriminator of
ContraD) and of the ``BaseDiscriminator`` head contract (models/gan/base.py:79-150; plain ``nn.Linear`` heads - the
StyleGAN2 registry entries do not use spectral norm) on the sm_100a kernels.

NCHW image in [0,1] -> (x*2-1, zero-padded to 32 channels, NHWC) -> FromRGB GEMM -> ResBlocks -> minibatch stddev ->
last 3x3 convolution -> features [B, 4*4*C] in (h,w,c) order -> three MLP heads whose first-layer weights are
re-ordered from the reference's (c,h,w) flattening.  Every operator is double-differentiable (R1 penalty)."""
import math

import torch
from torch import nn

from .... import sg2_functional as SF
from .layers import ConvLayer


class FromRGB(ConvLayer):
    """discriminator.py:17-19."""

    def __init__(self, out_channel):
        super().__init__(3, out_channel, 1, activate=True)


class ResBlock(nn.Module):
    """discriminator.py:60-76: out = (conv2(conv1(x)) + skip(x)) / sqrt(2); the division is folded into conv2's
    activation gain and into the skip convolution's weights."""

    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3, activate=True)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, blur_kernel=blur_kernel, downsample=True, activate=True)
        self.skip = ConvLayer(in_channel, out_channel, 1, blur_kernel=blur_kernel, downsample=True, activate=False)

    def forward(self, input):
        out = self.conv1(input)
        skip = self.skip(input, mul=1 / math.sqrt(2))
        return self.conv2(out, res=skip, mul=1 / math.sqrt(2))


class TinyDiscriminator(nn.Module):
    """models/gan/base.py:14-35 with plain linears (parameter container; evaluated by _heads)."""

    def __init__(self, n_features, n_classes=1, d_hidden=128):
        super().__init__()
        if n_classes > 1:
            raise NotImplementedError("class-conditional heads are not on the ContraD hot path")
        self.n_features, self.n_classes, self.d_hidden = n_features, n_classes, d_hidden
        self.l1 = nn.Linear(n_features, d_hidden)
        self.l2 = nn.Linear(d_hidden, 1)


class ResidualDiscriminatorP(nn.Module):
    """discriminator.py:191-235 + models/gan/base.py:79-150."""

    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], small32=False, n_classes=1, d_hidden=128,
                 d_project=128, mlp_linear=False):
        super().__init__()
        if not mlp_linear:
            raise NotImplementedError("only mlp_linear=True (every registry architecture, models/gan/__init__.py) is built")
        if small32:
            channels = {4: 512, 8: 512, 16: 256, 32: 128}
        else:
            channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: int(256 * channel_multiplier),
                        128: int(128 * channel_multiplier), 256: int(64 * channel_multiplier),
                        512: int(32 * channel_multiplier), 1024: int(16 * channel_multiplier)}
        self.n_features = channels[4] * 4 * 4
        self.d_penul, self.n_classes, self.d_hidden, self.d_project = self.n_features, n_classes, d_hidden, d_project
        self.linear = TinyDiscriminator(self.n_features, n_classes=n_classes, d_hidden=d_hidden)
        self.projection = nn.Sequential(nn.Linear(self.n_features, d_hidden), nn.LeakyReLU(0.1, inplace=True),
                                        nn.Linear(d_hidden, d_project))
        self.projection2 = nn.Sequential(nn.Linear(self.n_features, d_hidden), nn.LeakyReLU(0.1, inplace=True),
                                         nn.Linear(d_hidden, d_project))
        layers = [FromRGB(channels[size])]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            layers.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.layers = nn.Sequential(*layers)
        self.last_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self._feat_chw = (channels[4], 4, 4)

    # ---- backbone
    def _features_hwc(self, input):
        """features [B, 4*4*C] in the kernels' (h, w, c) flattening."""
        x = SF.Rgb2Nhwc.apply(input, 32, 2.0, -1.0, True)                 # `input * 2. - 1.` (discriminator.py:229)
        out = self.layers(x)
        std = SF.Stddev.apply(out)                                         # discriminator.py:22-33
        cpad = (out.shape[-1] + 1 + 31) // 32 * 32
        out = SF.StddevConcat.apply(out, std, cpad, True)
        out = self.last_conv(out)
        return out.view(out.shape[0], -1)

    def _to_reference_order(self, features):
        c, h, w = self._feat_chw
        return features.view(-1, h, w, c).permute(0, 3, 1, 2).reshape(-1, self.n_features)

    def penultimate(self, input):
        return self._to_reference_order(self._features_hwc(input))

    # ---- heads
    def _hwc_weight(self, w):
        """first-layer head weight with its columns re-ordered from (c,h,w) to (h,w,c); once per optimiser step"""
        c, h, ww = self._feat_chw
        n = w.shape[0]
        return SF.weight_memo(w, "hwc", lambda: SF.LinearMap.apply(
            w, lambda t: t.view(n, c, h, ww).permute(0, 2, 3, 1).reshape(n, -1),
            lambda g: g.reshape(n, h, ww, c).permute(0, 3, 1, 2).reshape(n, -1)))

    def _mlp(self, feat, l1, l2):
        hid = SF.BiasAct.apply(SF.MmNT.apply(feat, self._hwc_weight(l1.weight)), l1.bias, None, 0.1, 1.0, True)
        n_out = l2.weight.shape[0]
        rows = (n_out + 31) // 32 * 32
        pad = rows - n_out
        w2 = SF.weight_memo(l2.weight, "rows", lambda: SF.LinearMap.apply(
            l2.weight, lambda t: torch.nn.functional.pad(t, (0, 0, 0, pad)), lambda g: g[:n_out]))
        b2 = SF.weight_memo(l2.bias, "rows", lambda: SF.LinearMap.apply(
            l2.bias, lambda t: torch.nn.functional.pad(t, (0, pad)), lambda g: g[:n_out]))
        out = SF.MmNT.apply(hid, w2, b2)
        return out if out.shape[1] == n_out else out[:, :n_out]

    def forward(self, inputs, y=None, penultimate=False, projection=False, projection2=False, finetuning=False,
                sg_linear=False):
        """models/gan/base.py:107-150."""
        if y is not None:
            raise NotImplementedError("class-conditional discriminators are not on the ContraD hot path")
        if finetuning:
            is_train = self.training
            self.eval()
            with torch.no_grad():
                features = self._features_hwc(inputs)
            features = features.detach()
            self.train(is_train)
        else:
            features = self._features_hwc(inputs)
        features_d = features.detach() if sg_linear else features
        output = self._mlp(features_d, self.linear.l1, self.linear.l2)
        aux = {}
        # the reference always evaluates both projection heads (their `* 0.` nuisance term adds exactly zero to the
        # output and to every gradient); here they run only when asked for
        if penultimate:
            aux["penultimate"] = self._to_reference_order(features)
        if projection:
            aux["projection"] = self._mlp(features, self.projection[0], self.projection[2])
        if projection2:
            aux["projection2"] = self._mlp(features, self.projection2[0], self.projection2[2])
        if aux:
            return output, aux
        return output
