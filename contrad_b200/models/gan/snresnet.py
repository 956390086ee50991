"""Mirror of models/gan/snresnet.py (``D_SNResNet18``, the second discriminator architecture of the paper, SURVEY 8f
row f4) on the sm_100a kernels; state_dict keys / shapes are those of the reference after ``spectral_norm``
(``<layer>.weight_orig / .bias / .weight_u / .weight_v``).

Data flow: NCHW image -> SIMT first layer (x*2-1, 3->64, LeakyReLU) -> NHWC TF32 activations -> 8 BasicBlocks:
  3x3 stride-1 convs      tcgen05 implicit GEMM from the spectral-norm kernels' packed W/sigma (forward + data-gradient
                          packs, weight gradient back in the forward-pack layout)
  3x3 stride-2 convs      zero-pad top/left (cb200_upfirdn2d with a 1-tap FIR) -> 3x3/s2 patch matrix -> tcgen05 GEMM
  1x1 stride-2 shortcuts  decimation (cb200_upfirdn2d, down=2) -> tcgen05 GEMM with the bias in the epilogue
  residual sum + LeakyReLU  cb200_axpby + cb200_bias_act
-> 4x4 average pool (cb200_upfirdn2d, 4x4 box FIR, down=4) -> features [B, 512] -> the fused three-head GEMMs.
Spectral norm: one batched power iteration + packing launch for all 27 layers (as in D_SNDCGAN)."""
import torch
import torch.nn as nn

from ... import sg2_functional as SF
from ...functional import ConvFirstFn, ConvPackedFn, SNLayerSpec
from .base import BaseDiscriminator, SNConv2d


class BasicBlock(nn.Module):
    """models/gan/snresnet.py:21-44 (parameter container; evaluated by D_SNResNet18._backbone)."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = SNConv2d(in_planes, planes, 3, stride, 1, init="default")
        self.conv2 = SNConv2d(planes, planes, 3, 1, 1, init="default")
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(SNConv2d(in_planes, self.expansion * planes, 1, stride, 0, init="default"))
        self.stride = stride


class SNResNet(BaseDiscriminator):
    """models/gan/snresnet.py:46-89."""
    SLOPE = 0.1

    def __init__(self, block, num_blocks, n_classes=1, disable_sn=False, **kwargs):
        if disable_sn:
            raise NotImplementedError("disable_sn is not used by any reference config")
        self.in_planes = 64
        self.n_features = 512 * block.expansion
        super().__init__(self.n_features, n_classes=n_classes, head_init="default", **kwargs)
        self._feat_chw = (self.n_features, 1, 1)
        self.conv1 = SNConv2d(3, 64, 3, 1, 1, init="default")
        self.layer1 = self._make_layer(block, 64, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 128, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, 256, num_blocks[2], stride=2)
        self.layer4 = self._make_layer(block, 512, num_blocks[3], stride=2)
        self.register_buffer("_one", torch.ones(1, 1), persistent=False)              # 1-tap FIR: pad / decimate
        self.register_buffer("_box4", torch.full((4, 4), 1.0 / 16), persistent=False)  # F.avg_pool2d(out, 4)

    def _make_layer(self, block, planes, num_blocks, stride):
        layers = []
        for s in [stride] + [1] * (num_blocks - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    def _conv_layers(self):
        """(name, module) of every convolution in the order of the spectral-norm specs."""
        out = [("conv1", self.conv1)]
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4), start=1):
            for bi, blk in enumerate(layer):
                out.append(("layer%d.%d.conv1" % (li, bi), blk.conv1))
                out.append(("layer%d.%d.conv2" % (li, bi), blk.conv2))
                if len(blk.shortcut):
                    out.append(("layer%d.%d.shortcut.0" % (li, bi), blk.shortcut[0]))
        return out

    def _sn_specs(self):
        specs = []
        for name, m in self._conv_layers():
            if name == "conv1":
                kind = "conv_first"
            elif m.ks == 3 and m.stride == 1:
                kind = "conv"                # forward + data-gradient packs
            else:
                kind = "conv_plain"          # GEMM matrix only (3x3 stride 2 via patches, 1x1)
            specs.append(SNLayerSpec(name, m, kind, m.ks, m.stride))
        return specs + self._head_specs()

    def _backbone(self, holder, inputs, packs):
        names = [n for n, _ in self._conv_layers()]
        pack = dict(zip(names, packs))
        dgrad = holder["side"]["dgrad"]
        slope = SNResNet.SLOPE

        def conv(name, m, x):
            """Convolution WITHOUT its bias (added by the caller's bias_act / GEMM epilogue)."""
            B, H, W, C = x.shape
            if m.ks == 3 and m.stride == 1:
                return ConvPackedFn.apply(x, pack[name], dgrad[name])
            if m.ks == 3:                                                   # stride 2, padding 1
                t = SF.UpFirDn.apply(x, self._one, 1, 1, (1, 0, 1, 0), None, False, True, 1.0, False)   # [B, H+1, W+1, C]
                u = SF.PatchS2.apply(t, False)
                y = SF.MmNT.apply(u.view(-1, 9 * C), pack[name])
                return y.view(B, H // 2, W // 2, -1)
            t = x if m.stride == 1 else SF.UpFirDn.apply(x, self._one, 1, m.stride, (0, 0, 0, 0), None, False, True, 1.0, False)
            y = SF.MmNT.apply(t.reshape(-1, C), pack[name], m.bias)         # 1x1 shortcut: bias in the GEMM epilogue
            return y.view(t.shape[0], t.shape[1], t.shape[2], -1)

        out = ConvFirstFn.apply(inputs, pack["conv1"], self.conv1.bias, dgrad["conv1"], slope)
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4), start=1):
            for bi, blk in enumerate(layer):
                p = "layer%d.%d" % (li, bi)
                h = SF.BiasAct.apply(conv(p + ".conv1", blk.conv1, out), blk.conv1.bias, None, slope, 1.0, True)
                h = conv(p + ".conv2", blk.conv2, h)
                sc = conv(p + ".shortcut.0", blk.shortcut[0], out) if len(blk.shortcut) else out
                out = SF.BiasAct.apply(SF.Axpby.apply(h, sc, 1.0, 1.0, 0.0), blk.conv2.bias, None, slope, 1.0, True)
        out = SF.UpFirDn.apply(out, self._box4, 1, 4, (0, 0, 0, 0), None, False, True, 1.0, True)    # avg_pool2d(out, 4)
        return out.reshape(out.shape[0], -1)

    def _to_reference_order(self, features):
        return features


def D_SNResNet18(**kwargs):
    return SNResNet(BasicBlock, [2, 2, 2, 2], **kwargs)


def D_SNResNet34(**kwargs):
    return SNResNet(BasicBlock, [3, 4, 6, 3], **kwargs)
