"""Mirror of models/gan/sndcgan.py: ``D_SNDCGAN`` (:69-148) on the sm_100a kernels and ``G_SNDCGAN``
(:13-66).  state_dict keys / shapes are those of the reference, so checkpoints interchange.

Discriminator data flow: NCHW image -> SIMT first layer (x*2-1 folded in) -> NHWC TF32 activations ->
six tcgen05 implicit-GEMM convolutions with fused bias + LeakyReLU(0.1) -> features [B, 4*4*512] in
(h,w,c) order -> one tensor-core GEMM for the three head MLPs.  Spectral norm: batched power iteration
+ packing kernels; sigma stays on the device."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import staging
from ...functional import GSNDCGANFn, SNDCGANBackboneFn, SNLayerSpec
from .base import BaseDiscriminator, SNConv2d


class G_SNDCGAN(nn.Module):
    """models/gan/sndcgan.py:13-66.  The torch modules below are parameter / buffer containers (same names,
    shapes and initialisation as the reference, so state_dicts interchange and
    nn.SyncBatchNorm.convert_sync_batchnorm works); the arithmetic is contrad_b200.functional.GSNDCGANFn."""

    def __init__(self, image_size, ngf=64, nz=128):
        super().__init__()
        self.image_size, self.ngf, self.nz = image_size, ngf, nz
        s_h, s_w, nc = image_size
        if nc != 3:
            raise NotImplementedError("RGB outputs only")
        self.s_hb, self.s_wb = s_h // 8, s_w // 8
        self.linear = nn.Linear(nz, ngf * 8 * self.s_hb * self.s_wb)
        self.norm_init = nn.BatchNorm2d(ngf * 8 * self.s_hb * self.s_wb)
        self.main = nn.Sequential(
            nn.ConvTranspose2d(ngf * 8, ngf * 4, 4, 2, 1), nn.BatchNorm2d(ngf * 4), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(ngf * 4, ngf * 2, 4, 2, 1), nn.BatchNorm2d(ngf * 2), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(ngf * 2, ngf, 4, 2, 1), nn.BatchNorm2d(ngf), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(ngf, nc, 3, 1, 1), nn.Tanh())
        self.reset_parameters()

    def _bns(self):
        return [self.norm_init, self.main[1], self.main[4], self.main[7]]

    def forward(self, z):
        bns = self._bns()
        holder = {"training": self.training,
                  "sync": self.training and any(isinstance(b, nn.SyncBatchNorm) for b in bns)
                  and torch.distributed.is_available() and torch.distributed.is_initialized()
                  and torch.distributed.get_world_size() > 1,
                  "bn_states": [(b.running_mean, b.running_var) for b in bns],
                  "s_hb": self.s_hb, "s_wb": self.s_wb}
        if self.training:
            torch._foreach_add_([b.num_batches_tracked for b in bns], 1)       # one launch for the four counters
        c = self.main
        return GSNDCGANFn.apply(holder, z, self.linear.weight, self.linear.bias, bns[0].weight, bns[0].bias,
                                c[0].weight, c[0].bias, bns[1].weight, bns[1].bias,
                                c[3].weight, c[3].bias, bns[2].weight, bns[2].bias,
                                c[6].weight, c[6].bias, bns[3].weight, bns[3].bias,
                                c[9].weight, c[9].bias)

    def sample_latent(self, n_samples):
        """models/gan/sndcgan.py:50-52: U(-1,1) drawn on the CPU generator (same stream as the reference), staged
        through a ring of pinned buffers so the host->device copy does not block the host."""
        device = next(self.parameters()).device
        nz = self.nz
        return staging.stage(lambda out: out.uniform_(-1, 1), device, shape=(n_samples, nz), fill=True)

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, (nn.ConvTranspose2d, nn.Linear)):
                nn.init.normal_(m.weight.data, 0.0, 0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias.data, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight.data, 1.0)
                nn.init.constant_(m.bias.data, 0.0)


class D_SNDCGAN(BaseDiscriminator):
    _strict_capable = True        # the strict-precision ("3xTF32") generator step is wired through this backbone and heads

    def __init__(self, image_size, ndf=64, n_classes=1, normalize=False, disable_sn=False, mlp_linear=False,
                 d_hidden=128):
        if normalize or disable_sn:
            raise NotImplementedError("normalize / disable_sn variants are not used by any reference config")
        if ndf != 64:
            raise NotImplementedError("the first-layer kernel is specialised for ndf=64 (the registry value)")
        s_h, s_w, nc = image_size
        if nc != 3:
            raise NotImplementedError("RGB inputs only")
        self.image_size, self.ndf = image_size, ndf
        self.s_hb, self.s_wb = s_h // 8, s_w // 8
        self.n_features = ndf * 8 * self.s_hb * self.s_wb
        super().__init__(self.n_features, n_classes=n_classes, d_hidden=d_hidden, mlp_linear=mlp_linear)
        self._feat_chw = (ndf * 8, self.s_hb, self.s_wb)
        act = lambda: nn.LeakyReLU(0.1, inplace=True)     # placeholders keep the reference's `main.<2i>` indices
        self.main = nn.Sequential(
            SNConv2d(nc, ndf, 3, 1, 1), act(),
            SNConv2d(ndf, ndf * 2, 4, 2, 1), act(), SNConv2d(ndf * 2, ndf * 2, 3, 1, 1), act(),
            SNConv2d(ndf * 2, ndf * 4, 4, 2, 1), act(), SNConv2d(ndf * 4, ndf * 4, 3, 1, 1), act(),
            SNConv2d(ndf * 4, ndf * 8, 4, 2, 1), act(), SNConv2d(ndf * 8, ndf * 8, 3, 1, 1), act())

    def _sn_specs(self):
        specs = []
        for i in range(0, len(self.main), 2):
            m = self.main[i]
            specs.append(SNLayerSpec("main.%d" % i, m, "conv_first" if i == 0 else "conv", m.ks, m.stride))
        return specs + self._head_specs()

    def _backbone(self, holder, inputs, packs):
        convs = [self.main[i] for i in range(0, len(self.main), 2)]
        wb = []
        for pack, m in zip(packs, convs):
            wb += [pack, m.bias]
        return SNDCGANBackboneFn.apply(holder, inputs, *wb)

    def _to_reference_order(self, features):
        c, h, w = self._feat_chw
        return features.view(-1, h, w, c).permute(0, 3, 1, 2).reshape(-1, self.n_features)
