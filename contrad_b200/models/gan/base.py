"""Mirror of models/gan/base.py: ``BaseDiscriminator`` (forward contract :107-150) with the three MLP
heads fused into sm_100a tensor-core GEMMs, plus the spectrally-normalised layer containers that keep
the reference's state_dict layout (``<layer>.weight_orig / .bias / .weight_u / .weight_v``)."""
from abc import ABCMeta, abstractmethod

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...functional import HeadsFn, SNLayerSpec, SNPackFn


def _unit(v, eps=1e-12):
    return v / v.norm().clamp_min(eps)


class _SNParams(nn.Module):
    """Parameter container of one spectrally-normalised layer: ``weight_orig``, ``bias`` (parameters),
    ``weight_u``, ``weight_v`` (buffers) - the keys torch.nn.utils.spectral_norm produces."""

    def __init__(self, weight_shape, init="sndcgan"):
        super().__init__()
        self.weight_shape = tuple(weight_shape)
        self.init = init
        fan = 1
        for s in weight_shape[1:]:
            fan *= s
        self.bias = nn.Parameter(torch.zeros(weight_shape[0]))
        self.weight_orig = nn.Parameter(torch.empty(*weight_shape))
        self.register_buffer("weight_u", torch.empty(weight_shape[0]))
        self.register_buffer("weight_v", torch.empty(fan))
        self.reset_parameters()

    def reset_parameters(self):
        """models/gan/sndcgan.py:130-148: N(0, 0.02) weights, zero bias, fresh unit-norm u / v
        (torch.nn.utils.spectral_norm draws them from N(0,1))."""
        with torch.no_grad():
            if self.init == "default":       # nn.Conv2d / nn.Linear defaults (models/gan/snresnet.py keeps them)
                fan = self.weight_v.numel()
                bound = 1.0 / fan ** 0.5
                self.weight_orig.uniform_(-bound, bound)
                self.bias.uniform_(-bound, bound)
            else:
                self.weight_orig.normal_(0.0, 0.02)
                self.bias.zero_()
            self.weight_u.copy_(_unit(torch.empty_like(self.weight_u).normal_(0, 1)))
            self.weight_v.copy_(_unit(torch.empty_like(self.weight_v).normal_(0, 1)))


class SNConv2d(_SNParams):
    def __init__(self, cin, cout, ks, stride, padding, init="sndcgan"):
        super().__init__((cout, cin, ks, ks), init=init)
        self.ks, self.stride, self.padding = ks, stride, padding

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=%d, padding=%d (spectral norm)" % (
            self.weight_shape[1], self.weight_shape[0], self.ks, self.stride, self.padding)


class SNLinear(_SNParams):
    def __init__(self, fin, fout, init="sndcgan"):
        super().__init__((fout, fin), init=init)

    def extra_repr(self):
        return "in_features=%d, out_features=%d (spectral norm)" % (self.weight_shape[1], self.weight_shape[0])


class TinyDiscriminator(nn.Module):
    """models/gan/base.py:14-35 (parameter container; evaluated inside HeadsFn)."""

    def __init__(self, n_features, n_classes=1, d_hidden=128, init="sndcgan"):
        super().__init__()
        if n_classes > 1:
            raise NotImplementedError("class-conditional heads are not on the ContraD hot path")
        self.n_features, self.n_classes, self.d_hidden = n_features, n_classes, d_hidden
        self.l1 = SNLinear(n_features, d_hidden, init=init)
        self.l2 = SNLinear(d_hidden, 1, init=init)


class LinearDiscriminator(nn.Module):
    """models/gan/base.py:36-51 (linear evaluation / fine-tuning scripts; a plain nn.Linear head, no kernel of this
    library is involved)."""

    def __init__(self, n_features, n_classes=1):
        super().__init__()
        self.n_features, self.n_classes = n_features, n_classes
        self.linear = nn.Linear(n_features, 1)
        if n_classes > 1:
            self.linear_y = nn.Embedding(n_classes, n_features)

    def forward(self, inputs, y=None):
        d = self.linear(inputs)
        if y is not None:
            d = d + (inputs * self.linear_y(y)).sum(1, keepdim=True)
        return d


class LinearWrapper(nn.Linear):
    """models/gan/base.py:54-59 (`test_lineval.py`)."""

    def forward(self, inputs, y=None):
        return super().forward(inputs)


class NullDiscriminator(nn.Module):
    """models/gan/base.py:62-68."""

    def forward(self, inputs, y=None):
        return inputs.sum(1, keepdim=True)


def projection(D, inputs):
    """models/gan/base.py:73-76 (used by training/gan/simclr_only.py): the projection head's output; `d.mean() * 0`
    keeps the unused linear head in the graph so that DDP finds a gradient for every parameter."""
    d, aux = D(inputs, projection=True)
    return aux["projection"] + d.mean() * 0


class BaseDiscriminator(nn.Module, metaclass=ABCMeta):
    def __init__(self, d_penul, n_classes=1, d_hidden=128, d_project=128, mlp_linear=False, head_init="sndcgan"):
        super().__init__()
        if not mlp_linear:
            raise NotImplementedError("only mlp_linear=True (every registry architecture, models/gan/__init__.py) is built")
        self.d_penul, self.n_classes, self.d_hidden, self.d_project = d_penul, n_classes, d_hidden, d_project
        self.linear = TinyDiscriminator(d_penul, n_classes=n_classes, d_hidden=d_hidden, init=head_init)
        self.projection = nn.Sequential(SNLinear(d_penul, d_hidden, init=head_init), nn.LeakyReLU(0.1, inplace=True),
                                        SNLinear(d_hidden, d_project, init=head_init))
        self.projection2 = nn.Sequential(SNLinear(d_penul, d_hidden, init=head_init), nn.LeakyReLU(0.1, inplace=True),
                                         SNLinear(d_hidden, d_project, init=head_init))

    # ---- hooks the concrete discriminator implements
    @abstractmethod
    def _sn_specs(self):
        """Ordered list of SNLayerSpec: conv layers, then the three head1 layers, then the three head2 layers."""

    @abstractmethod
    def _backbone(self, holder, inputs, packs):
        """features in the kernels' (h, w, c) flattening."""

    @abstractmethod
    def _to_reference_order(self, features):
        """(h, w, c) -> the reference's (c, h, w) flattening (for the `penultimate` aux output)."""

    def _head_specs(self):
        return [SNLayerSpec("linear.l1", self.linear.l1, "head1"),
                SNLayerSpec("projection.0", self.projection[0], "head1"),
                SNLayerSpec("projection2.0", self.projection2[0], "head1"),
                SNLayerSpec("linear.l2", self.linear.l2, "head2"),
                SNLayerSpec("projection.2", self.projection[2], "head2"),
                SNLayerSpec("projection2.2", self.projection2[2], "head2")]

    def _packs(self, which="all"):
        """Packed W/sigma matrices of the spectrally-normalised layers.  which = "conv" | "heads" | "all": the forward
        packs the backbone and the heads as TWO autograd nodes, the heads' node created AFTER the backbone's forward, so
        that in the backward pass the head layers' weight gradients are final right after the heads' backward (autograd
        runs the later-created node first) and their all-reduce can overlap the backbone's backward (engine.GradSync)."""
        specs = self._sn_specs()
        if which != "all":
            is_head = lambda s: s.kind in ("head1", "head2")
            specs = [s for s in specs if is_head(s) == (which == "heads")]
        holder = {"specs": specs, "feat_chw": self._feat_chw, "allow_strict": getattr(self, "_strict_capable", False)}
        packs = SNPackFn.apply(holder, self.training, *[s.module.weight_orig for s in specs])
        return holder, packs

    def early_gradient_parameters(self):
        """Parameters whose gradients are complete before the backbone's backward starts (see `_packs`)."""
        for m in (self.linear, self.projection, self.projection2):
            for p in m.parameters():
                yield p

    def penultimate(self, inputs):
        holder, packs = self._packs("conv")
        return self._to_reference_order(self._backbone(holder, inputs, packs))

    def forward(self, inputs, y=None, penultimate=False, projection=False, projection2=False,
                finetuning=False, sg_linear=False):
        """models/gan/base.py:107-150."""
        if y is not None:
            raise NotImplementedError("class-conditional discriminators are not on the ContraD hot path")
        if finetuning:
            # models/gan/base.py:112-118: the backbone runs in eval mode (no power iteration on its layers)
            is_train = self.training
            self.eval()
            with torch.no_grad():
                holder_c, packs_c = self._packs("conv")
                features = self._backbone(holder_c, inputs, packs_c)
            features = features.detach()
            self.train(is_train)
        else:
            holder_c, packs_c = self._packs("conv")
            features = self._backbone(holder_c, inputs, packs_c)
        holder, packs = self._packs("heads")
        wcat, w_l2, w_p1, w_p2 = packs[0:4]
        bcat = torch.cat([self.linear.l1.bias, self.projection[0].bias, self.projection2[0].bias])
        output, project, project2 = HeadsFn.apply(holder, bool(sg_linear), features, wcat, bcat,
                                                  w_l2, self.linear.l2.bias, w_p1, self.projection[2].bias,
                                                  w_p2, self.projection2[2].bias)
        aux = {}
        if penultimate:
            aux["penultimate"] = self._to_reference_order(features)
        if projection:
            aux["projection"] = project
        if projection2:
            aux["projection2"] = project2
        if aux:
            return output, aux
        return output

    def reset_parameters(self, root=None):
        root = self if root is None else root
        for m in root.modules():
            if isinstance(m, _SNParams):
                m.reset_parameters()
