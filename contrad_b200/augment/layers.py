"""Layer classes of the SimCLR chain.  Each class keeps the reference's constructor signature (they are
bound by gin by CLASS NAME, configs/defaults/augment.gin) and can run alone; ``FusedSimCLR`` runs the
whole chain in one kernel launch."""
import math
import numbers

import gin
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from .. import staging
from ..functional import (AugmentSimCLRFn, AugmentSimCLRMixedFn, CutOutFn, DiffAugFn, GaussianBlurFn, NoiseClampFn,
                          ShiftFlipFn)

_N_FIELDS = 11   # sx, sy, bx, by, flip, cj_on, contrast, hue, sat, val, gray_on


class _ShapeOnly(object):
    """Stand-in for a tensor where only `.shape` is consulted (host samplers captured by the staging recorder
    must not keep a device tensor alive)."""

    def __init__(self, shape):
        self.shape = tuple(shape)


def _identity_block(batch, device):
    p = torch.zeros(_N_FIELDS, batch, device=device)
    p[0] = 1.0; p[1] = 1.0; p[4] = 1.0; p[6] = 1.0; p[8] = 1.0; p[9] = 1.0
    return p


@gin.configurable
class NoAugment(nn.Module):
    def forward(self, input):
        return input


@gin.configurable(whitelist=["sigma"])
class Gaussian(nn.Module):
    """augment/__init__.py:40-49: additive Gaussian noise, clamped to [0, 1].  The noise is torch.randn_like (the
    reference's random stream); the scale-add-clamp and its gradient mask are one kernel each."""

    def __init__(self, sigma):
        super().__init__()
        self.sigma = sigma

    def forward(self, input):
        return NoiseClampFn.apply(input, torch.randn_like(input), self.sigma)


class _ShiftFlip(nn.Module):
    """Shared body of HorizontalFlipRandomCrop / RandomCrop (augment/spatial.py:14-67): theta = [[sign, 0, bias_x],
    [0, 1, bias_y]] with bias = randint(-max_pixels, max_pixels + 1) / (width / 2), sampled by nearest-neighbour
    grid_sample.  One gather kernel instead of affine_grid + grid_sample."""

    _flip = False

    def __init__(self, max_pixels, width, padding_mode):
        super().__init__()
        if padding_mode not in ("zeros", "border", "reflection"):
            raise ValueError("padding_mode must be 'zeros', 'border' or 'reflection' (F.grid_sample), got %r"
                             % (padding_mode,))
        self.max_pixels = max_pixels
        self.width = width
        self.register_buffer("_eye", torch.eye(2, 3))      # state_dict compatibility
        self.padding_mode = padding_mode

    def sample(self, input):
        """[3, N] = {sign, bias_x, bias_y}; device draws in the reference order (sign first, spatial.py:31-33)."""
        n, dev = input.size(0), input.device
        params = torch.empty(3, n, device=dev)
        if self._flip:
            params[0] = torch.bernoulli(torch.ones(n, device=dev) * 0.5) * 2 - 1
        else:
            params[0] = 1.0
        r_bias = torch.randint(-self.max_pixels, self.max_pixels + 1, (n, 2), device=dev).float() / (self.width / 2)
        params[1:3] = r_bias.t()
        return params

    def forward(self, input):
        return ShiftFlipFn.apply(input, self.sample(input), self.padding_mode)


@gin.configurable
class HorizontalFlipRandomCrop(_ShiftFlip):
    """augment/spatial.py:14-40 (`--aug hfrt`)."""
    _flip = True


@gin.configurable
class RandomCrop(_ShiftFlip):
    """augment/spatial.py:43-67."""
    _flip = False


@gin.configurable
class RandomResizeCropLayer(nn.Module):
    """Inception crop parameters (augment/spatial.py:96-148)."""

    def __init__(self, scale, ratio=(3. / 4., 4. / 3.)):
        super().__init__()
        self.register_buffer("_eye", torch.eye(2, 3))      # kept for state_dict compatibility
        self.scale = scale
        self.ratio = ratio

    def sample(self, inputs):
        """numpy draws in the reference order (spatial.py:119-136). Returns float32 [4, B] on the host."""
        n, _, width, height = inputs.shape
        area = height * width
        target_area = np.random.uniform(*self.scale, n * 10) * area
        log_ratio = (math.log(self.ratio[0]), math.log(self.ratio[1]))
        aspect = np.exp(np.random.uniform(*log_ratio, n * 10))
        w = np.round(np.sqrt(target_area * aspect))
        h = np.round(np.sqrt(target_area / aspect))
        keep = (0 < w) * (w <= width) * (0 < h) * (h <= height)
        w, h = w[keep], h[keep]
        if len(w) > n:
            sel = np.random.choice(len(w), n, replace=False)
            w, h = w[sel], h[sel]
        k = len(w)
        bias_w = np.random.randint(w - width, width - w + 1) / width
        bias_h = np.random.randint(h - height, height - h + 1) / height
        block = np.zeros((4, n), dtype=np.float32)
        block[0], block[1] = 1.0, 1.0
        block[0, :k] = w / width
        block[1, :k] = h / height
        block[2, :k] = bias_w
        block[3, :k] = bias_h
        return torch.from_numpy(block)

    def forward(self, inputs):
        p = _identity_block(inputs.shape[0], inputs.device)
        shape = _ShapeOnly(inputs.shape)
        p[0:4] = staging.stage(lambda: self.sample(shape), inputs.device)
        return AugmentSimCLRFn.apply(inputs, p, 0)


@gin.configurable
class HorizontalFlipLayer(nn.Module):
    """augment/spatial.py:70-93."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_eye", torch.eye(2, 3))

    def sample(self, inputs):
        n = inputs.size(0)
        return torch.bernoulli(torch.ones(n, device=inputs.device) * 0.5) * 2 - 1

    def forward(self, inputs):
        p = _identity_block(inputs.shape[0], inputs.device)
        p[4] = self.sample(inputs)
        return AugmentSimCLRFn.apply(inputs, p, 0)


@gin.configurable
class ColorJitterLayer(nn.Module):
    """augment/color_jitter.py:15-78 (ranges validated like the reference's _check_input)."""

    def __init__(self, brightness, contrast, saturation, hue):
        super().__init__()
        self.brightness = self._range(brightness, "brightness")
        self.contrast = self._range(contrast, "contrast")
        self.saturation = self._range(saturation, "saturation")
        self.hue = self._range(hue, "hue", center=0, bound=(-0.5, 0.5), clip_first_on_zero=False)

    @staticmethod
    def _range(value, name, center=1, bound=(0, float("inf")), clip_first_on_zero=True):
        if isinstance(value, numbers.Number):
            if value < 0:
                raise ValueError("If {} is a single number, it must be non negative.".format(name))
            value = [center - value, center + value]
            if clip_first_on_zero:
                value[0] = max(value[0], 0)
        elif isinstance(value, (tuple, list)) and len(value) == 2:
            if not bound[0] <= value[0] <= value[1] <= bound[1]:
                raise ValueError("{} values should be between {}".format(name, bound))
        else:
            raise TypeError("{} should be a single number or a list/tuple with lenght 2.".format(name))
        if value[0] == value[1] == center:
            value = None
        return value

    @staticmethod
    def draw_order():
        """color_jitter.py:65-70: one host draw per call decides [contrast, hsv] (0) or [hsv, contrast] (1)."""
        return 0 if np.random.rand() > 0.5 else 1

    def sample(self, inputs, order=None):
        """Returns (order, contrast, hue, sat, val) drawn like color_jitter.py:44-75."""
        n = inputs.size(0)
        if order is None:
            order = self.draw_order()

        def draw_contrast():
            if self.contrast:
                return inputs.new_empty(n, 1, 1, 1).uniform_(*self.contrast).view(n)
            return inputs.new_ones(n)

        def draw_hsv():
            f_h = inputs.new_zeros(n, 1, 1)
            f_s = inputs.new_ones(n, 1, 1)
            f_v = inputs.new_ones(n, 1, 1)
            if self.hue:
                f_h.uniform_(*self.hue)
            if self.saturation:
                f_s = f_s.uniform_(*self.saturation)
            if self.brightness:
                f_v = f_v.uniform_(*self.brightness)
            return f_h.view(n), f_s.view(n), f_v.view(n)

        if order == 0:
            f_c = draw_contrast()
            f_h, f_s, f_v = draw_hsv()
        else:
            f_h, f_s, f_v = draw_hsv()
            f_c = draw_contrast()
        return order, f_c, f_h, f_s, f_v

    def forward(self, inputs):
        p = _identity_block(inputs.shape[0], inputs.device)
        order, p[6], p[7], p[8], p[9] = self.sample(inputs)
        p[5] = 1.0
        return AugmentSimCLRFn.apply(inputs, p, order)


@gin.configurable
class RandomColorGrayLayer(nn.Module):
    """augment/__init__.py:81-91."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_weight", torch.tensor([[0.299, 0.587, 0.114]]).view(1, 3, 1, 1))

    def forward(self, inputs):
        p = _identity_block(inputs.shape[0], inputs.device)
        p[10] = 1.0
        return AugmentSimCLRFn.apply(inputs, p, 0)


class RandomApply(nn.Module):
    """augment/__init__.py:94-103: per-sample Bernoulli(p) mask blending x and fn(x)."""

    def __init__(self, fn, p):
        super().__init__()
        self.fn = fn
        self.p = p

    def sample(self, inputs):
        return torch.bernoulli(inputs.new_full((inputs.size(0),), self.p))

    def forward(self, inputs):
        mask = self.sample(inputs).view(-1, 1, 1, 1)
        return inputs * (1 - mask) + self.fn(inputs) * mask


class FusedSimCLR(nn.Sequential):
    """nn.Sequential(RRC, HFlip, RandomApply(CJ), RandomApply(Gray)) whose forward is ONE kernel.
    The child modules hold the gin-configured hyper-parameters and draw the random numbers (in the
    reference order); the arithmetic is cb200_augment_simclr_fwd/bwd."""

    def _draw_cfg(self):
        """{p_flip, p_jitter, p_gray, contrast, hue, saturation, value ranges} of the child modules (gin-configured)."""
        apply_cj, apply_gray = self[2], self[3]
        cj = apply_cj.fn
        rng = lambda r, centre: (centre, centre) if r is None else (r[0], r[1])
        return ((0.5, apply_cj.p, apply_gray.p) + rng(cj.contrast, 1.0) + rng(cj.hue, 0.0) + rng(cj.saturation, 1.0)
                + rng(cj.brightness, 1.0))

    def sample_params(self, inputs):
        """The parameter block of one call.

        Eager calls draw every factor with the reference's own torch / numpy calls in the reference's order (a given
        seed consumes the RNG streams exactly like `augment.simclr()` does: tests/test_host_logic.py replays the
        fixtures' seeds).  Under `staging.Recorder` (the step is being turned into a CUDA graph) the ~20 small ATen ops
        of that sequence would each become a graph node, so the block is built by ONE device draw of [7, n] uniforms
        mapped by cb200_augment_simclr_params (bernoulli(p) = u < p, uniform_(lo, hi) = lo + (hi - lo) u: the same
        distributions; seed parity with the eager sequence is given up there - DESIGN 6) plus the host-staged crop
        boxes and jitter order (row 11, kernel order = -1)."""
        rrc, flip, apply_cj, apply_gray = self[0], self[1], self[2], self[3]
        n, dev = inputs.shape[0], inputs.device
        shape = _ShapeOnly(inputs.shape)
        if staging.recording():
            boxes = staging.stage(lambda: rrc.sample(shape), dev, shape=(4, n))
            order_src = staging.stage(lambda: torch.tensor([float(apply_cj.fn.draw_order())]), dev, shape=(1,))
            u = torch.rand(7, n, device=dev)
            return K.augment_simclr_params(boxes, u, order_src, _N_FIELDS + 1, self._draw_cfg()), -1
        p = torch.empty(_N_FIELDS, n, device=dev)
        p[0:4] = staging.stage(lambda: rrc.sample(shape), dev, shape=(4, n))
        p[4] = flip.sample(inputs)
        p[5] = apply_cj.sample(inputs)
        order, p[6], p[7], p[8], p[9] = apply_cj.fn.sample(inputs)
        p[10] = apply_gray.sample(inputs)
        return p, order

    def forward(self, inputs):
        if inputs.dim() != 4 or inputs.shape[1] != 3:
            raise ValueError("FusedSimCLR expects [B,3,H,W] images, got %s" % (tuple(inputs.shape),))
        params, order = self.sample_params(inputs)
        return AugmentSimCLRFn.apply(inputs, params, order)

    def forward_views(self, images_u8, reps=1, extra=None):
        """Row f3: `self(torch.cat([images_u8.float() / 255] * reps + [extra]))` in ONE launch that reads
        the dataset's uint8 bytes directly - ToTensor (datasets.py:10-21), the fp32 upload (train_gan.py:153-154) and
        the concatenation (training/gan/contrad.py:38-40) are folded into the kernel.  Random draws are made for the
        whole `reps * n + len(extra)` batch in the reference order, so a given seed produces the views the reference
        produces on the converted, concatenated batch.  Only `extra` (fp32, e.g. G(z)) is differentiable."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[1] != 3:
            raise ValueError("forward_views expects uint8 [n,3,H,W] images, got %s %s"
                             % (images_u8.dtype, tuple(images_u8.shape)))
        if extra is not None and (extra.dtype != torch.float32 or tuple(extra.shape[1:]) != tuple(images_u8.shape[1:])):
            raise ValueError("extra must be float32 [m,3,H,W] of the same image size, got %s %s"
                             % (extra.dtype, tuple(extra.shape)))
        n_views = int(reps) * images_u8.shape[0]
        total = n_views + (0 if extra is None else extra.shape[0])
        # shape / device / dtype carrier for the samplers (no B x 3 x H x W allocation)
        carrier = torch.empty(1, device=images_u8.device, dtype=torch.float32).expand(total, *images_u8.shape[1:])
        params, order = self.sample_params(carrier)
        return AugmentSimCLRMixedFn.apply(images_u8, n_views, extra, params, order)


def gaussian_taps(kernel_size, sigma):
    """The normalised 1-D Gaussian whose outer product is kornia's `get_gaussian_kernel2d((k, k), (sigma, sigma))`."""
    x = torch.arange(kernel_size, dtype=torch.float32) - kernel_size // 2
    g = torch.exp(-x.pow(2.0) / (2.0 * float(sigma) ** 2))
    return g / g.sum()


@gin.configurable
class GaussianBlur(nn.Module):
    """augment/__init__.py:52-78: kernel size 2*floor((H/10)/2)+1, ONE sigma ~ U(sigma_range) per call (numpy), dense
    outer-product Gaussian with 'reflect' padding - evaluated separably by cb200_gaussian_blur."""

    def __init__(self, sigma_range):
        super().__init__()
        self.sigma_range = sigma_range

    @staticmethod
    def kernel_size(height):
        return int((height // 10) / 2) * 2 + 1

    def taps(self, inputs):
        k = self.kernel_size(inputs.shape[2])
        lo, hi = self.sigma_range
        return staging.stage(lambda: gaussian_taps(k, np.random.uniform(lo, hi)), inputs.device, shape=(k,))

    def forward(self, inputs, on=None):
        if on is None:
            on = inputs.new_ones(inputs.shape[0])
        return GaussianBlurFn.apply(inputs, self.taps(inputs), on)


@gin.configurable
class CutOut(nn.Module):
    """augment/spatial.py:151-181."""

    def __init__(self, length):
        super().__init__()
        if length % 2 == 0:
            raise ValueError("Currently CutOut only accepts odd lengths: length % 2 == 1")
        self.length = length
        self.register_buffer("_weight", torch.ones(1, 1, self.length))     # state_dict compatibility
        self._padding = (length - 1) // 2

    def sample(self, inputs):
        n, _, h, w = inputs.shape
        h_center = torch.randint(h, (n, 1), device=inputs.device)
        w_center = torch.randint(w, (n, 1), device=inputs.device)
        return h_center.view(n).float(), w_center.view(n).float()

    def forward(self, inputs, on=None):
        n = inputs.shape[0]
        params = torch.empty(3, n, device=inputs.device)
        params[0] = 1.0 if on is None else on
        params[1], params[2] = self.sample(inputs)
        return CutOutFn.apply(inputs, params, self.length)


class FusedSimCLRHQ(FusedSimCLR):
    """`simclr_hq` / `simclr_hq_cutout` (augment/__init__.py:115-133): the four fused stages, then
    RandomApply(GaussianBlur, .5) and optionally RandomApply(CutOut, .5).  Random draws in the reference order:
    ... gray mask, blur mask (device), sigma (numpy), cutout mask (device), centres (device)."""

    def forward(self, inputs):
        return self._tail(FusedSimCLR.forward(self, inputs), inputs)

    def forward_views(self, images_u8, reps=1, extra=None):
        out = FusedSimCLR.forward_views(self, images_u8, reps, extra)
        return self._tail(out, out)

    def _tail(self, out, inputs):
        apply_blur = self[4]
        on = apply_blur.sample(inputs)
        out = apply_blur.fn(out, on=on)
        if len(self) > 5:
            apply_cut = self[5]
            on = apply_cut.sample(inputs)
            out = apply_cut.fn(out, on=on)
        return out


class DiffAugLayer(nn.Module):
    """augment/__init__.py:136-142 + third_party/diffaug.py: DiffAugment(inputs, policy) with policy a comma-separated
    subset of 'color', 'translation', 'cutout'.  The per-sample draws are made on the device in the reference order
    (brightness, saturation, contrast; shift along H, along W; cutout offset along H, along W); the arithmetic is two
    launches (cb200_diffaug_fwd).  The stages must be listed in the canonical order (the only policy the reference uses is
    'color,cutout'); any other order raises."""

    _ORDER = ("color", "translation", "cutout")

    def __init__(self, policy=""):
        super().__init__()
        self.policy = policy
        stages = [p for p in policy.split(",")] if policy else []
        for p in stages:
            if p not in self._ORDER:
                raise KeyError(p)                                   # AUGMENT_FNS[p] in the reference
        if stages != [p for p in self._ORDER if p in stages]:
            raise NotImplementedError("DiffAugLayer: stages must be a subset of %s in that order (got %r)"
                                      % (",".join(self._ORDER), policy))
        self.stages = stages

    def sample(self, x):
        """[7, B] draws in the reference's torch RNG order (third_party/diffaug.py:24-76); unused rows stay zero."""
        n, _, h, w = x.shape
        dev = x.device
        p = torch.zeros(7, n, device=dev)
        if "color" in self.stages:
            for row in range(3):                                    # rand_brightness, rand_saturation, rand_contrast
                p[row] = torch.rand(n, 1, 1, 1, dtype=x.dtype, device=dev).view(n)
        if "translation" in self.stages:
            sh, sw = int(h * 0.125 + 0.5), int(w * 0.125 + 0.5)
            p[3] = torch.randint(-sh, sh + 1, size=[n, 1, 1], device=dev).view(n).float()
            p[4] = torch.randint(-sw, sw + 1, size=[n, 1, 1], device=dev).view(n).float()
        if "cutout" in self.stages:
            ch, cw = int(h * 0.5 + 0.5), int(w * 0.5 + 0.5)
            p[5] = torch.randint(0, h + (1 - ch % 2), size=[n, 1, 1], device=dev).view(n).float()
            p[6] = torch.randint(0, w + (1 - cw % 2), size=[n, 1, 1], device=dev).view(n).float()
        return p

    def forward(self, inputs):
        if not self.stages:
            return inputs                                           # DiffAugment with an empty policy is the identity
        if inputs.dim() != 4 or inputs.shape[1] != 3:
            raise ValueError("DiffAugLayer expects [B,3,H,W] images, got %s" % (tuple(inputs.shape),))
        flags = sum({"color": 1, "translation": 2, "cutout": 4}[p] for p in self.stages)
        return DiffAugFn.apply(inputs, self.sample(inputs), flags)
