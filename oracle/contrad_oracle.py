"""CPU oracle for the ContraD per-step training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-fp32 / numpy restatement of the reference algorithm
(jh-jeong/ContraD @ 4ac8ce5) for the path named by BASELINE.json `north_star`:

    SimCLR two-view augmentation -> D conv backbone (spectral norm) -> projection MLPs
    -> NT-Xent / supcon-fake -> D/G GAN losses -> backward -> Adam

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product (``contrad_b200``) never does and
fails loudly when its CUDA library is missing.

Pinning: the reference has no tests, golden vectors or known-answer fixtures for this path
(SURVEY.md section 4 / 8c), and all its dense arithmetic lives in an un-vendored, unpinned
PyTorch.  The oracle is therefore pinned against *outputs of the reference itself*, executed
in the build container on torch 2.11 CPU with fixed seeds by ``tests/golden/make_golden.py``
(committed), and stored as small fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them.  Every function cites the reference
file:line it restates (paths relative to the reference root).

Everything is fp32, NCHW, images in [0, 1].  The functions are written on explicit
parameter dictionaries (state_dict keys of the reference) so that the same
oracle can be fed the product's parameters.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# 1. SimCLR augmentation chain
# --------------------------------------------------------------------------------------

PARAM_FIELDS = ("sx", "sy", "bx", "by", "flip", "cj_on", "contrast", "hue", "sat", "val", "gray_on")


def sample_simclr_params(batch, height, width, device="cpu",
                         scale=(0.2, 1.0), ratio=(3. / 4., 4. / 3.),
                         brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1,
                         p_jitter=0.8, p_gray=0.2):
    """Draw the per-sample parameters of the `simclr` chain, consuming numpy's global RNG and
    torch's default generator of `device` in exactly the reference's order:

      RRC      augment/spatial.py:119-143   np uniform x2, np choice, np randint x2
      HFlip    augment/spatial.py:87-89     torch bernoulli(0.5)[B]
      Apply CJ augment/__init__.py:101-102  torch bernoulli(p)[B]
      CJ order augment/color_jitter.py:67   np rand() once per batch
      contrast augment/color_jitter.py:46   torch uniform_[B,1,1,1]
      hsv      augment/color_jitter.py:56-61 torch uniform_ x3 (hue, saturation, brightness)
      Apply Gr augment/__init__.py:101-102  torch bernoulli(p)[B]

    Returns (params, order): params is a dict of float32 [B] tensors on `device` with keys
    PARAM_FIELDS; order == 0 means [contrast, hsv], 1 means [hsv, contrast].
    The reference names shape[2] "width" and shape[3] "height" (spatial.py:113); with the
    square images of every config the two coincide, and the same quirk is kept here.
    """
    B = batch
    ref_width, ref_height = height, width          # spatial.py:113  N, _, width, height = inputs.shape
    area = ref_height * ref_width
    target_area = np.random.uniform(scale[0], scale[1], B * 10) * area
    aspect = np.exp(np.random.uniform(math.log(ratio[0]), math.log(ratio[1]), B * 10))
    w = np.round(np.sqrt(target_area * aspect))
    h = np.round(np.sqrt(target_area / aspect))
    ok = (0 < w) * (w <= ref_width) * (0 < h) * (h <= ref_height)
    w, h = w[ok], h[ok]
    if len(w) > B:
        pick = np.random.choice(len(w), B, replace=False)
        w, h = w[pick], h[pick]
    n_valid = len(w)
    bias_x = np.random.randint(w - ref_width, ref_width - w + 1) / ref_width
    bias_y = np.random.randint(h - ref_height, ref_height - h + 1) / ref_height

    sx = torch.ones(B, dtype=torch.float32)
    sy = torch.ones(B, dtype=torch.float32)
    bx = torch.zeros(B, dtype=torch.float32)
    by = torch.zeros(B, dtype=torch.float32)
    sx[:n_valid] = torch.from_numpy(w / ref_width).float()
    sy[:n_valid] = torch.from_numpy(h / ref_height).float()
    bx[:n_valid] = torch.from_numpy(np.asarray(bias_x, dtype=np.float64)).float()
    by[:n_valid] = torch.from_numpy(np.asarray(bias_y, dtype=np.float64)).float()

    dev = torch.device(device)
    flip = torch.bernoulli(torch.ones(B, device=dev) * 0.5) * 2 - 1
    cj_on = torch.bernoulli(torch.full((B,), p_jitter, device=dev))
    order = 0 if np.random.rand() > 0.5 else 1

    def _range(v, center=1.0, clip0=True):
        lo, hi = center - v, center + v
        if clip0:
            lo = max(lo, 0.0)
        return lo, hi

    def draw_contrast():
        if contrast:
            return torch.empty(B, 1, 1, 1, device=dev).uniform_(*_range(contrast)).view(B)
        return torch.ones(B, device=dev)

    def draw_hsv():
        f_h = torch.zeros(B, 1, 1, device=dev)
        f_s = torch.ones(B, 1, 1, device=dev)
        f_v = torch.ones(B, 1, 1, device=dev)
        if hue:
            f_h.uniform_(*_range(hue, center=0.0, clip0=False))
        if saturation:
            f_s.uniform_(*_range(saturation))
        if brightness:
            f_v.uniform_(*_range(brightness))
        return f_h.view(B), f_s.view(B), f_v.view(B)

    if order == 0:
        f_c = draw_contrast()
        f_h, f_s, f_v = draw_hsv()
    else:
        f_h, f_s, f_v = draw_hsv()
        f_c = draw_contrast()
    gray_on = torch.bernoulli(torch.full((B,), p_gray, device=dev))

    params = {
        "sx": sx.to(dev), "sy": sy.to(dev), "bx": bx.to(dev), "by": by.to(dev),
        "flip": flip, "cj_on": cj_on, "contrast": f_c, "hue": f_h, "sat": f_s, "val": f_v,
        "gray_on": gray_on,
    }
    return params, order


def pack_params(params):
    """[11, B] float32 SoA block in PARAM_FIELDS order (the layout the CUDA kernel takes)."""
    return torch.stack([params[k].float() for k in PARAM_FIELDS], dim=0).contiguous()


def _reflect(coord, size):
    """grid_sample(padding_mode='reflection', align_corners=False) coordinate fold:
    reflect about -0.5 and size-0.5, then clip to [0, size-1] (ATen GridSampler.h
    reflect_coordinates/clip_coordinates; call site augment/spatial.py:146)."""
    span = float(size)
    t = (coord + 0.5).abs()
    extra = torch.fmod(t, span)
    flips = torch.floor(t / span)
    even = torch.remainder(flips, 2.0) == 0
    folded = torch.where(even, extra - 0.5, span - extra - 0.5)
    return folded.clamp(0.0, span - 1.0)


def resized_crop(x, sx, sy, bx, by):
    """RandomResizeCropLayer.forward with explicit theta (augment/spatial.py:138-146):
    affine_grid(align_corners=False) + bilinear grid_sample with reflection padding."""
    B, C, H, W = x.shape
    j = torch.arange(W, dtype=torch.float32, device=x.device)
    i = torch.arange(H, dtype=torch.float32, device=x.device)
    base_x = (2.0 * j + 1.0) / W - 1.0
    base_y = (2.0 * i + 1.0) / H - 1.0
    gx = sx.view(B, 1) * base_x.view(1, W) + bx.view(B, 1)            # [B, W]
    gy = sy.view(B, 1) * base_y.view(1, H) + by.view(B, 1)            # [B, H]
    px = _reflect(((gx + 1.0) * W - 1.0) / 2.0, W)
    py = _reflect(((gy + 1.0) * H - 1.0) / 2.0, H)
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    wx1 = px - x0
    wy1 = py - y0
    wx0 = 1.0 - wx1
    wy0 = 1.0 - wy1
    x0i = x0.long()
    y0i = y0.long()
    x1i = x0i + 1
    y1i = y0i + 1
    # a tap outside the image contributes zero (only x1 == W / y1 == H can occur after the clip)
    wx1 = torch.where(x1i <= W - 1, wx1, torch.zeros_like(wx1))
    wy1 = torch.where(y1i <= H - 1, wy1, torch.zeros_like(wy1))
    x1i = x1i.clamp(max=W - 1)
    y1i = y1i.clamp(max=H - 1)

    def rows(idx):      # gather rows: [B, C, H, W] -> [B, C, H(out), W]
        return torch.gather(x, 2, idx.view(B, 1, H, 1).expand(B, C, H, W))

    def cols(t, idx):   # gather cols
        return torch.gather(t, 3, idx.view(B, 1, 1, W).expand(B, C, H, W))

    top, bot = rows(y0i), rows(y1i)
    wx0e, wx1e = wx0.view(B, 1, 1, W), wx1.view(B, 1, 1, W)
    wy0e, wy1e = wy0.view(B, 1, H, 1), wy1.view(B, 1, H, 1)
    out = (cols(top, x0i) * wx0e * wy0e + cols(top, x1i) * wx1e * wy0e
           + cols(bot, x0i) * wx0e * wy1e + cols(bot, x1i) * wx1e * wy1e)
    return out


def hflip(x, sign):
    """HorizontalFlipLayer.forward (augment/spatial.py:84-93).  The reference resamples with
    theta=[[+-1,0,0],[0,1,0]]; the sample points fall on pixel centres, so the result is the
    exact mirror / identity (max deviation 0 at 32x32, 3e-5 at 512x512: SURVEY row a3)."""
    flipped = torch.flip(x, dims=[3])
    return torch.where(sign.view(-1, 1, 1, 1) < 0, flipped, x)


class _StraightThroughHSV(torch.autograd.Function):
    """RandomHSVFunction (augment/color_jitter.py:81-104): forward rgb->hsv, jitter, hsv->rgb
    (augment/utils.py:27-38,55-63); backward is the identity on the image."""

    @staticmethod
    def forward(ctx, x, f_h, f_s, f_v):
        r, g, b = x[:, 0], x[:, 1], x[:, 2]
        cmax = x.max(1)[0]
        cmin = x.min(1)[0]
        hue = torch.atan2(math.sqrt(3) * (g - b), 2 * r - g - b)
        hue = torch.remainder(hue, 2 * math.pi) / (2 * math.pi)
        sat = 1 - cmin / (cmax + 1e-8)
        val = cmax
        hsv = torch.stack([hue, sat, val], dim=1)
        hsv = torch.where(torch.isfinite(hsv), hsv, torch.zeros_like(hsv))
        h = torch.remainder(hsv[:, 0] + f_h.view(-1, 1, 1) * (255. / 360.), 1.0)
        s = hsv[:, 1] * f_s.view(-1, 1, 1)
        v = hsv[:, 2] * f_v.view(-1, 1, 1)
        h, s, v = h.clamp(0, 1), s.clamp(0, 1), v.clamp(0, 1)
        c = v * s
        outs = []
        for n in (5.0, 3.0, 1.0):
            k = torch.remainder(n + h * 6.0, 6.0)
            t = torch.min(k, 4.0 - k).clamp(0, 1)
            outs.append(v - c * t)
        return torch.stack(outs, dim=1)

    @staticmethod
    def backward(ctx, grad):
        return grad.clone(), None, None, None


def adjust_contrast(x, factor):
    """ColorJitterLayer.adjust_contrast (augment/color_jitter.py:44-49)."""
    mean = x.mean(dim=[2, 3], keepdim=True)
    return ((x - mean) * factor.view(-1, 1, 1, 1) + mean).clamp(0, 1)


def color_jitter(x, f_c, f_h, f_s, f_v, order):
    """ColorJitterLayer.transform (augment/color_jitter.py:65-75)."""
    if order == 0:
        x = adjust_contrast(x, f_c)
        return _StraightThroughHSV.apply(x, f_h, f_s, f_v)
    x = _StraightThroughHSV.apply(x, f_h, f_s, f_v)
    return adjust_contrast(x, f_c)


def to_gray(x):
    """RandomColorGrayLayer.forward (augment/__init__.py:81-91)."""
    l = 0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3]
    return torch.cat([l, l, l], dim=1)


def _blend(x, fx, mask):
    """RandomApply.forward (augment/__init__.py:100-103)."""
    m = mask.view(-1, 1, 1, 1)
    return x * (1 - m) + fx * m


def augment_simclr(x, params, order):
    """The `simclr` chain (augment/__init__.py:106-112) on explicit parameters."""
    x = resized_crop(x, params["sx"], params["sy"], params["bx"], params["by"])
    x = hflip(x, params["flip"])
    x = _blend(x, color_jitter(x, params["contrast"], params["hue"], params["sat"], params["val"], order),
               params["cj_on"])
    x = _blend(x, to_gray(x), params["gray_on"])
    return x


def sample_hq_params(batch, height, width, device="cpu", sigma_range=(0.1, 2.0), p_blur=0.5, p_cut=0.5, cutout=False):
    """The draws `simclr_hq` / `simclr_hq_cutout` make AFTER those of `simclr` (call sample_simclr_params first):
      Apply blur  augment/__init__.py:101-102  torch bernoulli(0.5)[B]
      sigma       augment/__init__.py:73       np uniform(*sigma_range), once per batch
      Apply cut   augment/__init__.py:101-102  torch bernoulli(0.5)[B]                     (simclr_hq_cutout only)
      centres     augment/spatial.py:169-170   torch randint(h, (B,1)), randint(w, (B,1))
    Returns a dict with blur_on [B], sigma (float) and, with cutout, cut_on / h_center / w_center [B]."""
    dev = torch.device(device)
    out = {"blur_on": torch.bernoulli(torch.full((batch,), p_blur, device=dev)),
           "sigma": float(np.random.uniform(*sigma_range))}
    if cutout:
        out["cut_on"] = torch.bernoulli(torch.full((batch,), p_cut, device=dev))
        out["h_center"] = torch.randint(height, (batch, 1), device=dev).view(batch)
        out["w_center"] = torch.randint(width, (batch, 1), device=dev).view(batch)
    return out


def gaussian_blur(x, sigma):
    """GaussianBlur.forward (augment/__init__.py:64-78) with kornia's two entry points restated (kornia is unpinned
    in the reference and absent here - parity unpinned): k = 2*int((H//10)/2)+1, the DENSE k x k kernel
    outer(g, g) of the normalised 1-D Gaussian g, 'reflect' padding, depthwise correlation."""
    b, c, h, w = x.shape
    k = int((h // 10) / 2) * 2 + 1
    t = torch.arange(k, dtype=torch.float32, device=x.device) - k // 2
    g = torch.exp(-t.pow(2.0) / (2.0 * float(sigma) ** 2))
    g = g / g.sum()
    kern = torch.outer(g, g)
    r = k // 2
    xp = F.pad(x, (r, r, r, r), mode="reflect") if r > 0 else x
    return F.conv2d(xp, kern.expand(c, 1, k, k).contiguous(), groups=c)


def cutout(x, h_center, w_center, length):
    """CutOut.forward (augment/spatial.py:163-181): 1 - outer(box_h, box_w) with boxes of `length` around the centres."""
    b, _, h, w = x.shape
    half = (length - 1) // 2
    rows = torch.arange(h, device=x.device)[None, :]
    cols = torch.arange(w, device=x.device)[None, :]
    mh = ((rows - h_center.view(b, 1).long()).abs() <= half).float()
    mw = ((cols - w_center.view(b, 1).long()).abs() <= half).float()
    mask = 1.0 - mh[:, None, :, None] * mw[:, None, None, :]
    return x * mask


def augment_simclr_hq(x, params, order, hq, cutout_length=None):
    """`simclr_hq` (augment/__init__.py:115-122) / `simclr_hq_cutout` (:125-133) on explicit parameters."""
    x = augment_simclr(x, params, order)
    x = _blend(x, gaussian_blur(x, hq["sigma"]), hq["blur_on"])
    if "cut_on" in hq:
        x = _blend(x, cutout(x, hq["h_center"], hq["w_center"], cutout_length), hq["cut_on"])
    return x


# --------------------------------------------------------------------------------------
# 2. Spectral norm + SNDCGAN discriminator / generator on explicit parameter dicts
# --------------------------------------------------------------------------------------

# ---- rows f3 / f4 of SURVEY 8f: uint8 input, hfrt / RandomCrop, Gaussian noise, baseline modes --------------------

def to_tensor_u8(x_u8):
    """The dataset transform `ToTensor` on NCHW bytes (datasets.py:10-21 -> torchvision to_tensor:
    `img.to(float32).div(255)`)."""
    return x_u8.to(torch.float32).div(255)


def sample_shift_flip(batch, max_pixels=4, width=32, flip=True, device="cpu"):
    """Random draws of HorizontalFlipRandomCrop (flip=True, augment/spatial.py:31-33) / RandomCrop (flip=False,
    spatial.py:60-61) in the reference order.  Returns [3, B] = {sign, bias_x, bias_y}."""
    params = torch.empty(3, batch, device=device)
    if flip:
        params[0] = torch.bernoulli(torch.ones(batch, device=device) * 0.5) * 2 - 1
    else:
        params[0] = 1.0
    r_bias = torch.randint(-max_pixels, max_pixels + 1, (batch, 2), device=device).float() / (width / 2)
    params[1:3] = r_bias.t()
    return params


def _nearest_index(g, size, padding_mode):
    """grid_sample(mode='nearest', align_corners=False) index pipeline of one axis (ATen GridSampler.h:
    grid_sampler_unnormalize -> clip / reflect -> nearbyint -> bounds check).  Returns (index, valid)."""
    c = ((g + 1.0) * size - 1.0) / 2.0
    if padding_mode == "border":
        c = c.clamp(0.0, float(size) - 1.0)
    elif padding_mode == "reflection":
        c = _reflect(c, size)
    elif padding_mode != "zeros":
        raise ValueError(padding_mode)
    r = torch.round(c)                      # round-half-even == nearbyint
    valid = (r >= 0) & (r <= size - 1)
    return r.clamp(0, size - 1).long(), valid


def shift_flip(x, params, padding_mode="reflection"):
    """HorizontalFlipRandomCrop.forward / RandomCrop.forward on explicit draws (augment/spatial.py:25-40,54-67):
    theta = [[sign, 0, bias_x], [0, 1, bias_y]], affine_grid + nearest grid_sample, restated as an index gather."""
    B, C, H, W = x.shape
    sign, bx, by = params[0], params[1], params[2]
    j = torch.arange(W, dtype=torch.float32, device=x.device)
    i = torch.arange(H, dtype=torch.float32, device=x.device)
    gx = sign.view(B, 1) * ((2.0 * j + 1.0) / W - 1.0).view(1, W) + bx.view(B, 1)
    gy = ((2.0 * i + 1.0) / H - 1.0).view(1, H) + by.view(B, 1)
    xi, xv = _nearest_index(gx, W, padding_mode)        # [B, W]
    yi, yv = _nearest_index(gy, H, padding_mode)        # [B, H]
    rows = torch.gather(x, 2, yi.view(B, 1, H, 1).expand(B, C, H, W))
    out = torch.gather(rows, 3, xi.view(B, 1, 1, W).expand(B, C, H, W))
    mask = (yv.view(B, 1, H, 1) & xv.view(B, 1, 1, W)).to(x.dtype)
    return out * mask


def gaussian_noise(x, noise, sigma):
    """Gaussian.forward on an explicit noise draw (augment/__init__.py:46-49)."""
    return (x + noise * sigma).clamp(0, 1)


def sample_diffaug(batch, height, width, stages=("color", "cutout"), device="cpu"):
    """Draws of DiffAugment in the reference's torch RNG order (third_party/diffaug.py:24-76).  [7, B]:
    r_brightness, r_saturation, r_contrast, shift along H, along W, cutout offset along H, along W."""
    p = torch.zeros(7, batch, device=device)
    if "color" in stages:
        for row in range(3):
            p[row] = torch.rand(batch, 1, 1, 1, device=device).view(batch)
    if "translation" in stages:
        sh, sw = int(height * 0.125 + 0.5), int(width * 0.125 + 0.5)
        p[3] = torch.randint(-sh, sh + 1, size=[batch, 1, 1], device=device).view(batch).float()
        p[4] = torch.randint(-sw, sw + 1, size=[batch, 1, 1], device=device).view(batch).float()
    if "cutout" in stages:
        ch, cw = int(height * 0.5 + 0.5), int(width * 0.5 + 0.5)
        p[5] = torch.randint(0, height + (1 - ch % 2), size=[batch, 1, 1], device=device).view(batch).float()
        p[6] = torch.randint(0, width + (1 - cw % 2), size=[batch, 1, 1], device=device).view(batch).float()
    return p


def diffaug(x, p, stages=("color", "cutout")):
    """DiffAugment(x, policy) on explicit draws (third_party/diffaug.py:8-76), stages in the order color, translation,
    cutout.  Plain slicing / masking instead of the reference's meshgrid fancy indexing."""
    B, C, H, W = x.shape
    x = 2.0 * x - 1.0
    if "color" in stages:
        x = x + (p[0].view(B, 1, 1, 1) - 0.5)                                           # rand_brightness :24-26
        m = x.mean(dim=1, keepdim=True)
        x = (x - m) * (p[1].view(B, 1, 1, 1) * 2) + m                                   # rand_saturation :29-32
        m = x.mean(dim=[1, 2, 3], keepdim=True)
        x = (x - m) * (p[2].view(B, 1, 1, 1) + 0.5) + m                                 # rand_contrast :35-38
    if "translation" in stages:                                                         # rand_translation :41-54
        i = torch.arange(H, device=x.device).view(1, H) + p[3].long().view(B, 1)        # source row of output row
        j = torch.arange(W, device=x.device).view(1, W) + p[4].long().view(B, 1)
        vi, vj = (i >= 0) & (i < H), (j >= 0) & (j < W)
        rows = torch.gather(x, 2, i.clamp(0, H - 1).view(B, 1, H, 1).expand(B, C, H, W))
        out = torch.gather(rows, 3, j.clamp(0, W - 1).view(B, 1, 1, W).expand(B, C, H, W))
        x = out * (vi.view(B, 1, H, 1) & vj.view(B, 1, 1, W)).to(x.dtype)
    if "cutout" in stages:                                                              # rand_cutout :57-72
        sh, sw = int(H * 0.5 + 0.5), int(W * 0.5 + 0.5)
        lo_h = p[5].long().view(B, 1) - sh // 2
        lo_w = p[6].long().view(B, 1) - sw // 2
        ii = torch.arange(H, device=x.device).view(1, H)
        jj = torch.arange(W, device=x.device).view(1, W)
        cut_h = (ii >= lo_h.clamp(min=0)) & (ii <= (lo_h + sh - 1).clamp(max=H - 1))
        cut_w = (jj >= lo_w.clamp(min=0)) & (jj <= (lo_w + sw - 1).clamp(max=W - 1))
        mask = 1.0 - (cut_h.view(B, 1, H, 1) & cut_w.view(B, 1, 1, W)).to(x.dtype)
        x = x * mask
    return 0.5 * x + 0.5


def penalty_cr(sd_d, d_real, images, aug, lbd, training=True):
    """penalty.consistency (penalty.py:47-49) with `aug` = the augmentation on explicit draws."""
    d_aug, _ = d_sndcgan_forward(sd_d, aug(images), training=training)
    return lbd * ((d_real - d_aug) ** 2).mean()


def penalty_bcr(sd_d, d_real, d_gen, all_images, aug, lbd, lbd2, training=True):
    """penalty.balanced_consistency (penalty.py:52-60)."""
    d_aug_all, _ = d_sndcgan_forward(sd_d, aug(all_images), training=training)
    n = all_images.shape[0] // 2
    return lbd * ((d_real - d_aug_all[:n]) ** 2).mean() + lbd2 * ((d_gen - d_aug_all[n:]) ** 2).mean()


def loss_d_baseline(sd_d, mode, images, gen_images, augs, loss="nonsat", penalty="none", lbd=10.0, lbd2=10.0,
                    training=True):
    """training/gan/{std,aug,aug_both}.py loss_D_fn.  `augs` is the list of augmentation callables (explicit draws)
    in call order: the mode's own `P.augment_fn` call first (aug / aug_both), then the penalty's.
    Returns (d_loss, penalty, d_real.mean(), d_gen.mean()).  NOTE: in train mode every D forward advances the
    spectral-norm power iteration in `sd_d` (as the reference's hooks do)."""
    augs = list(augs)
    gen_images = gen_images.detach()
    n = images.shape[0]
    if mode == "std":                                   # std.py:11-13
        all_images = torch.cat([images, gen_images], dim=0)
        d_inputs = all_images
    elif mode == "aug":                                 # aug.py:11-13
        all_images = torch.cat([augs.pop(0)(images), gen_images], dim=0)
        d_inputs = all_images
    elif mode == "aug_both":                            # aug_both.py:12-14
        all_images = torch.cat([images, gen_images], dim=0)
        d_inputs = augs.pop(0)(all_images)
    else:
        raise NotImplementedError(mode)
    d_all, _ = d_sndcgan_forward(sd_d, d_inputs, training=training)
    d_real, d_gen = d_all[:n], d_all[n:]
    d_loss = gan_d_loss(d_real, d_gen, loss)
    if penalty == "none":
        pen = torch.zeros(1)
    elif penalty == "cr":
        pen = penalty_cr(sd_d, d_real, images, augs.pop(0), lbd, training)
    elif penalty == "bcr":
        pen = penalty_bcr(sd_d, d_real, d_gen, all_images, augs.pop(0), lbd, lbd2, training)
    else:
        raise NotImplementedError(penalty)
    return d_loss, pen, d_real.mean(), d_gen.mean()


def _l2normalize(v, eps=1e-12):
    return v / v.norm().clamp_min(eps)


def spectral_normalize(weight, u, v, training=True, eps=1e-12):
    """torch.nn.utils.spectral_norm (torch/nn/utils/spectral_norm.py:92-114), as applied by
    models/gan/sndcgan.py:111-118: one power iteration in train mode (u, v updated IN PLACE,
    no grad), sigma = u^T W v with grad flowing to W only, returns W / sigma."""
    w_mat = weight.reshape(weight.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(_l2normalize(torch.mv(w_mat.t(), u), eps))
            u.copy_(_l2normalize(torch.mv(w_mat, v), eps))
    u_c, v_c = u.clone(), v.clone()
    sigma = torch.dot(u_c, torch.mv(w_mat, v_c))
    return weight / sigma


D_CONVS = (("main.0", 1, 1), ("main.2", 2, 1), ("main.4", 1, 1), ("main.6", 2, 1),
           ("main.8", 1, 1), ("main.10", 2, 1), ("main.12", 1, 1))      # (key, stride, pad)
D_LINEARS = ("linear.l1", "linear.l2", "projection.0", "projection.2", "projection2.0", "projection2.2")


def _sn_weight(sd, key, training):
    return spectral_normalize(sd[key + ".weight_orig"], sd[key + ".weight_u"], sd[key + ".weight_v"],
                              training=training)


def d_sndcgan_penultimate(sd, x, training=True):
    """D_SNDCGAN.penultimate (models/gan/sndcgan.py:122-128; layers :91-109)."""
    h = x * 2. - 1.
    for key, stride, pad in D_CONVS:
        h = F.conv2d(h, _sn_weight(sd, key, training), sd[key + ".bias"], stride=stride, padding=pad)
        h = F.leaky_relu(h, 0.1)
    return h.reshape(h.shape[0], -1)


def d_heads(sd, features, sg_linear=False, training=True):
    """BaseDiscriminator.forward heads (models/gan/base.py:123-133) with TinyDiscriminator
    (base.py:14-35, mlp_linear=True) and the two projection MLPs (base.py:92-101)."""
    feats_d = features.detach() if sg_linear else features

    def mlp(x, k1, k2):
        hid = F.leaky_relu(F.linear(x, _sn_weight(sd, k1, training), sd[k1 + ".bias"]), 0.1)
        return F.linear(hid, _sn_weight(sd, k2, training), sd[k2 + ".bias"])

    out = mlp(feats_d, "linear.l1", "linear.l2")
    proj = mlp(features, "projection.0", "projection.2")
    proj2 = mlp(features, "projection2.0", "projection2.2")
    out = out + (proj.mean() + proj2.mean()) * 0.
    return out, proj, proj2


def d_sndcgan_forward(sd, x, sg_linear=False, training=True):
    feats = d_sndcgan_penultimate(sd, x, training)
    out, proj, proj2 = d_heads(sd, feats, sg_linear=sg_linear, training=training)
    return out, {"penultimate": feats, "projection": proj, "projection2": proj2}


# D_SNResNet18 (models/gan/snresnet.py:21-93): conv1 3->64, then 4 stages of 2 BasicBlocks (64, 128, 256, 512 planes;
# the first block of stages 2-4 has stride 2 and a 1x1 stride-2 shortcut conv); every conv / linear spectrally normalised.
RESNET18_BLOCKS = tuple(("layer%d.%d" % (s + 1, b), cin, planes, stride)
                        for s, (planes, first_stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)))
                        for b, (cin, stride) in enumerate((((64, 64, 128, 256)[s], first_stride), (planes, 1))))


def d_snresnet18_penultimate(sd, x, training=True):
    """SNResNet.penultimate (snresnet.py:76-89) with BasicBlock.forward (:39-44)."""
    def conv(key, h, stride, pad):
        return F.conv2d(h, _sn_weight(sd, key, training), sd[key + ".bias"], stride=stride, padding=pad)

    h = F.leaky_relu(conv("conv1", x * 2. - 1., 1, 1), 0.1)
    for name, cin, planes, stride in RESNET18_BLOCKS:
        out = F.leaky_relu(conv(name + ".conv1", h, stride, 1), 0.1)
        out = conv(name + ".conv2", out, 1, 1)
        sc = conv(name + ".shortcut.0", h, stride, 0) if (stride != 1 or cin != planes) else h
        h = F.leaky_relu(out + sc, 0.1)
    h = F.avg_pool2d(h, 4)
    return h.reshape(h.shape[0], -1)


def d_snresnet18_forward(sd, x, sg_linear=False, training=True):
    feats = d_snresnet18_penultimate(sd, x, training)
    out, proj, proj2 = d_heads(sd, feats, sg_linear=sg_linear, training=training)
    return out, {"penultimate": feats, "projection": proj, "projection2": proj2}


def make_d_resnet18_state(d_hidden=1024, d_project=128, generator=None):
    """A D_SNResNet18(mlp_linear=True, d_hidden=1024) state_dict (models/gan/__init__.py:8-12): nn.Conv2d / nn.Linear
    default initialisation (uniform +-1/sqrt(fan_in) for weights and biases), fresh unit-norm u / v."""
    shapes = {"conv1": (64, 3, 3, 3)}
    for name, cin, planes, stride in RESNET18_BLOCKS:
        shapes[name + ".conv1"] = (planes, cin, 3, 3)
        shapes[name + ".conv2"] = (planes, planes, 3, 3)
        if stride != 1 or cin != planes:
            shapes[name + ".shortcut.0"] = (planes, cin, 1, 1)
    shapes.update({"linear.l1": (d_hidden, 512), "linear.l2": (1, d_hidden), "projection.0": (d_hidden, 512),
                   "projection.2": (d_project, d_hidden), "projection2.0": (d_hidden, 512),
                   "projection2.2": (d_project, d_hidden)})
    sd = {}
    for key, shp in shapes.items():
        fan = 1
        for d in shp[1:]:
            fan *= d
        bound = 1.0 / math.sqrt(fan)
        sd[key + ".weight_orig"] = (torch.rand(*shp, generator=generator) * 2 - 1) * bound
        sd[key + ".bias"] = (torch.rand(shp[0], generator=generator) * 2 - 1) * bound
        sd[key + ".weight_u"] = _l2normalize(torch.empty(shp[0]).normal_(0, 1, generator=generator))
        sd[key + ".weight_v"] = _l2normalize(torch.empty(fan).normal_(0, 1, generator=generator))
    return sd


def _batch_norm_train(x, sd, key, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d in train mode (batch statistics; running stats updated with the unbiased
    variance), as used by G_SNDCGAN (models/gan/sndcgan.py:25-36)."""
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                        sd[key + ".weight"], sd[key + ".bias"], True, momentum, eps)


def g_sndcgan_forward(sd, z, ngf=64, s_hb=4, s_wb=4):
    """G_SNDCGAN.forward (models/gan/sndcgan.py:41-48)."""
    h = F.linear(z, sd["linear.weight"], sd["linear.bias"])
    h = h.view(h.shape[0], h.shape[1], 1, 1)
    h = F.relu(_batch_norm_train(h, sd, "norm_init"))
    h = h.view(-1, ngf * 8, s_hb, s_wb)
    for conv, bn in (("main.0", "main.1"), ("main.3", "main.4"), ("main.6", "main.7")):
        h = F.conv_transpose2d(h, sd[conv + ".weight"], sd[conv + ".bias"], stride=2, padding=1)
        h = F.relu(_batch_norm_train(h, sd, bn))
    h = F.conv_transpose2d(h, sd["main.9.weight"], sd["main.9.bias"], stride=1, padding=1)
    return 0.5 * torch.tanh(h) + 0.5


def sample_latent(n, nz=128):
    """G_SNDCGAN.sample_latent (models/gan/sndcgan.py:50-52): U(-1,1) drawn on the CPU generator."""
    return torch.empty(n, nz).uniform_(-1, 1)


# --------------------------------------------------------------------------------------
# 3. Contrastive + GAN losses
# --------------------------------------------------------------------------------------

def nt_xent(out1, out2, temperature=0.1):
    """training/criterion.py:24-45 (single process; the distributed branch only concatenates
    all-gathered rows in rank order, see gather semantics below)."""
    n = out1.shape[0]
    z = torch.cat([out1, out2], dim=0)
    sim = (z @ z.t()) / temperature
    sim = sim.masked_fill(torch.eye(2 * n, dtype=torch.bool, device=z.device), -5e4)
    lsm = F.log_softmax(sim, dim=1)
    idx = torch.arange(n, device=z.device)
    return -(lsm[idx, idx + n] + lsm[idx + n, idx]).sum() / (2 * n)


def supcon_fake(out1, out2, others, temperature=0.1):
    """training/gan/contrad.py:8-32: rows 2N..3N (the fakes) against all 3N columns; positives
    are the other fakes with weight 1/(N-1)."""
    n = out1.shape[0]
    z = torch.cat([out1, out2, others], dim=0)
    sim = (z @ z.t()) / temperature
    sim = sim.masked_fill(torch.eye(3 * n, dtype=torch.bool, device=z.device), -5e4)
    rows = sim[2 * n:]
    mask = torch.zeros_like(rows)
    mask[:, 2 * n:] = 1
    mask[torch.arange(n), torch.arange(n) + 2 * n] = 0
    mask = mask / mask.sum(1, keepdim=True)
    lsm = F.log_softmax(rows, dim=1)
    return -(lsm * mask).sum(1).mean()


def gather_rank_major(chunks):
    """third_party/gather_layer.py:8-23 + the torch.cat at training/criterion.py:31-32: the
    gathered tensor is the rank-major concatenation; backward keeps grads[rank]."""
    return torch.cat(list(chunks), dim=0)


def gan_d_loss(d_real, d_gen, kind):
    """training/gan/contrad.py:52-64."""
    if kind == "nonsat":
        return F.softplus(d_gen).mean() + F.softplus(-d_real).mean()
    if kind == "wgan":
        return d_gen.mean() - d_real.mean()
    if kind == "hinge":
        return F.relu(1. + d_gen).mean() + F.relu(1. - d_real).mean()
    if kind == "lsgan":
        return 0.5 * (((d_real - 1.0) ** 2).mean() + (d_gen ** 2).mean())
    raise NotImplementedError(kind)


def gan_g_loss(d_gen, kind):
    """training/gan/contrad.py:75-80."""
    if kind == "nonsat":
        return F.softplus(-d_gen).mean()
    if kind == "lsgan":
        return 0.5 * ((d_gen - 1.0) ** 2).mean()
    return -d_gen.mean()


def loss_d(sd_d, images, gen_images, aug_params, aug_order, temp=0.1, lbd_a=1.0, loss="nonsat",
           training=True):
    """training/gan/contrad.py:35-70 on explicit augmentation parameters.
    Returns (L_con+ + lbd_a * L_con-, L_dis, extras)."""
    n = images.shape[0]
    cat = torch.cat([images, images, gen_images.detach()], dim=0)
    d_all, aux = d_sndcgan_forward(sd_d, augment_simclr(cat, aug_params, aug_order),
                                   sg_linear=True, training=training)
    views = F.normalize(aux["projection"])
    l_pos = nt_xent(views[:n], views[n:2 * n], temp)
    reals = F.normalize(aux["projection2"])
    l_neg = supcon_fake(reals[:n], reals[n:2 * n], reals[2 * n:], temp)
    d_real, d_gen = d_all[:n], d_all[2 * n:3 * n]
    l_dis = gan_d_loss(d_real, d_gen, loss)
    extras = {"l_con_pos": l_pos, "l_con_neg": l_neg, "d_real": d_real.mean(), "d_gen": d_gen.mean(),
              "d_all": d_all, "projection": aux["projection"], "projection2": aux["projection2"]}
    return l_pos + lbd_a * l_neg, l_dis, extras


def loss_g(sd_d, gen_images, aug_params, aug_order, loss="nonsat", training=True):
    """training/gan/contrad.py:73-82."""
    d_gen, _ = d_sndcgan_forward(sd_d, augment_simclr(gen_images, aug_params, aug_order), training=training)
    return gan_g_loss(d_gen, loss)


# --------------------------------------------------------------------------------------
# 4. Parameters, Adam and the full train step (train_gan.py:141-179)
# --------------------------------------------------------------------------------------

D_SHAPES = {
    "main.0": (64, 3, 3, 3), "main.2": (128, 64, 4, 4), "main.4": (128, 128, 3, 3),
    "main.6": (256, 128, 4, 4), "main.8": (256, 256, 3, 3), "main.10": (512, 256, 4, 4),
    "main.12": (512, 512, 3, 3),
}


def make_d_state(ndf=64, d_hidden=512, d_project=128, s=4, generator=None):
    """A D_SNDCGAN state_dict (keys/shapes of SURVEY A.6) filled from `generator` in a fixed,
    name-sorted order: weights N(0, 0.02), biases 0, u/v unit-normalised N(0,1)
    (models/gan/sndcgan.py:130-148 init law; the draw ORDER is this repo's own so the same
    tensors can be rebuilt without the reference)."""
    shapes = {}
    chans = [3, ndf, ndf * 2, ndf * 2, ndf * 4, ndf * 4, ndf * 8, ndf * 8]
    ks = [3, 4, 3, 4, 3, 4, 3]
    for i, (key, _, _) in enumerate(D_CONVS):
        shapes[key] = (chans[i + 1], chans[i], ks[i], ks[i])
    nfeat = ndf * 8 * s * s
    shapes.update({"linear.l1": (d_hidden, nfeat), "linear.l2": (1, d_hidden),
                   "projection.0": (d_hidden, nfeat), "projection.2": (d_project, d_hidden),
                   "projection2.0": (d_hidden, nfeat), "projection2.2": (d_project, d_hidden)})
    sd = {}
    for key in sorted(shapes):
        shp = shapes[key]
        fan = int(np.prod(shp[1:]))
        sd[key + ".weight_orig"] = torch.empty(shp).normal_(0.0, 0.02, generator=generator)
        sd[key + ".bias"] = torch.zeros(shp[0])
        sd[key + ".weight_u"] = _l2normalize(torch.empty(shp[0]).normal_(0, 1, generator=generator))
        sd[key + ".weight_v"] = _l2normalize(torch.empty(fan).normal_(0, 1, generator=generator))
    return sd


def make_g_state(ngf=64, nz=128, s=4, generator=None):
    """A G_SNDCGAN state_dict: Linear/ConvT N(0,0.02), BN weight 1 bias 0
    (models/gan/sndcgan.py:54-66)."""
    sd = {}
    c0 = ngf * 8 * s * s
    sd["linear.weight"] = torch.empty(c0, nz).normal_(0.0, 0.02, generator=generator)
    sd["linear.bias"] = torch.zeros(c0)

    def bn(key, c):
        sd[key + ".weight"] = torch.ones(c)
        sd[key + ".bias"] = torch.zeros(c)
        sd[key + ".running_mean"] = torch.zeros(c)
        sd[key + ".running_var"] = torch.ones(c)
        sd[key + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    bn("norm_init", c0)
    chans = [ngf * 8, ngf * 4, ngf * 2, ngf]
    for i, conv in enumerate(("main.0", "main.3", "main.6")):
        sd[conv + ".weight"] = torch.empty(chans[i], chans[i + 1], 4, 4).normal_(0.0, 0.02, generator=generator)
        sd[conv + ".bias"] = torch.zeros(chans[i + 1])
        bn("main.%d" % (3 * i + 1), chans[i + 1])
    sd["main.9.weight"] = torch.empty(ngf, 3, 3, 3).normal_(0.0, 0.02, generator=generator)
    sd["main.9.bias"] = torch.zeros(3)
    return sd


def trainable(sd):
    return {k: v for k, v in sd.items()
            if v.is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))}


def set_requires_grad(sd, flag):
    """utils.set_grad (utils.py:125-127)."""
    for v in trainable(sd).values():
        v.requires_grad_(flag)


class Adam(object):
    """torch.optim.Adam as configured by train_gan.py:273-274 (no weight decay, no amsgrad)."""

    def __init__(self, params, lr, betas=(0.5, 0.999), eps=1e-8):
        self.params = list(params)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        c1 = 1 - b1 ** self.t
        c2 = 1 - b2 ** self.t
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(c2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / c1)


def warmup_lr(step, warmup, lr):
    """train_gan.py:88-93."""
    return min(1., (step + 1) / warmup) * lr if warmup > 0 else lr


def grad_norm(sd):
    sq = [p.grad.double().pow(2).sum() for p in trainable(sd).values() if p.grad is not None]
    return float(torch.stack(sq).sum().sqrt()) if sq else 0.0


def train_step(sd_g, sd_d, opt_g, opt_d, images, z_d, z_g, aug_d, aug_g, step=1,
               temp=0.1, lbd_a=1.0, loss="nonsat", lr=2e-4, warmup=3000):
    """One full step of train_gan.py:141-179 (n_critic=1) on explicit latents and augmentation
    parameters: D-step (G forward no-grad, loss_D, backward, Adam) then G-step (G forward,
    loss_G through the frozen D, backward, Adam).  aug_d / aug_g = (params, order)."""
    opt_g.lr = warmup_lr(step, warmup, lr)
    opt_d.lr = warmup_lr(step, warmup, lr)
    out = {}
    # ---- D step
    set_requires_grad(sd_g, False)
    set_requires_grad(sd_d, True)
    with torch.no_grad():
        gen = g_sndcgan_forward(sd_g, z_d)
    l_con, l_dis, ex = loss_d(sd_d, images, gen, aug_d[0], aug_d[1], temp, lbd_a, loss)
    opt_d.zero_grad()
    (l_con + l_dis).backward()
    out.update(l_con_pos=float(ex["l_con_pos"].detach()), l_con_neg=float(ex["l_con_neg"].detach()),
               l_dis=float(l_dis.detach()), d_real=float(ex["d_real"].detach()), d_gen=float(ex["d_gen"].detach()),
               d_grad_norm=grad_norm(sd_d))
    opt_d.step()
    # ---- G step
    set_requires_grad(sd_g, True)
    set_requires_grad(sd_d, False)
    gen = g_sndcgan_forward(sd_g, z_g)
    l_gen = loss_g(sd_d, gen, aug_g[0], aug_g[1], loss)
    opt_g.zero_grad()
    l_gen.backward()
    out.update(l_gen=float(l_gen.detach()), g_grad_norm=grad_norm(sd_g))
    opt_g.step()
    return out
