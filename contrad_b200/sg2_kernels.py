"""Raw (non-autograd) bindings of the StyleGAN2-side C ABI (include/contrad_b200.h, csrc/sg2_ops.cu): shape
checks, output allocation, stream plumbing.  ``contrad_b200.sg2_functional`` builds the autograd Functions on
these and on the tensor-core bindings of ``contrad_b200.kernels``.  No CPU path: every call needs CUDA tensors."""
import ctypes

import torch

from . import precision
from ._capi import f32, i32, i64, lib, ptr, stream_ptr
from .kernels import _call, _f32c


def _ro(round_out):
    """TF32 rounding of a producer's output: off in the full strict precision mode (contrad_b200/precision.py), where
    activations stay fp32 and the GEMM Functions split them into hi + lo operands at consumption."""
    return 1 if (round_out and not precision.strict_full()) else 0


def _strides4(t, nhwc):
    """{n, c, h, w} element strides of a contiguous 4-D tensor stored as NHWC or NCHW."""
    s = t.stride()
    order = (s[0], s[3], s[1], s[2]) if nhwc else (s[0], s[1], s[2], s[3])
    return (ctypes.c_longlong * 4)(*order)


def upfirdn_out_size(in_size, ksize, up, down, pad0, pad1):
    """op/upfirdn2d.py:131-132."""
    return (in_size * up + pad0 + pad1 - ksize) // down + 1


def upfirdn2d(x, fir, up, down, pad, out_hw=None, nhwc=True, flip=False, gain=1.0, round_out=False):
    """x: [N,H,W,C] (nhwc) or [N,C,H,W]; fir: [kh,kw] device tensor; pad = (x0, x1, y0, y1) like the reference.
    out_hw overrides the output size (the backward pass asks for exactly the forward input size)."""
    x = _f32c(x, "x")
    fir = _f32c(fir, "fir")
    if nhwc:
        N, Hi, Wi, C = x.shape
    else:
        N, C, Hi, Wi = x.shape
    kh, kw = fir.shape
    px0, px1, py0, py1 = pad
    if out_hw is None:
        out_hw = (upfirdn_out_size(Hi, kh, up, down, py0, py1), upfirdn_out_size(Wi, kw, up, down, px0, px1))
    Ho, Wo = out_hw
    y = torch.empty((N, Ho, Wo, C) if nhwc else (N, C, Ho, Wo), device=x.device, dtype=torch.float32)
    _call("upfirdn2d", 0, 4 * (x.numel() + y.numel()), lib().cb200_upfirdn2d, ptr(x), _strides4(x, nhwc), ptr(y),
          _strides4(y, nhwc), ptr(fir), i32(N), i32(C), i32(Hi), i32(Wi), i32(Ho), i32(Wo), i32(up), i32(down), i32(px0),
          i32(py0), i32(kh), i32(kw), i32(1 if flip else 0), f32(gain), i32(1 if nhwc else 0), i32(_ro(round_out)),
          stream_ptr())
    return y


def patch_s2_gather(x, round_out=False):
    """x [B, 2Ho+1, 2Wo+1, C] -> u [B, Ho, Wo, 9, C] (3x3 stride-2 patches, tap-major / channel-minor)."""
    x = _f32c(x, "x")
    B, Hi, Wi, C = x.shape
    assert Hi % 2 == 1 and Wi % 2 == 1, x.shape
    Ho, Wo = (Hi - 1) // 2, (Wi - 1) // 2
    u = torch.empty(B, Ho, Wo, 9, C, device=x.device, dtype=torch.float32)
    _call("patch_s2_gather", 0, 4 * (x.numel() + u.numel()), lib().cb200_patch_s2_gather, ptr(x), ptr(u), i32(B), i32(Ho), i32(Wo),
          i32(C), i32(_ro(round_out)), stream_ptr())
    return u


def patch_s2_scatter(u, round_out=False):
    """u [B, Ho, Wo, 9, C] -> x [B, 2Ho+1, 2Wo+1, C] (transpose of patch_s2_gather)."""
    u = _f32c(u, "u")
    B, Ho, Wo, nine, C = u.shape
    assert nine == 9
    x = torch.empty(B, 2 * Ho + 1, 2 * Wo + 1, C, device=u.device, dtype=torch.float32)
    _call("patch_s2_scatter", 0, 4 * (x.numel() + u.numel()), lib().cb200_patch_s2_scatter, ptr(u), ptr(x), i32(B), i32(Ho),
          i32(Wo), i32(C), i32(_ro(round_out)), stream_ptr())
    return x


def bias_act(x, bias, slope, gain, res=None, round_out=False):
    """y = lrelu_slope(x + bias[c]) * gain (+ res); the channel is the last dimension."""
    x = _f32c(x, "x")
    C = x.shape[-1]
    if res is not None:
        res = _f32c(res, "res")
        assert res.shape == x.shape
    y = torch.empty_like(x)
    _call("bias_act", 0, 8 * x.numel(), lib().cb200_bias_act, ptr(x), ptr(bias), ptr(None), ptr(res), ptr(y), i64(x.numel()),
          i32(C), i32(0), f32(slope), f32(gain), i32(_ro(round_out)), stream_ptr())
    return y


def bias_act_grad(g, ref, bias, slope, gain, round_out=False):
    """g * ((ref + bias[c]) > 0 ? gain : gain * slope)."""
    g = _f32c(g, "g")
    ref = _f32c(ref, "ref")
    assert g.shape == ref.shape, (g.shape, ref.shape)
    y = torch.empty_like(g)
    _call("bias_act", 0, 12 * g.numel(), lib().cb200_bias_act, ptr(g), ptr(bias), ptr(ref), ptr(None), ptr(y), i64(g.numel()),
          i32(g.shape[-1]), i32(1), f32(slope), f32(gain), i32(_ro(round_out)), stream_ptr())
    return y


def modulate(x, s, batch=None, alpha=1.0, round_out=False):
    """y[b, ..., c] = x[b, ..., c] * s[b, c] * alpha; x with leading dimension 1 is broadcast over the batch of s."""
    x = _f32c(x, "x")
    s = _f32c(s, "s")
    B, C = s.shape
    assert x.shape[-1] == C and x.shape[0] in (1, B), (x.shape, s.shape)
    P = x[0].numel() // C
    y = torch.empty((B,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
    _call("modulate", 0, 8 * y.numel(), lib().cb200_modulate, ptr(x), i64(0 if x.shape[0] == 1 and B > 1 else P * C), ptr(s),
          ptr(y), i32(B), i64(P), i32(C), f32(alpha), i32(_ro(round_out)), stream_ptr())
    return y


def mul_reduce(a, w):
    """out[b, c] = sum over the middle dimensions of a[b, ..., c] * w[b or 0, ..., c]."""
    a = _f32c(a, "a")
    w = _f32c(w, "w")
    B, C = a.shape[0], a.shape[-1]
    P = a[0].numel() // C
    assert w.shape[-1] == C and w[0].numel() == P * C and w.shape[0] in (1, B)
    out = torch.empty(B, C, device=a.device, dtype=torch.float32)
    _call("mul_reduce", 0, 8 * a.numel(), lib().cb200_mul_reduce, ptr(a), ptr(w), i64(0 if w.shape[0] == 1 and B > 1 else P * C),
          ptr(out), i32(B), i64(P), i32(C), stream_ptr())
    return out


def mod_epilogue(x, demod, noise, noise_weight, bias, slope=0.2, gain=2 ** 0.5, round_out=False):
    """y = lrelu(x * demod[b,c] + noise[b,p] * noise_weight[0] + bias[c]) * gain on x [B, H, W, C]."""
    x = _f32c(x, "x")
    B, C = x.shape[0], x.shape[-1]
    P = x[0].numel() // C
    if noise is not None:
        noise = _f32c(noise, "noise")
        assert noise.numel() == B * P, (noise.shape, x.shape)
    y = torch.empty_like(x)
    _call("mod_epilogue", 0, 8 * x.numel(), lib().cb200_mod_epilogue, ptr(x), ptr(demod), ptr(noise), ptr(noise_weight), ptr(bias),
          ptr(y), i32(B), i64(P), i32(C), f32(slope), f32(gain), i32(_ro(round_out)), stream_ptr())
    return y


def noise_grad(g, noise):
    """sum_{b,p} noise[b,p] * sum_c g[b,p,c] -> [1]."""
    g = _f32c(g, "g")
    noise = _f32c(noise, "noise")
    C = g.shape[-1]
    rows = g.numel() // C
    assert noise.numel() == rows
    out = torch.empty(1, device=g.device, dtype=torch.float32)
    _call("noise_grad", 0, 4 * g.numel(), lib().cb200_noise_grad, ptr(g), ptr(noise), ptr(out), i64(rows), i32(C), stream_ptr())
    return out


def _stddev_groups(B):
    G = min(B, 4)
    assert B % G == 0, "minibatch stddev needs batch % min(batch, 4) == 0 (discriminator.py:24-27)"
    return G, B // G


def stddev_fwd(x):
    x = _f32c(x, "x")
    B = x.shape[0]
    _, M = _stddev_groups(B)
    std = torch.empty(M, device=x.device, dtype=torch.float32)
    _call("stddev_fwd", 0, 4 * x.numel(), lib().cb200_stddev_fwd, ptr(x), ptr(std), i32(B), i64(x[0].numel()), stream_ptr())
    return std


def stddev_bwd(dstd, x):
    x = _f32c(x, "x")
    dstd = _f32c(dstd, "dstd")
    dx = torch.empty_like(x)
    _call("stddev_bwd", 0, 8 * x.numel(), lib().cb200_stddev_bwd, ptr(dstd), ptr(x), ptr(dx), i32(x.shape[0]), i64(x[0].numel()),
          stream_ptr())
    return dx


def stddev_bwd_bwd(gg, dstd, x):
    x = _f32c(x, "x")
    gg = _f32c(gg, "gg")
    dstd = _f32c(dstd, "dstd")
    d_dstd = torch.empty_like(dstd)
    d_x = torch.empty_like(x)
    _call("stddev_bwd_bwd", 0, 12 * x.numel(), lib().cb200_stddev_bwd_bwd, ptr(gg), ptr(dstd), ptr(x), ptr(d_dstd), ptr(d_x),
          i32(x.shape[0]), i64(x[0].numel()), stream_ptr())
    return d_dstd, d_x


def stddev_concat(x, std, cpad, round_out=False):
    """x [B, H, W, C], std [M] -> [B, H, W, cpad] with channel C = std[b % M] and zeros above."""
    x = _f32c(x, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, H, W, cpad, device=x.device, dtype=torch.float32)
    _call("stddev_concat", 0, 4 * (x.numel() + y.numel()), lib().cb200_stddev_concat, ptr(x), ptr(std), ptr(y), i32(B), i64(H * W),
          i32(C), i32(cpad), i32(_ro(round_out)), stream_ptr())
    return y


def stddev_split(dy, C):
    """dy [B, H, W, Cp] -> (dx [B, H, W, C], dstd [M])."""
    dy = _f32c(dy, "dy")
    B, H, W, Cp = dy.shape
    _, M = _stddev_groups(B)
    dx = torch.empty(B, H, W, C, device=dy.device, dtype=torch.float32)
    dstd = torch.empty(M, device=dy.device, dtype=torch.float32)
    _call("stddev_split", 0, 4 * (dy.numel() + dx.numel()), lib().cb200_stddev_split, ptr(dy), ptr(dx), ptr(dstd), i32(B),
          i64(H * W), i32(C), i32(Cp), stream_ptr())
    return dx, dstd


def rgb_to_nhwc(x, cpad=32, scale=1.0, shift=0.0, round_out=False):
    x = _f32c(x, "x")
    B, C, H, W = x.shape
    assert C == 3
    y = torch.empty(B, H, W, cpad, device=x.device, dtype=torch.float32)
    _call("rgb_to_nhwc", 0, 4 * (x.numel() + y.numel()), lib().cb200_rgb_to_nhwc, ptr(x), ptr(y), i32(B), i32(H), i32(W), i32(cpad),
          f32(scale), f32(shift), i32(_ro(round_out)), stream_ptr())
    return y


def nhwc_to_rgb(src, res=None, scale=1.0):
    src = _f32c(src, "src")
    B, H, W, cpad = src.shape
    out = torch.empty(B, 3, H, W, device=src.device, dtype=torch.float32)
    if res is not None:
        res = _f32c(res, "res")
        assert res.shape == out.shape
    _call("nhwc_to_rgb", 0, 4 * (src.numel() + out.numel()), lib().cb200_nhwc_to_rgb, ptr(src), ptr(res), ptr(out), i32(B), i32(H),
          i32(W), i32(cpad), f32(scale), stream_ptr())
    return out


def pixelnorm(x, round_out=False):
    x = _f32c(x, "x")
    rows, d = x.shape
    y = torch.empty_like(x)
    _call("pixelnorm", 0, 8 * x.numel(), lib().cb200_pixelnorm, ptr(x), ptr(y), i32(rows), i32(d), i32(_ro(round_out)),
          stream_ptr())
    return y


def row_sqsum(x):
    x = _f32c(x, "x")
    B = x.shape[0]
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    _call("row_sqsum", 0, 4 * x.numel(), lib().cb200_row_sqsum, ptr(x), ptr(out), i32(B), i64(x[0].numel()), stream_ptr())
    return out


def row_scale(x, s, alpha=1.0):
    x = _f32c(x, "x")
    s = _f32c(s, "s")
    y = torch.empty_like(x)
    _call("row_scale", 0, 8 * x.numel(), lib().cb200_row_scale, ptr(x), ptr(s), ptr(y), i32(x.shape[0]), i64(x[0].numel()),
          f32(alpha), stream_ptr())
    return y


def axpby(a, b=None, alpha=1.0, beta=1.0, gamma=0.0, round_out=False):
    a = _f32c(a, "a")
    if b is not None:
        b = _f32c(b, "b")
        assert b.shape == a.shape
    out = torch.empty_like(a)
    _call("axpby", 0, 8 * a.numel(), lib().cb200_axpby, ptr(a), ptr(b), ptr(out), i64(a.numel()), f32(alpha), f32(beta), f32(gamma),
          i32(_ro(round_out)), stream_ptr())
    return out


def pad_channels(x, cp):
    """x [..., C] -> [..., cp] zero-extended in the last dimension."""
    x = _f32c(x, "x")
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty(tuple(x.shape[:-1]) + (cp,), device=x.device, dtype=torch.float32)
    _call("pad_channels", 0, 4 * (x.numel() + y.numel()), lib().cb200_pad_channels, ptr(x), ptr(y), i64(rows), i32(C), i32(cp),
          stream_ptr())
    return y


class _EmaTensor(ctypes.Structure):
    _fields_ = [("dst", ctypes.c_void_p), ("src", ctypes.c_void_p), ("numel", ctypes.c_longlong)]


def ema_lerp(pairs, decay):
    """pairs: list of (dst, src) contiguous fp32 CUDA tensors; dst <- decay * dst + (1 - decay) * src, one launch
    per 64 tensors (utils.py:130-143 `accumulate`)."""
    arr = (_EmaTensor * len(pairs))()
    nbytes = 0
    for i, (d, s) in enumerate(pairs):
        assert d.is_contiguous() and s.is_contiguous() and d.numel() == s.numel()
        assert d.dtype == torch.float32 and s.dtype == torch.float32
        arr[i] = _EmaTensor(ptr(d).value, ptr(s).value, d.numel())             # ptr() refuses CPU tensors
        nbytes += 12 * d.numel()
    _call("ema_lerp", 0, nbytes, lib().cb200_ema_lerp, arr, i32(len(pairs)), f32(decay), stream_ptr())
