"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the C-ABI bindings that the StyleGAN2 autograd Functions call
(``contrad_b200.kernels`` / ``contrad_b200.sg2_kernels``), written independently of the CUDA sources from the
definitions in include/contrad_b200.h.

Two uses:
  * `-m "not gpu"` tests monkeypatch these into the product modules to exercise the HOST logic of
    ``contrad_b200.sg2_functional`` and the StyleGAN2 module mirrors (operator wiring, weight re-layouts, the
    double-backward structure behind the R1 penalty) against the oracle, without a GPU;
  * `-m gpu` tests use them as the per-kernel reference for the CUDA kernels.
The product never imports this file."""
import contextlib

import torch
import torch.nn.functional as F


def _lrelu(x, slope):
    return x if slope == 1.0 else torch.where(x > 0, x, x * slope)


# ------------------------------------------------------------------ contrad_b200.kernels stand-ins
def gemm_nt(a, bw, bias=None, slope=1.0, round_out=False, out=None, dact=None, colsum=None):
    y = a @ bw.t()
    if bias is not None:
        y = y + bias
    return _lrelu(y, slope)


def gemm_tn_wgrad(dy, x, out=None):
    return dy.t() @ x


def conv2d_nhwc_fwd(x, wmat, bias, ks, stride, slope=1.0, round_out=False):
    cout, cin = wmat.shape[0], x.shape[3]
    w = wmat.view(cout, ks, ks, cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=stride, padding=1)
    return _lrelu(y, slope).permute(0, 2, 3, 1).contiguous()


def conv2d_nhwc_dgrad(dy, wmat_t, in_shape, ks, stride, act_in=None, bias_out=None, slope=1.0, round_out=False, colsum=None):
    assert ks == 3 and stride == 1 and act_in is None and bias_out is None
    B, H, W, cin = in_shape
    cout = dy.shape[3]
    w = wmat_t.view(cin, 3, 3, cout).permute(3, 0, 1, 2)                # [Cout, Cin, kh, kw]
    dx = F.conv_transpose2d(dy.permute(0, 3, 1, 2), w, stride=1, padding=1)
    return dx.permute(0, 2, 3, 1).contiguous()


def conv2d_nhwc_wgrad(x, dy, ks, stride):
    assert ks == 3 and stride == 1
    cin, cout = x.shape[3], dy.shape[3]
    dw = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.permute(0, 3, 1, 2), stride=1, padding=1)
    return dw.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()


def colsum(x2d):
    return x2d.sum(0)


def round_tf32_(x):
    return x.clone()


# ------------------------------------------------------------------ contrad_b200.sg2_kernels stand-ins
def upfirdn_out_size(in_size, ksize, up, down, pad0, pad1):
    return (in_size * up + pad0 + pad1 - ksize) // down + 1


def upfirdn2d(x, fir, up, down, pad, out_hw=None, nhwc=True, flip=False, gain=1.0, round_out=False):
    """Zero-stuff by `up`, pad (negative = crop), correlate with the flipped kernel (flip=False) or the kernel
    as given (flip=True), keep every `down`-th sample; out_hw crops / zero-extends to the requested size."""
    xc = x.permute(0, 3, 1, 2) if nhwc else x
    N, C, H, W = xc.shape
    px0, px1, py0, py1 = pad
    kh, kw = fir.shape
    Ho = upfirdn_out_size(H, kh, up, down, py0, py1) if out_hw is None else out_hw[0]
    Wo = upfirdn_out_size(W, kw, up, down, px0, px1) if out_hw is None else out_hw[1]
    z = xc.new_zeros(N, C, H * up, W * up)
    z[:, :, ::up, ::up] = xc
    # enough bottom/right padding for the requested output size
    need_h = (Ho - 1) * down + kh
    need_w = (Wo - 1) * down + kw
    canvas = xc.new_zeros(N, C, max(need_h, 1), max(need_w, 1))
    ys, xs = max(py0, 0), max(px0, 0)           # where z[0] lands on the canvas
    zy0, zx0 = max(-py0, 0), max(-px0, 0)       # first z row/col kept when the pad is negative
    hh = min(z.shape[2] - zy0, canvas.shape[2] - ys)
    ww = min(z.shape[3] - zx0, canvas.shape[3] - xs)
    if hh > 0 and ww > 0:
        canvas[:, :, ys:ys + hh, xs:xs + ww] = z[:, :, zy0:zy0 + hh, zx0:zx0 + ww]
    k = fir if flip else torch.flip(fir, [0, 1])
    out = F.conv2d(canvas.reshape(N * C, 1, canvas.shape[2], canvas.shape[3]), (k * gain).view(1, 1, kh, kw))
    out = out[:, :, ::down, ::down][:, :, :Ho, :Wo].reshape(N, C, Ho, Wo)
    return out.permute(0, 2, 3, 1).contiguous() if nhwc else out.contiguous()


def patch_s2_gather(x, round_out=False):
    B, Hi, Wi, C = x.shape
    Ho, Wo = (Hi - 1) // 2, (Wi - 1) // 2
    u = x.new_empty(B, Ho, Wo, 9, C)
    for kh in range(3):
        for kw in range(3):
            u[:, :, :, kh * 3 + kw] = x[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2]
    return u


def patch_s2_scatter(u, round_out=False):
    B, Ho, Wo, _, C = u.shape
    x = u.new_zeros(B, 2 * Ho + 1, 2 * Wo + 1, C)
    for kh in range(3):
        for kw in range(3):
            x[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2] += u[:, :, :, kh * 3 + kw]
    return x


def bias_act(x, bias, slope, gain, res=None, round_out=False):
    t = x if bias is None else x + bias
    y = _lrelu(t, slope) * gain
    return y if res is None else y + res


def bias_act_grad(g, ref, bias, slope, gain, round_out=False):
    t = ref if bias is None else ref + bias
    return g * torch.where(t > 0, torch.full_like(t, gain), torch.full_like(t, gain * slope))


def _bshape(s, x):
    return s.view(s.shape[0], *([1] * (x.dim() - 2)), s.shape[1])


def modulate(x, s, batch=None, alpha=1.0, round_out=False):
    return (x * _bshape(s, x) * alpha).contiguous()


def mul_reduce(a, w):
    return (a * w).reshape(a.shape[0], -1, a.shape[-1]).sum(1)


def mod_epilogue(x, demod, noise, noise_weight, bias, slope=0.2, gain=2 ** 0.5, round_out=False):
    t = x
    if demod is not None:
        t = t * _bshape(demod, x)
    if noise is not None:
        t = t + noise.reshape(x.shape[0], x.shape[1], x.shape[2], 1) * noise_weight.reshape(())
    if bias is not None:
        t = t + bias
    return _lrelu(t, slope) * gain


def noise_grad(g, noise):
    return (g.sum(-1).reshape(-1) * noise.reshape(-1)).sum().reshape(1)


def _groups(B):
    G = min(B, 4)
    return G, B // G


def stddev_fwd(x):
    B = x.shape[0]
    G, M = _groups(B)
    v = x.reshape(G, M, -1)
    return torch.sqrt(v.var(0, unbiased=False) + 1e-8).mean(1)


def stddev_bwd(dstd, x):
    xs = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        (dx,) = torch.autograd.grad(stddev_fwd(xs), xs, dstd)
    return dx


def stddev_bwd_bwd(gg, dstd, x):
    xs = x.detach().clone().requires_grad_(True)
    ds = dstd.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        (dx,) = torch.autograd.grad(stddev_fwd(xs), xs, ds, create_graph=True)
        d_ds, d_x = torch.autograd.grad(dx, [ds, xs], gg)
    return d_ds, d_x


def stddev_concat(x, std, cpad, round_out=False):
    B, H, W, C = x.shape
    _, M = _groups(B)
    y = x.new_zeros(B, H, W, cpad)
    y[..., :C] = x
    y[..., C] = std[torch.arange(B) % M].view(B, 1, 1)
    return y


def stddev_split(dy, C):
    B = dy.shape[0]
    _, M = _groups(B)
    dstd = dy.new_zeros(M)
    dstd.index_add_(0, torch.arange(B) % M, dy[..., C].reshape(B, -1).sum(1))
    return dy[..., :C].contiguous(), dstd


def rgb_to_nhwc(x, cpad=32, scale=1.0, shift=0.0, round_out=False):
    B, _, H, W = x.shape
    y = x.new_zeros(B, H, W, cpad)
    y[..., :3] = (x * scale + shift).permute(0, 2, 3, 1)
    return y


def nhwc_to_rgb(src, res=None, scale=1.0):
    out = (src[..., :3] * scale).permute(0, 3, 1, 2).contiguous()
    return out if res is None else out + res


def pixelnorm(x, round_out=False):
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def row_sqsum(x):
    return x.reshape(x.shape[0], -1).pow(2).sum(1)


def row_scale(x, s, alpha=1.0):
    return x * (s * alpha).view(-1, *([1] * (x.dim() - 1)))


def axpby(a, b=None, alpha=1.0, beta=1.0, gamma=0.0, round_out=False):
    y = alpha * a + gamma
    return y if b is None else y + beta * b


def pad_channels(x, cp):
    return F.pad(x, (0, cp - x.shape[-1]))


def ema_lerp(pairs, decay):
    for d, s in pairs:
        d.mul_(decay).add_(s, alpha=1 - decay)


def split_tf32(x, mode):
    """include/contrad_b200.h cb200_split_tf32: hi = rn_tf32(x), lo = rn_tf32(x - hi); mode 0 hi|lo|hi, 1 hi|hi, 2 [hi ; lo]."""
    from contrad_b200 import kernels as K
    hi = K.round_tf32(x)
    lo = K.round_tf32(x - hi)
    if mode == 0:
        return torch.cat([hi, lo, hi], dim=-1).contiguous()
    if mode == 1:
        return torch.cat([hi, hi], dim=-1).contiguous()
    return torch.stack([hi, lo], dim=0).contiguous()


K_NAMES = ("gemm_nt", "gemm_tn_wgrad", "conv2d_nhwc_fwd", "conv2d_nhwc_dgrad", "conv2d_nhwc_wgrad", "colsum", "round_tf32_",
           "split_tf32")
S_NAMES = ("upfirdn2d", "patch_s2_gather", "patch_s2_scatter", "bias_act", "bias_act_grad", "modulate", "mul_reduce",
           "mod_epilogue", "noise_grad", "stddev_fwd", "stddev_bwd", "stddev_bwd_bwd", "stddev_concat", "stddev_split",
           "rgb_to_nhwc", "nhwc_to_rgb", "pixelnorm", "row_sqsum", "row_scale", "axpby", "ema_lerp", "pad_channels")


@contextlib.contextmanager
def patched(exact_weights=True):
    """Route the product's kernel bindings to the CPU stand-ins above (tests only).  exact_weights: also turn
    the TF32 rounding of weight operands into the identity, so results are plain fp32."""
    import sys
    from contrad_b200 import kernels as K
    from contrad_b200 import sg2_kernels as S
    me = sys.modules[__name__]
    saved = [(K, n, getattr(K, n)) for n in K_NAMES] + [(S, n, getattr(S, n)) for n in S_NAMES]
    saved.append((K, "round_tf32", K.round_tf32))
    try:
        for mod, n, _ in saved[:-1]:
            setattr(mod, n, getattr(me, n))
        if exact_weights:
            K.round_tf32 = lambda t: t.contiguous()
        yield
    finally:
        for mod, n, fn in saved:
            setattr(mod, n, fn)
