"""Differentiable primitives of the StyleGAN2 side of the path.

Unlike the SNDCGAN Functions (one monolithic Function per network), these are small operators whose backward is
itself expressed with the same operators, so that autograd can differentiate the *backward* pass: the R1 penalty
(`train_stylegan2.py:106-113`, `train_stylegan2_contraD.py:129-136`) takes `autograd.grad(..., create_graph=True)`
of D w.r.t. its input and back-propagates through that gradient.  The families are closed under differentiation:

    MmNT / MmNN / MmTN           x @ w^T, g @ w, g^T @ x             (tcgen05 GEMMs)
    Conv3x3 / Dgrad / Wgrad      3x3 stride-1 pad-1 NHWC convolution  (tcgen05 implicit GEMMs)
    UpFirDn                      its backward is UpFirDn with up <-> down and the flipped kernel
    PatchS2 / PatchS2T           3x3 stride-2 patch gather / scatter
    BiasAct / BiasActGrad        FusedLeakyReLU; the mask does not depend on the cotangent
    Stddev / StddevBwd, StddevConcat / StddevSplit, Rgb2Nhwc / Nhwc2Rgb, Axpby, Modulate

Every forward/backward is one call into the C ABI (`contrad_b200.kernels`, `contrad_b200.sg2_kernels`) on NHWC
activations; torch ops appear only on weight-shaped tensors (scaling, re-layout, TF32 rounding)."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K
from . import precision
from . import sg2_kernels as S


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _cached(w, key, make):
    """Per-tensor memo of derived weight layouts: the forward GEMM, the data-gradient GEMM and the second-order passes
    of one layer all receive the SAME (scaled) weight tensor, so each re-layout is computed once per forward.  The memo
    lives on the tensor object and is dropped with it (or when the tensor is modified in place)."""
    memo = getattr(w, "_cb200_memo", None)
    if memo is None or memo[0] != w._version:
        memo = (w._version, {})
        try:
            w._cb200_memo = memo
        except AttributeError:
            return make()
    if key not in memo[1]:
        memo[1][key] = make()
    return memo[1][key]


_WEIGHT_EPOCH = [0]


def bump_weight_epoch():
    """Called by optimisers / EMA updates that write parameters through the C ABI (raw pointers: torch's version
    counters do not see them), so that `weight_memo` entries derived from the old values are dropped."""
    _WEIGHT_EPOCH[0] += 1


def weight_memo(param, key, make):
    """Per-parameter memo of tensors derived from it (runtime-scaled weight, zero-padded / permuted variants ...),
    valid until the parameter changes (version counter or `bump_weight_epoch`) or the autograd mode differs.  A training
    step evaluates D up to four times with the same weights (fakes, reals, the R1 pass, the G step): the derived
    tensors - and, through `_cached`, their GEMM re-layouts - are then built once per optimiser step instead of once per
    forward.  The entries keep their autograd history, so gradients of all the forwards flow into the parameter; only
    results of scale / pad / permute / reshape chains may be stored (their backward nodes hold no tensors, so they can
    be back-propagated through any number of times, e.g. under gradient accumulation).

    Forwards under `torch.no_grad()` are NEVER served from the memo: the EMA generator is only ever evaluated that way
    (evaluate/gan.py:57-58, FID), and the reference's `utils.accumulate` (utils.py:130-143) updates it through
    `param.data.mul_().add_()`, which neither torch's version counter nor `bump_weight_epoch` sees - a memo entry
    would freeze g_ema at its first evaluation."""
    if not torch.is_grad_enabled():
        return make()
    tag = (param._version, _WEIGHT_EPOCH[0], param.requires_grad, torch.is_grad_enabled())
    memo = getattr(param, "_cb200_wmemo", None)
    if memo is None or memo[0] != tag:
        memo = (tag, {})
        param._cb200_wmemo = memo
    if key not in memo[1]:
        memo[1][key] = make()
    return memo[1][key]


class LinearMap(Function):
    """y = fwd(w) for a LINEAR map given with its adjoint `bwd` (two callables closing over Python scalars / shapes
    only): weight scaling, zero-padding, re-layout.  The node saves no tensors, so a memoised result can be
    back-propagated through repeatedly; `bwd` runs ordinary torch ops, so higher-order differentiation works."""

    @staticmethod
    def forward(ctx, w, fwd, bwd):
        ctx.bwd = bwd
        return fwd(w)

    @staticmethod
    def backward(ctx, g):
        return ctx.bwd(g), None, None


def _rw(w):
    """Weight-shaped operand of a tensor-core GEMM: contiguous and rounded to TF32 (nearest)."""
    return _cached(w, "rw", lambda: K.round_tf32(_c(w.detach())))


def _rw_t(w):
    return _cached(w, "rw_t", lambda: K.round_tf32(_c(w.detach().t())))


# ---- strict precision ("3xTF32", contrad_b200/precision.py level "full"): every GEMM Function below evaluates
# (a_hi + a_lo)(b_hi + b_lo) ~ a_hi b_hi + a_lo b_hi + a_hi b_lo through the SAME tcgen05 entry points on operands
# concatenated along the reduction axis; producers stop rounding (sg2_kernels._ro, RoundTF32).  Because the Function
# families are closed under differentiation, the R1 double backward is compensated as well.
def _strict():
    return precision.strict_full()


def _hi_lo(t):
    hi = K.round_tf32(t)
    return hi, K.round_tf32(t - hi)


def _w3_cols(w2d, key):
    """[N, K] weight -> [N, 3K] = w_hi | w_hi | w_lo (pairs with an activation split as hi | lo | hi along K)."""
    def make():
        hi, lo = _hi_lo(_c(w2d.detach()))
        return torch.cat([hi, hi, lo], dim=1).contiguous()
    return _cached(w2d, key, make)


def _rows3(t, first):
    """Three row-stacked copies for reductions over the ROW axis: first=True -> [hi ; lo ; hi], else [hi ; hi ; lo]."""
    parts = K.split_tf32(_c(t), 2)
    hi, lo = parts[0], parts[1]
    return torch.cat([hi, lo, hi] if first else [hi, hi, lo], dim=0)


def _pack_fwd(w):
    return _cached(w, "pack_fwd", lambda: K.pack_fwd_weight(_rw(w)))


def _pack_dgrad(w):
    return _cached(w, "pack_dgrad", lambda: K.pack_dgrad_weight(_rw(w), 1))


def _round_operand(x):
    """Cotangent about to enter a GEMM: rounded to TF32 (straight-through) - or left alone in the strict precision mode,
    where the consuming GEMM Function splits it into hi + lo itself."""
    return x if _strict() else RoundTF32.apply(x)


class RoundTF32(Function):
    """Round an activation-shaped GEMM operand to TF32 (nearest); identity gradient.  Forward activations are rounded
    by the kernels that produce them; cotangents arriving from autograd (sums of branches, loss gradients) are not,
    and the tensor core would TRUNCATE them - a systematic shrink of ~2^-11 per GEMM that adds up along the backward
    pass - so every backward closure rounds its incoming cotangent once."""

    @staticmethod
    def forward(ctx, x):
        return K.round_tf32_(_c(x))

    @staticmethod
    def backward(ctx, g):
        return g


# ------------------------------------------------------------------------------------------------ GEMM family
class MmNT(Function):
    """out[M,N] = a[M,K] @ w[N,K]^T (+ bias[N])          N, K multiples of 32."""

    @staticmethod
    def forward(ctx, a, w, bias=None):
        a = _c(a)
        ctx.save_for_backward(a, w)
        ctx.has_bias = bias is not None
        if _strict():
            return K.gemm_nt(K.split_tf32(a, 0), _w3_cols(w, "w3"), None if bias is None else _c(bias.detach()))
        return K.gemm_nt(a, _rw(w), None if bias is None else _c(bias.detach()))

    @staticmethod
    def backward(ctx, dy):
        a, w = ctx.saved_tensors
        dy = _round_operand(dy)
        da = MmNN.apply(dy, w) if ctx.needs_input_grad[0] else None
        dw = MmTN.apply(dy, a) if ctx.needs_input_grad[1] else None
        db = ColSum.apply(dy) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return da, dw, db


class MmNN(Function):
    """out[M,K] = g[M,N] @ w[N,K]."""

    @staticmethod
    def forward(ctx, g, w):
        g = _c(g)
        ctx.save_for_backward(g, w)
        if _strict():
            wt3 = _cached(w, "wt3", lambda: (lambda hi, lo: torch.cat([hi, hi, lo], dim=1).contiguous())(*_hi_lo(_c(w.detach().t()))))
            return K.gemm_nt(K.split_tf32(g, 0), wt3)
        return K.gemm_nt(g, _rw_t(w))

    @staticmethod
    def backward(ctx, gg):
        g, w = ctx.saved_tensors
        gg = _round_operand(gg)
        d_g = MmNT.apply(gg, w) if ctx.needs_input_grad[0] else None
        d_w = MmTN.apply(g, gg) if ctx.needs_input_grad[1] else None
        return d_g, d_w


class MmTN(Function):
    """out[N,K] = g[M,N]^T @ a[M,K]   (weight gradients; reduction over the rows)."""

    @staticmethod
    def forward(ctx, g, a):
        g, a = _c(g), _c(a)
        ctx.save_for_backward(g, a)
        n, k = g.shape[1], a.shape[1]
        if _strict():                                  # reduction over the rows: stack [g_hi; g_lo; g_hi] against [a_hi; a_hi; a_lo]
            g, a = _rows3(g, True), _rows3(a, False)
        if n % 128 == 0:
            return K.gemm_tn_wgrad(g, a)
        if k % 128 == 0:                               # the kernel wants 128 | rows of the result: compute the transpose
            return K.gemm_tn_wgrad(a, g).t().contiguous()
        # thin layers on both sides (e.g. 32 -> 64 channels at 512x512): zero-extend the cotangent to 128 columns
        return K.gemm_tn_wgrad(S.pad_channels(g, (n + 127) // 128 * 128), a)[:n].contiguous()

    @staticmethod
    def backward(ctx, gw):
        g, a = ctx.saved_tensors
        gw = _c(gw)
        d_g = MmNT.apply(a, gw) if ctx.needs_input_grad[0] else None
        d_a = MmNN.apply(g, gw) if ctx.needs_input_grad[1] else None
        return d_g, d_a


class ColSum(Function):
    """out[N] = sum_m x[m, N] (bias gradients)."""

    @staticmethod
    def forward(ctx, x):
        return K.colsum(_c(x).view(-1, x.shape[-1]))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        raise NotImplementedError("ColSum is a leaf of the gradient graph (bias gradients are not differentiated again)")


# ------------------------------------------------------------------------------------------------ 3x3 convolution family
class Conv3x3(Function):
    """y[B,H,W,Cout] = conv(x[B,H,W,Cin], w[Cout,Cin,3,3]), stride 1, padding 1 (`F.conv2d`, layers.py:115-121)."""

    @staticmethod
    def forward(ctx, x, w):
        x = _c(x)
        ctx.save_for_backward(x, w)
        if _strict():                                  # input channels concatenated: x_hi | x_lo | x_hi against w_hi | w_hi | w_lo
            def pack3():
                hi, lo = _hi_lo(_c(w.detach()))
                return K.pack_fwd_weight(torch.cat([hi, hi, lo], dim=1).contiguous())
            return K.conv2d_nhwc_fwd(K.split_tf32(x, 0), _cached(w, "pack_fwd3", pack3), None, 3, 1)
        return K.conv2d_nhwc_fwd(x, _pack_fwd(w), None, 3, 1)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _round_operand(dy)
        dx = Conv3x3Dgrad.apply(dy, w) if ctx.needs_input_grad[0] else None
        dw = Conv3x3Wgrad.apply(x, dy) if ctx.needs_input_grad[1] else None
        return dx, dw


class Conv3x3Dgrad(Function):
    """dx[B,H,W,Cin] = conv^T(dy[B,H,W,Cout], w[Cout,Cin,3,3])."""

    @staticmethod
    def forward(ctx, dy, w):
        dy = _c(dy)
        ctx.save_for_backward(dy, w)
        B, H, W, _ = dy.shape
        if _strict():                                  # output channels (the reduction axis here) concatenated
            def pack3():
                hi, lo = _hi_lo(_c(w.detach()))
                return K.pack_dgrad_weight(torch.cat([hi, hi, lo], dim=0).contiguous(), 1)
            return K.conv2d_nhwc_dgrad(K.split_tf32(dy, 0), _cached(w, "pack_dgrad3", pack3), (B, H, W, w.shape[1]), 3, 1)
        return K.conv2d_nhwc_dgrad(dy, _pack_dgrad(w), (B, H, W, w.shape[1]), 3, 1)

    @staticmethod
    def backward(ctx, g):
        dy, w = ctx.saved_tensors
        g = _round_operand(g)
        d_dy = Conv3x3.apply(g, w) if ctx.needs_input_grad[0] else None
        d_w = Conv3x3Wgrad.apply(g, dy) if ctx.needs_input_grad[1] else None
        return d_dy, d_w


class Conv3x3Wgrad(Function):
    """dw[Cout,Cin,3,3] = sum over pixels of dy (x) shifted x."""

    @staticmethod
    def forward(ctx, x, dy):
        x, dy = _c(x), _c(dy)
        ctx.save_for_backward(x, dy)
        cout, cin = dy.shape[3], x.shape[3]
        if _strict():                                  # reduction over pixels: stack the batch, [x_hi; x_hi; x_lo] vs [dy_hi; dy_lo; dy_hi]
            x, dy = _rows3(x, False), _rows3(dy, True)
        if cout % 128:          # the kernel tiles Cout by 128: 32- / 64-channel layers present dY zero-extended
            dwp = K.conv2d_nhwc_wgrad(x, S.pad_channels(dy, (cout + 127) // 128 * 128), 3, 1)[:cout]
        else:
            dwp = K.conv2d_nhwc_wgrad(x, dy, 3, 1)
        return dwp.view(cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous()

    @staticmethod
    def backward(ctx, gw):
        x, dy = ctx.saved_tensors
        d_x = Conv3x3Dgrad.apply(dy, gw) if ctx.needs_input_grad[0] else None
        d_dy = Conv3x3.apply(x, gw) if ctx.needs_input_grad[1] else None
        return d_x, d_dy


# ------------------------------------------------------------------------------------------------ upfirdn2d
class UpFirDn(Function):
    """op/upfirdn2d.py:86-142.  pad = (x0, x1, y0, y1); out_hw None = the reference's output size."""

    @staticmethod
    def forward(ctx, x, fir, up, down, pad, out_hw, flip, nhwc, gain, round_out):
        x = _c(x)
        ctx.save_for_backward(fir)
        ctx.cfg = (up, down, tuple(pad), flip, nhwc, gain)
        ctx.in_hw = (x.shape[1], x.shape[2]) if nhwc else (x.shape[2], x.shape[3])
        y = S.upfirdn2d(x, fir, up, down, pad, out_hw=out_hw, nhwc=nhwc, flip=flip, gain=gain, round_out=round_out)
        ctx.out_hw = (y.shape[1], y.shape[2]) if nhwc else (y.shape[2], y.shape[3])
        return y

    @staticmethod
    def backward(ctx, dy):
        (fir,) = ctx.saved_tensors
        up, down, (px0, px1, py0, py1), flip, nhwc, gain = ctx.cfg
        kh, kw = fir.shape
        (in_h, in_w), (out_h, out_w) = ctx.in_hw, ctx.out_hw
        # op/upfirdn2d.py:122-127
        g_pad = (kw - px0 - 1, in_w * up - out_w * down + px0 - up + 1, kh - py0 - 1, in_h * up - out_h * down + py0 - up + 1)
        dx = UpFirDn.apply(dy, fir, down, up, g_pad, ctx.in_hw, not flip, nhwc, gain, False)
        return dx, None, None, None, None, None, None, None, None, None


def upfirdn2d_nhwc(x, fir, up=1, down=1, pad=(0, 0), gain=1.0, round_out=False):
    return UpFirDn.apply(x, fir, up, down, (pad[0], pad[1], pad[0], pad[1]), None, False, True, gain, round_out)


# ------------------------------------------------------------------------------------------------ stride-2 patches
class PatchS2(Function):
    """x [B,2Ho+1,2Wo+1,C] -> [B,Ho,Wo,9,C]."""

    @staticmethod
    def forward(ctx, x, round_out):
        return S.patch_s2_gather(_c(x), round_out=round_out)

    @staticmethod
    def backward(ctx, du):
        return PatchS2T.apply(du, False), None


class PatchS2T(Function):
    """u [B,Ho,Wo,9,C] -> [B,2Ho+1,2Wo+1,C]."""

    @staticmethod
    def forward(ctx, u, round_out):
        return S.patch_s2_scatter(_c(u), round_out=round_out)

    @staticmethod
    def backward(ctx, dx):
        return PatchS2.apply(dx, False), None


# ------------------------------------------------------------------------------------------------ bias + leaky relu
class BiasAct(Function):
    """y = lrelu_slope(x + bias) * gain (+ res)   (FusedLeakyReLU, op/fused_act.py:74-94; channel = last dim)."""

    @staticmethod
    def forward(ctx, x, bias, res, slope, gain, round_out):
        x = _c(x)
        ctx.save_for_backward(x, bias)
        ctx.cfg = (slope, gain)
        return S.bias_act(x, None if bias is None else _c(bias.detach()), slope, gain, res=res, round_out=round_out)

    @staticmethod
    def backward(ctx, dy):
        x, bias = ctx.saved_tensors
        slope, gain = ctx.cfg
        dx = db = None
        if ctx.needs_input_grad[0] or (bias is not None and ctx.needs_input_grad[1]):
            dx = BiasActGrad.apply(dy, x, bias, slope, gain)
            if bias is not None and ctx.needs_input_grad[1]:
                db = ColSum.apply(dx)
        dres = dy if ctx.needs_input_grad[2] else None
        return dx, db, dres, None, None, None


class BiasActGrad(Function):
    """g * lrelu'(ref + bias) * gain: backward of BiasAct and, being linear in g, its own backward."""

    @staticmethod
    def forward(ctx, g, ref, bias, slope, gain):
        ctx.save_for_backward(ref, bias)
        ctx.cfg = (slope, gain)
        return S.bias_act_grad(_c(g), ref, None if bias is None else _c(bias.detach()), slope, gain, round_out=True)

    @staticmethod
    def backward(ctx, gg):
        ref, bias = ctx.saved_tensors
        slope, gain = ctx.cfg
        return BiasActGrad.apply(gg, ref, bias, slope, gain), None, None, None, None


# ------------------------------------------------------------------------------------------------ minibatch stddev
class Stddev(Function):
    """x [B,H,W,C] -> std [B / min(B,4)]   (discriminator.py:22-31)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.save_for_backward(x)
        return S.stddev_fwd(x)

    @staticmethod
    def backward(ctx, dstd):
        (x,) = ctx.saved_tensors
        return StddevBwd.apply(_c(dstd), x)


class StddevBwd(Function):
    @staticmethod
    def forward(ctx, dstd, x):
        ctx.save_for_backward(dstd, x)
        return S.stddev_bwd(dstd, x)

    @staticmethod
    @once_differentiable
    def backward(ctx, gg):
        dstd, x = ctx.saved_tensors
        d_dstd, d_x = S.stddev_bwd_bwd(_c(gg), dstd, x)
        return d_dstd, d_x


class StddevConcat(Function):
    """(x [B,H,W,C], std [M]) -> [B,H,W,cpad]: `torch.cat([input, stddev], 1)` (discriminator.py:31-33), zero padded."""

    @staticmethod
    def forward(ctx, x, std, cpad, round_out):
        ctx.c = x.shape[-1]
        return S.stddev_concat(_c(x), _c(std), cpad, round_out=round_out)

    @staticmethod
    def backward(ctx, dy):
        dx, dstd = StddevSplit.apply(dy, ctx.c)
        return dx, dstd, None, None


class StddevSplit(Function):
    @staticmethod
    def forward(ctx, dy, c):
        ctx.cpad = dy.shape[-1]
        return S.stddev_split(_c(dy), c)

    @staticmethod
    def backward(ctx, g_dx, g_dstd):
        return StddevConcat.apply(g_dx, g_dstd, ctx.cpad, False), None


# ------------------------------------------------------------------------------------------------ layout / elementwise
class Rgb2Nhwc(Function):
    """x [B,3,H,W] -> [B,H,W,cpad], y = x*scale + shift in the first three channels (`input * 2. - 1.`)."""

    @staticmethod
    def forward(ctx, x, cpad, scale, shift, round_out):
        ctx.scale = scale
        return S.rgb_to_nhwc(_c(x), cpad=cpad, scale=scale, shift=shift, round_out=round_out)

    @staticmethod
    def backward(ctx, dy):
        return Nhwc2Rgb.apply(dy, None, ctx.scale), None, None, None, None


class Nhwc2Rgb(Function):
    """(src [B,H,W,cpad], res [B,3,H,W] or None) -> src[..., :3]*scale (+ res) as NCHW."""

    @staticmethod
    def forward(ctx, src, res, scale):
        ctx.cpad, ctx.scale = src.shape[-1], scale
        return S.nhwc_to_rgb(_c(src), res=res, scale=scale)

    @staticmethod
    def backward(ctx, dout):
        dsrc = Rgb2Nhwc.apply(dout, ctx.cpad, ctx.scale, 0.0, True) if ctx.needs_input_grad[0] else None
        return dsrc, (dout if ctx.needs_input_grad[1] else None), None


class Axpby(Function):
    """alpha*a + beta*b + gamma."""

    @staticmethod
    def forward(ctx, a, b, alpha, beta, gamma):
        ctx.cfg = (alpha, beta)
        return S.axpby(_c(a), b, alpha, beta, gamma)

    @staticmethod
    def backward(ctx, dy):
        alpha, beta = ctx.cfg
        da = (dy if alpha == 1.0 else Axpby.apply(dy, None, alpha, 0.0, 0.0)) if ctx.needs_input_grad[0] else None
        db = (dy if beta == 1.0 else Axpby.apply(dy, None, beta, 0.0, 0.0)) if ctx.needs_input_grad[1] else None
        return da, db, None, None, None


class RowSqSum(Function):
    """x [B, ...] -> sum of squares per sample (`grad.pow(2).reshape(B,-1).sum(1)`, train_stylegan2.py:112)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.save_for_backward(x)
        return S.row_sqsum(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return RowScale.apply(x, g, 2.0)


class RowScale(Function):
    """y[b, ...] = x[b, ...] * s[b] * alpha."""

    @staticmethod
    def forward(ctx, x, s, alpha):
        ctx.save_for_backward(s)
        ctx.alpha = alpha
        return S.row_scale(_c(x), _c(s), alpha)

    @staticmethod
    def backward(ctx, dy):
        (s,) = ctx.saved_tensors
        d_x = RowScale.apply(dy, s, ctx.alpha) if ctx.needs_input_grad[0] else None
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("RowScale: the scale is a leaf (cotangent of the R1 penalty)")
        return d_x, None, None


# ------------------------------------------------------------------------------------------------ modulated convolution
class Modulate(Function):
    """y[b,h,w,c] = x[b or 0,h,w,c] * s[b,c]   (generator.py:55-56, moved from the weights to the activations)."""

    @staticmethod
    def forward(ctx, x, s, round_out):
        x, s = _c(x), _c(s)
        ctx.save_for_backward(x, s)
        return S.modulate(x, s, round_out=round_out)

    @staticmethod
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        dy = _c(dy)
        dx = ds = None
        if ctx.needs_input_grad[0]:
            dx = Modulate.apply(dy, s, True)
            if x.shape[0] == 1 and dy.shape[0] > 1:          # ConstantInput: sum over the batch it was broadcast to
                dx = ColSum.apply(dx.view(dy.shape[0], -1)).view(x.shape)
        if ctx.needs_input_grad[1]:
            ds = MulReduce.apply(dy, x)
        return dx, ds, None


class MulReduce(Function):
    """out[b,c] = sum_{h,w} a[b,h,w,c] * w[b or 0,h,w,c]."""

    @staticmethod
    def forward(ctx, a, w):
        return S.mul_reduce(_c(a), _c(w))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        raise NotImplementedError("MulReduce is only used in first-order backward passes (the generator)")


class ModEpilogue(Function):
    """y = lrelu(x * demod[b,c] + noise[b,h,w] * noise_weight + bias[c]) * sqrt(2): demodulation (generator.py:58-60),
    NoiseInjection (:85-94) and FusedLeakyReLU (:112-116) in one pass.  First-order only (generator)."""

    SLOPE, GAIN = 0.2, 2 ** 0.5

    @staticmethod
    def forward(ctx, x, demod, noise, noise_weight, bias, round_out):
        x = _c(x)
        y = S.mod_epilogue(x, None if demod is None else _c(demod), noise, _c(noise_weight.detach()), _c(bias.detach()),
                           ModEpilogue.SLOPE, ModEpilogue.GAIN, round_out=round_out)
        ctx.save_for_backward(x, demod, noise, y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, demod, noise, y = ctx.saved_tensors
        ng = ctx.needs_input_grad
        dpre = S.bias_act_grad(_c(dy), y, None, ModEpilogue.SLOPE, ModEpilogue.GAIN)
        dx = ddemod = dnw = dbias = None
        if ng[0]:
            dx = S.modulate(dpre, demod, round_out=True) if demod is not None else dpre
        if demod is not None and ng[1]:
            ddemod = S.mul_reduce(dpre, x)
        if ng[3]:
            dnw = S.noise_grad(dpre, noise)
        if ng[4]:
            dbias = K.colsum(dpre.view(-1, dpre.shape[-1]))
        return dx, ddemod, None, dnw, dbias, None


def pixelnorm(z):
    """PixelNorm (layers.py:15-20) of the latent input; z carries no gradient on the training path."""
    if z.requires_grad:
        raise NotImplementedError("pixelnorm: gradients w.r.t. the latent input are not on the ContraD hot path")
    return S.pixelnorm(_c(z), round_out=True)
