"""torch.autograd.Functions of the ContraD hot path.  Every forward/backward is a sequence of calls
into the C ABI (contrad_b200.kernels); there is no ATen math on the activation path and no CPU path.

Layout convention inside the discriminator: activations are NHWC fp32, TF32-rounded by the producing
epilogue; weights are consumed as packed GEMM matrices written by the spectral-norm kernels.
"""
import os

import torch
from torch.autograd import Function

from . import kernels as K
from . import precision


# bias gradients as column sums fused into the producing GEMM epilogue (CB200_FUSED_COLSUM=0: separate kernel)
_FUSE_COLSUM = os.environ.get("CB200_FUSED_COLSUM", "1") != "0"


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ---- strict precision ("3xTF32", contrad_b200/precision.py): weight-side helpers.  Activations are split by
# cb200_split_tf32 (K.split_tf32); weights are small, so their hi / lo parts are built with torch ops on the packed
# matrices.  `groups` = how the reduction axis of the pack is organised: [rows, taps, C] with the split applied to C.
def _hi_lo(t):
    hi = K.round_tf32(t)
    return hi, K.round_tf32(t - hi)


def _split_weight_fwd(pack, taps):
    """[rows, taps*C] -> [rows, taps*3C] = per tap [w_hi | w_hi | w_lo]: pairs with activations split as hi | lo | hi."""
    rows = pack.shape[0]
    hi, lo = _hi_lo(pack.reshape(rows, taps, -1))
    return torch.cat([hi, hi, lo], dim=2).reshape(rows, -1).contiguous()


def _split_weight_bwd(pack, taps):
    """[rows, taps*C] -> [rows, taps*2C] = per tap [w_hi | w_lo]: pairs with a gradient duplicated as g | g."""
    rows = pack.shape[0]
    hi, lo = _hi_lo(pack.reshape(rows, taps, -1))
    return torch.cat([hi, lo], dim=2).reshape(rows, -1).contiguous()


# ------------------------------------------------------------------------------------------------
# augmentation  (reference: augment/__init__.py:106-112 chain; backward = its autograd)
# ------------------------------------------------------------------------------------------------
class AugmentSimCLRFn(Function):
    """Images up to 64x64 take the one-image-per-CTA kernels; larger ones (the 512x512 StyleGAN2 configs) the
    global-memory kernels, which also hand the per-image contrast means from the forward to the backward pass.
    `force_large` selects the second path for any size (parity tests)."""

    @staticmethod
    def forward(ctx, x, params, order, force_large=False):
        x = _c(x)
        ctx.order = order
        ctx.large = bool(force_large) or K.augment_needs_large_path(x.shape[2], x.shape[3])
        if ctx.large:
            y, means = K.augment_simclr_large_fwd(x, params, order)
            ctx.save_for_backward(x, params, means)
            return y
        ctx.save_for_backward(x, params)
        return K.augment_simclr_fwd(x, params, order)

    @staticmethod
    def backward(ctx, dy):
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.large:
                x, params, means = ctx.saved_tensors
                dx = K.augment_simclr_large_bwd(x, _c(dy), params, ctx.order, means)
            else:
                x, params = ctx.saved_tensors
                dx = K.augment_simclr_bwd(x, _c(dy), params, ctx.order)
        return dx, None, None, None


class AugmentSimCLRMixedFn(Function):
    """Row f3: augment(cat[ToTensor(x_u8)] * reps + [x_f32]) in one launch, without materialising the conversion or the
    concatenation (datasets.py:10-21, training/gan/contrad.py:38-41).  Only `x_f32` is differentiable; its gradient runs
    the ordinary backward kernels on the parameter columns of its views."""

    @staticmethod
    def forward(ctx, x_u8, n_u8_views, x_f32, params, order):
        y, means = K.augment_simclr_mixed_fwd(x_u8, n_u8_views, x_f32, params, order)
        ctx.order, ctx.n_u8_views = order, n_u8_views
        ctx.has_f32 = x_f32 is not None and x_f32.shape[0] > 0
        if ctx.has_f32 and x_f32.requires_grad:
            x_f32 = _c(x_f32)
            # sizes that need the global-memory backward always took the global-memory forward, so `means` is valid
            ctx.large = K.augment_needs_large_path(x_f32.shape[2], x_f32.shape[3])
            ctx.save_for_backward(x_f32, params[:, n_u8_views:].contiguous(), means[n_u8_views:].contiguous())
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = None
        if ctx.has_f32 and ctx.needs_input_grad[2]:
            x_f32, params, means = ctx.saved_tensors
            dy = _c(dy[ctx.n_u8_views:])
            if ctx.large:
                dx = K.augment_simclr_large_bwd(x_f32, dy, params, ctx.order, means)
            else:
                dx = K.augment_simclr_bwd(x_f32, dy, params, ctx.order)
        return None, None, dx, None, None


class GaussianBlurFn(Function):
    """RandomApply(GaussianBlur) (augment/__init__.py:52-78,100-103): y = on ? blur(x) : x."""

    @staticmethod
    def forward(ctx, x, taps, on):
        ctx.save_for_backward(taps, on)
        return K.gaussian_blur(_c(x), taps, on)

    @staticmethod
    def backward(ctx, dy):
        taps, on = ctx.saved_tensors
        return K.gaussian_blur(_c(dy), taps, on, adjoint=True), None, None


class CutOutFn(Function):
    """RandomApply(CutOut) (augment/spatial.py:151-181): a masking, so the backward is the same masking."""

    @staticmethod
    def forward(ctx, x, params, length):
        ctx.save_for_backward(params)
        ctx.length = length
        return K.cutout(_c(x), params, length)

    @staticmethod
    def backward(ctx, dy):
        (params,) = ctx.saved_tensors
        return K.cutout(_c(dy), params, ctx.length), None, None


class ShiftFlipFn(Function):
    """HorizontalFlipRandomCrop / RandomCrop (augment/spatial.py:14-67): a per-sample index map, so the backward is the
    transposed map (cb200_shift_flip_bwd)."""

    @staticmethod
    def forward(ctx, x, params, padding_mode):
        ctx.save_for_backward(params)
        ctx.padding_mode = padding_mode
        return K.shift_flip(_c(x), params, padding_mode)

    @staticmethod
    def backward(ctx, dy):
        (params,) = ctx.saved_tensors
        return K.shift_flip(_c(dy), params, ctx.padding_mode, adjoint=True), None, None


class DiffAugFn(Function):
    """DiffAugment (third_party/diffaug.py:8-21) on explicit draws: affine in x, so the backward needs only the draws."""

    @staticmethod
    def forward(ctx, x, params, flags):
        ctx.save_for_backward(params)
        ctx.flags = flags
        return K.diffaug(_c(x), params, flags)

    @staticmethod
    def backward(ctx, dy):
        (params,) = ctx.saved_tensors
        return K.diffaug(_c(dy), params, ctx.flags, adjoint=True), None, None


class NoiseClampFn(Function):
    """Gaussian (augment/__init__.py:40-49): clamp(x + noise * sigma, 0, 1)."""

    @staticmethod
    def forward(ctx, x, noise, sigma):
        x = _c(x)
        ctx.save_for_backward(x, noise)
        ctx.sigma = float(sigma)
        return K.noise_clamp_fwd(x, noise, ctx.sigma)

    @staticmethod
    def backward(ctx, dy):
        x, noise = ctx.saved_tensors
        return K.noise_clamp_bwd(x, noise, _c(dy), ctx.sigma), None, None


# ------------------------------------------------------------------------------------------------
# spectral norm + packing of every D_SNDCGAN weight  (models/gan/sndcgan.py:111-118)
# ------------------------------------------------------------------------------------------------
class SNLayerSpec(object):
    """Static description of one spectrally-normalised layer of the discriminator."""

    def __init__(self, name, module, kind, ks=1, stride=1):
        self.name, self.module, self.kind, self.ks, self.stride = name, module, kind, ks, stride


class SNPackFn(Function):
    """(weight_orig_1 ... weight_orig_L) -> packed W/sigma matrices.

    outputs (differentiable): conv packs in layer order, then Wcat (the three first-layer heads stacked,
    columns in (h,w,c) order), then the three second-layer head packs.
    side (python dict returned through `holder`): data-gradient packs, sigmas."""

    @staticmethod
    def forward(ctx, holder, training, *weights):
        specs = holder["specs"]
        dev = weights[0].device
        weights = [_c(w) for w in weights]
        # strict precision applies to passes through FROZEN weights (the generator step): no weight gradient is asked for
        strict = holder["strict"] = (bool(holder.get("allow_strict")) and
                                     (precision.strict_full() or
                                      (precision.strict_enabled() and not any(ctx.needs_input_grad[2:]))))
        rnd = not strict
        side = {"dgrad": {}, "sigma": {}}
        sig_all = torch.empty(len(specs), 2, device=dev, dtype=torch.float32)
        sig = [sig_all[i] for i in range(len(specs))]
        layers = [(w, s.module.weight_u, s.module.weight_v, sg) for s, w, sg in zip(specs, weights, sig)]
        for i in range(0, len(layers), 16):                      # the batched kernels take up to 16 layers per launch
            K.sn_power_iter_batched(layers[i:i + 16], training=training)
        # u / v of THIS forward (the backward of sigma needs them; the module buffers move on at the next forward):
        # one flat snapshot filled by a single batched copy, and only when some weight will receive a gradient
        if any(ctx.needs_input_grad[2:]):
            srcs = [t for s in specs for t in (s.module.weight_u, s.module.weight_v)]
            flat = torch.empty(sum(t.numel() for t in srcs), device=dev, dtype=torch.float32)
            views, off = [], 0
            for t in srcs:
                views.append(flat[off:off + t.numel()])
                off += t.numel()
            torch._foreach_copy_(views, srcs)
            saved_uv = [(views[2 * i], views[2 * i + 1]) for i in range(len(specs))]
        else:
            saved_uv = None
        head1 = [s for s in specs if s.kind == "head1"]
        outs, jobs = [], []
        wcat = wcat_t = None
        # zero-padded packs (3 -> 32 input channels of the first layer's data-gradient pack, 1 -> 32 rows of `linear.l2`)
        # come out of ONE zero-filled allocation
        pad_sizes = []
        for s, w in zip(specs, weights):
            if s.kind == "conv_first":
                pad_sizes.append(32 * 9 * w.shape[0])
            elif s.kind == "head2" and w.shape[0] % 32:
                pad_sizes.append(2 * 32 * w.shape[1])
        pad_pool = torch.zeros(sum(pad_sizes), device=dev) if pad_sizes else None
        pad_off = [0]

        def padded(rows, cols):
            t = pad_pool[pad_off[0]:pad_off[0] + rows * cols].view(rows, cols)
            pad_off[0] += rows * cols
            return t

        for s, w, sigma in zip(specs, weights, sig):
            side["sigma"][s.name] = sigma
            cout = w.shape[0]
            if s.kind == "conv_first":
                fwd = torch.empty(cout, 27, device=dev)
                dg = padded(32, 9 * cout)
                jobs.append(dict(w4=w.view(cout, 27, 1, 1), sigma=sigma, fwd=fwd, ld_fwd=27, round_out=False))
                jobs.append(dict(w4=w, sigma=sigma, dgrad=dg, dgrad_mode=1, round_out=rnd))
                outs.append(fwd)
                side["dgrad"][s.name] = dg
            elif s.kind == "conv":
                cin = w.shape[1]
                fwd = torch.empty(cout, s.ks * s.ks * cin, device=dev)
                if s.stride == 1:
                    dg = torch.empty(cin, s.ks * s.ks * cout, device=dev)
                    jobs.append(dict(w4=w, sigma=sigma, fwd=fwd, ld_fwd=fwd.shape[1], dgrad=dg, dgrad_mode=1, round_out=rnd))
                else:
                    dg = torch.empty(4 * cin, 4 * cout, device=dev)
                    jobs.append(dict(w4=w, sigma=sigma, fwd=fwd, ld_fwd=fwd.shape[1], dgrad=dg, dgrad_mode=2, round_out=rnd))
                outs.append(fwd)
                side["dgrad"][s.name] = dg
            elif s.kind == "conv_plain":         # any kernel size / stride: forward GEMM matrix only (SNResNet18)
                cin = w.shape[1]
                fwd = torch.empty(cout, s.ks * s.ks * cin, device=dev)
                jobs.append(dict(w4=w, sigma=sigma, fwd=fwd, ld_fwd=fwd.shape[1], round_out=rnd))
                outs.append(fwd)
            elif s.kind == "head1":
                nfeat = w.shape[1]
                c_last, sh, sw = holder["feat_chw"]
                if wcat is None:
                    wcat = torch.empty(len(head1) * cout, nfeat, device=dev)
                    wcat_t = torch.empty(nfeat, len(head1) * cout, device=dev)
                idx = head1.index(s)
                jobs.append(dict(w4=w.view(cout, c_last, sh, sw), sigma=sigma, fwd=wcat[idx * cout:], ld_fwd=nfeat,
                                 dgrad=wcat_t, dgrad_mode=3, ldt=wcat_t.shape[1], col0=idx * cout, round_out=rnd))
                if idx == len(head1) - 1:
                    outs.append(wcat)
                    side["dgrad"]["wcat"] = wcat_t
            else:   # head2: [cout, hidden]; the 1-output `linear.l2` is padded to 32 rows
                hid = w.shape[1]
                rows = cout if cout % 32 == 0 else 32
                fwd = padded(rows, hid) if rows != cout else torch.empty(rows, hid, device=dev)
                dg = padded(hid, rows) if rows != cout else torch.empty(hid, rows, device=dev)
                jobs.append(dict(w4=w.view(cout, hid, 1, 1), sigma=sigma, fwd=fwd, ld_fwd=hid, dgrad=dg, dgrad_mode=3,
                                 ldt=rows, col0=0, round_out=rnd))
                outs.append(fwd)
                side["dgrad"][s.name] = dg
        for i in range(0, len(jobs), 16):
            K.sn_pack_batched(jobs[i:i + 16])
        if strict:
            # error-compensated weights (the packs above are unrounded fp32): the forward GEMMs read [w_hi | w_hi | w_lo]
            # per tap (3x the reduction length, side["fwd3"]), the data-gradient GEMMs [w_hi | w_lo] (2x).  The
            # differentiable outputs stay the ordinary packs - they carry the weight gradients back to SNPackFn.backward.
            # The head packs are split where they are used (HeadsFn).
            side["fwd3"] = {}
            k = 0
            for s, w in zip(specs, weights):
                if s.kind == "conv_first":
                    side["dgrad"][s.name] = _split_weight_bwd(side["dgrad"][s.name], 9)
                    k += 1
                elif s.kind in ("conv", "conv_plain"):
                    side["fwd3"][s.name] = _split_weight_fwd(outs[k], s.ks * s.ks)
                    if s.name in side["dgrad"]:
                        side["dgrad"][s.name] = _split_weight_bwd(side["dgrad"][s.name], 9 if s.stride == 1 else 4)
                    k += 1
        holder["side"] = side
        ctx.holder = holder
        ctx.saved_uv = saved_uv
        ctx.sig = sig
        ctx.save_for_backward(*weights)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dpacks):
        specs = ctx.holder["specs"]
        weights = ctx.saved_tensors
        grads = [None] * len(specs)
        head1 = [s for s in specs if s.kind == "head1"]
        head2 = [s for s in specs if s.kind == "head2"]
        n_conv = sum(1 for s in specs if s.kind in ("conv_first", "conv", "conv_plain"))
        jobs = []
        for li, (s, w) in enumerate(zip(specs, weights)):
            if not ctx.needs_input_grad[2 + li]:
                continue
            u, v = ctx.saved_uv[li]
            cout = w.shape[0]
            if s.kind == "conv_first":
                g, w4 = dpacks[li], w.view(cout, 27, 1, 1)
            elif s.kind in ("conv", "conv_plain"):
                g, w4 = dpacks[li], w
            elif s.kind == "head1":
                g = dpacks[n_conv]
                c_last, sh, sw = ctx.holder["feat_chw"]
                w4 = w.view(cout, c_last, sh, sw)
                if g is not None:
                    idx = head1.index(s)
                    g = _c(g)[idx * cout:(idx + 1) * cout]
            else:
                g, w4 = dpacks[n_conv + 1 + head2.index(s)], w.view(cout, w.shape[1], 1, 1)
            if g is None:
                continue
            g = g if g.stride(-1) == 1 and (g.dim() < 2 or g.stride(0) >= g.shape[1]) else _c(g)
            dw = torch.empty_like(w)
            grads[li] = dw
            jobs.append(dict(dw_hat_packed=g, ld_fwd=g.stride(0) if g.dim() == 2 else g.shape[-1], w4=w4, u=u, v=v,
                             sigma=ctx.sig[li], dw=dw))
        for i in range(0, len(jobs), 16):
            K.sn_weight_bwd_batched(jobs[i:i + 16])
        return (None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# discriminator backbone  (models/gan/sndcgan.py:91-109,122-128)
# ------------------------------------------------------------------------------------------------
class SNDCGANBackboneFn(Function):
    """x [B,3,H,W] (NCHW, in [0,1]) -> features [B, Hf*Wf*C] in (h, w, c) order (post-LeakyReLU).

    args after x: for each conv layer (pack, bias); `holder` carries the data-gradient packs."""

    SLOPE = 0.1

    @staticmethod
    def forward(ctx, holder, x, *wb):
        specs = [s for s in holder["specs"] if s.kind in ("conv_first", "conv")]
        x = _c(x)
        strict = ctx.strict = bool(holder.get("strict"))
        acts = [K.conv_first_fwd(x, wb[0].view(-1, 3, 3, 3), None, wb[1], slope=SNDCGANBackboneFn.SLOPE,
                                 round_out=not strict)]
        for li, s in enumerate(specs[1:], start=1):
            # strict: activations stay fp32 and enter the GEMM as hi | lo | hi against [w_hi | w_hi | w_lo]
            a_in = K.split_tf32(acts[-1], 0) if strict else acts[-1]
            wpack = holder["side"]["fwd3"][s.name] if strict else wb[2 * li]
            acts.append(K.conv2d_nhwc_fwd(a_in, wpack, wb[2 * li + 1], s.ks, s.stride,
                                          slope=SNDCGANBackboneFn.SLOPE, round_out=not strict))
        ctx.specs = specs
        ctx.dgrad = [holder["side"]["dgrad"][s.name] for s in specs]
        ctx.save_for_backward(x, *acts, *[wb[2 * i] for i in range(len(specs))])
        feat = acts[-1]
        return feat.view(feat.shape[0], -1)

    @staticmethod
    def backward(ctx, dfeat):
        specs = ctx.specs
        L = len(specs)
        saved = ctx.saved_tensors
        x, acts = saved[0], saved[1:1 + L]
        need_x = ctx.needs_input_grad[1]
        need_w = [ctx.needs_input_grad[2 + 2 * i] for i in range(L)]
        need_b = [ctx.needs_input_grad[3 + 2 * i] for i in range(L)]
        grads_w, grads_b = [None] * L, [None] * L
        # g = gradient w.r.t. the pre-activation of the last layer
        g = K.lrelu_bwd(_c(dfeat).view_as(acts[-1]), acts[-1], SNDCGANBackboneFn.SLOPE, round_out=True)
        if need_b[L - 1]:
            grads_b[L - 1] = K.colsum(g.view(-1, g.shape[-1]))
        for li in range(L - 1, 0, -1):
            s = specs[li]
            a_in = acts[li - 1]
            if need_w[li]:
                if ctx.strict:      # activation compensated along the (pixel) reduction axis: [a_hi ; a_lo] against [g ; g]
                    grads_w[li] = K.conv2d_nhwc_wgrad(K.split_tf32(a_in, 2).view((2 * a_in.shape[0],) + tuple(a_in.shape[1:])),
                                                      torch.cat([g, g], dim=0), s.ks, s.stride)
                else:
                    grads_w[li] = K.conv2d_nhwc_wgrad(a_in, g, s.ks, s.stride)
            if li > 1 or need_x or need_w[0] or need_b[0]:
                # the data gradient (x lrelu') IS dL/d(pre-activation) of layer li-1: its column sum, fused into the
                # GEMM epilogue, is that layer's bias gradient (layer 0 gets its own from conv_first_wgrad)
                db_prev = None
                if li > 1 and need_b[li - 1] and _FUSE_COLSUM:
                    db_prev = torch.empty(a_in.shape[-1], device=a_in.device, dtype=torch.float32)
                g = K.conv2d_nhwc_dgrad(K.split_tf32(g, 1) if ctx.strict else g, ctx.dgrad[li], tuple(a_in.shape), s.ks,
                                        s.stride, act_in=a_in, slope=SNDCGANBackboneFn.SLOPE, round_out=True, colsum=db_prev)
                if li > 1 and need_b[li - 1] and db_prev is None:
                    db_prev = K.colsum(g.view(-1, g.shape[-1]))
                grads_b[li - 1] = db_prev
            else:
                g = None
        dx = None
        if g is not None:
            if need_w[0] or need_b[0]:
                dw0, db0 = K.conv_first_wgrad(x, g)
                grads_w[0] = dw0 if need_w[0] else None
                grads_b[0] = db0 if need_b[0] else None
            if need_x:
                B, H, W, _ = g.shape
                dpad = K.conv2d_nhwc_dgrad(K.split_tf32(g, 1) if ctx.strict else g, ctx.dgrad[0], (B, H, W, 32), 3, 1)
                dx = K.conv_first_dgrad_finish(dpad)
        out = [None, dx]
        for i in range(L):
            out += [grads_w[i], grads_b[i]]
        return tuple(out)


class ConvFirstFn(Function):
    """`LeakyReLU(Conv2d(3, 64, 3, 1, 1)(x * 2 - 1))` on an NCHW image -> NHWC activation (first layer of D_SNDCGAN and
    D_SNResNet18, models/gan/snresnet.py:77-79) from the packed W/sigma [64, 27]; first order."""

    @staticmethod
    def forward(ctx, x, w27, bias, dgrad_pack, slope):
        x = _c(x)
        y = K.conv_first_fwd(x, w27.view(-1, 3, 3, 3), None, bias, slope=slope, round_out=True)
        ctx.save_for_backward(x, y)
        ctx.dgrad, ctx.slope = dgrad_pack, slope
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        g = K.lrelu_bwd(_c(dy), y, ctx.slope, round_out=True)
        dw = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = K.conv_first_wgrad(x, g)
        if ctx.needs_input_grad[0]:
            B, H, W, _ = g.shape
            dx = K.conv_first_dgrad_finish(K.conv2d_nhwc_dgrad(g, ctx.dgrad, (B, H, W, 32), 3, 1))
        return dx, dw, db, None, None


class ConvPackedFn(Function):
    """3x3 stride-1 NHWC convolution from the packed matrices written by the spectral-norm kernels (forward pack
    [Cout, 9*Cin], data-gradient pack [Cin, 9*Cout]); the weight gradient comes back in the forward-pack layout, which
    is what SNPackFn.backward consumes.  First order."""

    @staticmethod
    def forward(ctx, x, wpack, dgrad_pack):
        x = _c(x)
        ctx.save_for_backward(x)
        ctx.dgrad = dgrad_pack
        return K.conv2d_nhwc_fwd(x, wpack, None, 3, 1)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = K.round_tf32_(_c(dy))
        dx = K.conv2d_nhwc_dgrad(dy, ctx.dgrad, tuple(x.shape), 3, 1) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            cout = dy.shape[-1]
            if cout % 128:          # the wgrad kernel tiles Cout by 128: 64-channel layers present dY zero-extended
                from . import sg2_kernels as S
                dw = K.conv2d_nhwc_wgrad(x, S.pad_channels(dy, (cout + 127) // 128 * 128), 3, 1)[:cout]
            else:
                dw = K.conv2d_nhwc_wgrad(x, dy, 3, 1)
        return dx, dw, None


# ------------------------------------------------------------------------------------------------
# generator  (models/gan/sndcgan.py:13-52)
# ------------------------------------------------------------------------------------------------
def _bn_forward(K_, x2d, bn_state, gamma, beta, remap_s, training, sync, round_out=True):
    """Train-mode BN(+ReLU): returns (y, stats, count).  bn_state = (running_mean, running_var) or None."""
    M = x2d.shape[0]
    if training:
        sums = K.bn_stats(x2d)
        count = float(M)
        if sync:
            import torch.distributed as dist
            dist.all_reduce(sums)
            count *= dist.get_world_size()
        rm, rv = bn_state if bn_state is not None else (None, None)
        stats = K.bn_finalize(sums, count, rm, rv)
    else:
        rm, rv = bn_state
        stats = torch.stack([rm, torch.rsqrt(rv + 1e-5)])
        count = float(M)
    return K.bn_apply_relu(x2d, stats, gamma, beta, remap_s=remap_s, round_out=round_out), stats, count


def _bn_backward(dy2d, y2d, x2d, stats, gamma, count, remap_s, sync):
    """Returns (dx, dgamma, dbeta); under SyncBN the sums are all-reduced for dx, the affine grads stay local
    (DDP averages them afterwards, as torch's SyncBatchNorm does)."""
    sums = K.bn_bwd_reduce(dy2d, y2d, x2d, stats, remap_s=remap_s)
    if sync:
        import torch.distributed as dist
        local = sums.clone()
        dbeta, dgamma = local[0], local[1]
        dist.all_reduce(sums)
    else:
        dbeta, dgamma = sums[0], sums[1]
    dx = K.bn_bwd_apply(dy2d, y2d, x2d, stats, gamma, sums, count, remap_s=remap_s)
    return dx, dgamma, dbeta


class GSNDCGANFn(Function):
    """z [N,nz] -> images [N,3,8*s_h,8*s_w] in [0,1].

    Linear and the four ConvTranspose2d layers run on the tcgen05 tap-GEMM kernels (a transposed convolution
    is the data-gradient of the convolution with the same weight tensor); BatchNorm+ReLU, tanh are HBM-bound
    SIMT kernels; activations are NHWC.  args: z, linear.{weight,bias}, then per block (convT.weight,
    convT.bias) interleaved with (bn.weight, bn.bias): see G_SNDCGAN.forward."""

    @staticmethod
    def forward(ctx, holder, z, w_lin, b_lin, g0, be0, w1, b1, g1, be1, w2, b2, g2, be2, w3, b3, g3, be3, w4, b4):
        training, sync = holder["training"], holder["sync"]
        bn_states = holder["bn_states"]           # [(running_mean, running_var)] * 4
        sh, sw = holder["s_hb"], holder["s_wb"]
        dev = z.device
        N = z.shape[0]
        convs = [_c(w1), _c(w2), _c(w3)]
        w4 = _c(w4)
        w_lin = _c(w_lin)
        # strict precision (contrad_b200/precision.py): the generator step, i.e. whenever a generator weight wants a gradient
        strict = precision.strict_full() or (precision.strict_enabled() and any(ctx.needs_input_grad[2:]))
        if strict:
            # forward operands: activation hi | lo | hi against weights [w_hi | w_hi | w_lo] along the reduction axis (the
            # INPUT channels `o` of a transposed convolution = dim 0 of its weight); data-gradient packs [w_hi | w_lo]
            hi, lo = _hi_lo(w_lin)
            wl = torch.cat([hi, hi, lo], dim=1).contiguous()
            jobs, tpacks, fpacks = [], [], []
            for w in convs:
                o, i = w.shape[0], w.shape[1]
                hi, lo = _hi_lo(w)
                fp = torch.empty(o, 16 * 2 * i, device=dev)
                tp = torch.empty(4 * i, 4 * 3 * o, device=dev)
                jobs.append(dict(w4=torch.cat([hi, hi, lo], dim=0).contiguous(), dgrad=tp, dgrad_mode=2))
                jobs.append(dict(w4=torch.cat([hi, lo], dim=1).contiguous(), fwd=fp, ld_fwd=16 * 2 * i))
                tpacks.append(tp); fpacks.append(fp)
            hi, lo = _hi_lo(w4)
            t4 = torch.zeros(32, 9 * 3 * w4.shape[0], device=dev)
            jobs.append(dict(w4=torch.cat([hi, hi, lo], dim=0).contiguous(), dgrad=t4, dgrad_mode=1))
            K.sn_pack_batched(jobs)
            sp = lambda t: K.split_tf32(t, 0)
            zr = _c(z)
        else:
            # ---- packs (TF32-rounded): linear as-is; convT weights as OIHW of the underlying conv
            wl = torch.empty_like(w_lin)
            jobs = [dict(w4=w_lin.view(w_lin.shape[0], w_lin.shape[1], 1, 1), fwd=wl, ld_fwd=w_lin.shape[1])]
            tpacks, fpacks = [], []
            for w in convs:
                o, i = w.shape[0], w.shape[1]
                fp = torch.empty(o, 16 * i, device=dev)
                tp = torch.empty(4 * i, 4 * o, device=dev)
                jobs.append(dict(w4=w, fwd=fp, ld_fwd=16 * i, dgrad=tp, dgrad_mode=2))
                tpacks.append(tp); fpacks.append(fp)
            t4 = torch.zeros(32, 9 * w4.shape[0], device=dev)
            jobs.append(dict(w4=w4, dgrad=t4, dgrad_mode=1))
            K.sn_pack_batched(jobs)
            sp = lambda t: t
            zr = K.round_tf32_(z)
        rnd = not strict
        h0 = K.gemm_nt(sp(zr), wl, b_lin)
        a, st0, cnt0 = _bn_forward(K, h0, bn_states[0], g0, be0, sh * sw, training, sync, round_out=rnd)
        xs, acts, stats, counts = [h0], [a], [st0], [cnt0]
        hw = (sh, sw)
        for li, (w, b, g, be) in enumerate(((convs[0], b1, g1, be1), (convs[1], b2, g2, be2), (convs[2], b3, g3, be3))):
            o, i = w.shape[0], w.shape[1]
            x_in = acts[-1].view(N, hw[0], hw[1], o)
            hw = (hw[0] * 2, hw[1] * 2)
            x = K.conv2d_nhwc_dgrad(sp(x_in), tpacks[li], (N, hw[0], hw[1], i), 4, 2, bias_out=b, slope=1.0)
            y, st, cnt = _bn_forward(K, x.view(-1, i), bn_states[li + 1], g, be, 0, training, sync, round_out=rnd)
            xs.append(x); acts.append(y); stats.append(st); counts.append(cnt)
        c_last = convs[2].shape[1]
        pre = K.conv2d_nhwc_dgrad(sp(acts[-1].view(N, hw[0], hw[1], c_last)), t4, (N, hw[0], hw[1], 32), 3, 1)
        out = K.g_final_fwd(pre, b4)
        ctx.strict = strict
        ctx.meta = (sh, sw, sync, counts, [tuple(w.shape) for w in convs])
        ctx.w_lin = w_lin if ctx.needs_input_grad[1] else None      # only a latent that asks for a gradient needs it (below)
        ctx.save_for_backward(zr, w4, g0, g1, g2, g3, out, *xs, *acts, *stats, *fpacks)
        return out

    @staticmethod
    def backward(ctx, dout):
        sh, sw, sync, counts, wshapes = ctx.meta
        sv = ctx.saved_tensors
        zr, w4, gammas, out = sv[0], sv[1], sv[2:6], sv[6]
        xs, acts, stats, fpacks = sv[7:11], sv[11:15], sv[15:19], sv[19:22]
        N = zr.shape[0]
        grads = [None] * 20
        dpre, db4 = K.g_final_bwd(dout, out)
        H, W = out.shape[2], out.shape[3]
        c_last = wshapes[2][1]
        a3 = acts[3].view(N, H, W, c_last)
        dw4, _ = K.conv_first_wgrad(dpre, a3, in_scale=1.0, in_shift=0.0)
        grads[18], grads[19] = dw4.view_as(w4), db4
        da = K.conv_first_fwd(dpre, w4, None, None, slope=1.0, round_out=False, in_scale=1.0, in_shift=0.0)
        hw = (H, W)
        for li in (2, 1, 0):
            o, i = wshapes[li][0], wshapes[li][1]
            dx, dgam, dbet = _bn_backward(da.view(-1, i), acts[li + 1], xs[li + 1].view(-1, i), stats[li + 1],
                                          gammas[li + 1], counts[li + 1], 0, sync)
            base = 6 + 4 * li                       # positions of (w, b, gamma, beta) of block li in forward's args
            grads[base + 2], grads[base + 3] = dgam, dbet
            dx4 = dx.view(N, hw[0], hw[1], i)
            a_in = acts[li].view(N, hw[0] // 2, hw[1] // 2, o)
            if ctx.strict:      # activation compensated over the (pixel) reduction axis: [a_hi ; a_lo] against [dx ; dx]
                dwp = K.conv2d_nhwc_wgrad(torch.cat([dx4, dx4], dim=0),
                                          K.split_tf32(a_in, 2).view(2 * N, hw[0] // 2, hw[1] // 2, o), 4, 2)
            else:
                dwp = K.conv2d_nhwc_wgrad(dx4, a_in, 4, 2)                    # [o, 16*i] forward-pack layout
            dw = torch.empty(o, i, 4, 4, device=dx.device)
            K.sn_weight_bwd(dwp, dwp.shape[1], dw, None, None, None, dw)
            grads[base], grads[base + 1] = dw, K.colsum(dx)
            da = K.conv2d_nhwc_fwd(K.split_tf32(dx4, 1) if ctx.strict else dx4, fpacks[li], None, 4, 2, slope=1.0,
                                   round_out=False)
            hw = (hw[0] // 2, hw[1] // 2)
        dh0, dgam, dbet = _bn_backward(da.view(N, -1), acts[0], xs[0], stats[0], gammas[0], counts[0], sh * sw, sync)
        grads[4], grads[5] = dgam, dbet
        if ctx.strict:
            grads[2] = K.gemm_tn_wgrad(torch.cat([dh0, dh0], dim=0), K.split_tf32(zr, 2).view(2 * N, -1))
        else:
            grads[2] = K.gemm_tn_wgrad(dh0, zr)
        grads[3] = K.colsum(dh0)
        if ctx.needs_input_grad[1]:
            # dL/dz = dh0 @ W_lin (latent optimisation / projection into the latent space; the training loop never asks: z
            # is sampled without grad, models/gan/sndcgan.py:50-52)
            wt = ctx.w_lin.detach().t().contiguous()                                   # [nz, 8192]
            if ctx.strict:
                hi, lo = _hi_lo(wt)
                grads[1] = K.gemm_nt(K.split_tf32(dh0, 0), torch.cat([hi, hi, lo], dim=1).contiguous())
            else:
                grads[1] = K.gemm_nt(K.round_tf32_(dh0), K.round_tf32(wt))
        return tuple(grads)


# ------------------------------------------------------------------------------------------------
# heads: linear (TinyDiscriminator), projection, projection2  (models/gan/base.py:14-35,92-101,123-133)
# ------------------------------------------------------------------------------------------------
_UNIT_ROWS = {}


def _unit_row(width, ones, device):
    """[1, width] row with `ones` leading ones (cached per device): `g * row` zero-extends a [B, ones] gradient."""
    key = (width, ones, device)
    row = _UNIT_ROWS.get(key)
    if row is None:
        row = torch.zeros(1, width, device=device)
        row[:, :ones] = 1.0
        if not (device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            _UNIT_ROWS[key] = row          # (a tensor born inside a capture belongs to the graph's pool: not cached)
    return row


class HeadsFn(Function):
    """features [B,F] -> (d [B,1], projection [B,P], projection2 [B,P]).

    One GEMM computes the three hidden layers (A read once): H = lrelu(feat @ Wcat^T + bcat), Wcat [3*hid, F].
    sg_linear: the `linear` head does not propagate into the features (features.detach(), base.py:123-126)."""

    SLOPE = 0.1

    @staticmethod
    def forward(ctx, holder, sg_linear, feat, wcat, bcat, w_l2, b_l2, w_p1, b_p1, w_p2, b_p2):
        feat = _c(feat)
        hid = wcat.shape[0] // 3
        ctx.set_materialize_grads(False)          # heads the loss does not use arrive as None (G step: only `d`)
        strict = ctx.strict = bool(holder.get("strict"))
        if strict:
            # error-compensated operands: A = hi | lo | hi, B = [w_hi | w_hi | w_lo] along the reduction axis
            H = K.gemm_nt(K.split_tf32(feat, 0), _split_weight_fwd(wcat, 1), bcat, slope=HeadsFn.SLOPE, round_out=False)
            sp = lambda t: K.split_tf32(t, 0)
            d32 = K.gemm_nt(sp(H[:, :hid]), _split_weight_fwd(w_l2, 1), None)
            p1 = K.gemm_nt(sp(H[:, hid:2 * hid]), _split_weight_fwd(w_p1, 1), b_p1)
            p2 = K.gemm_nt(sp(H[:, 2 * hid:]), _split_weight_fwd(w_p2, 1), b_p2)
        else:
            H = K.gemm_nt(feat, wcat, bcat, slope=HeadsFn.SLOPE, round_out=True)
            d32 = K.gemm_nt(H[:, :hid], w_l2, None)                 # 1 output row padded to 32; bias added on the slice
            p1 = K.gemm_nt(H[:, hid:2 * hid], w_p1, b_p1)
            p2 = K.gemm_nt(H[:, 2 * hid:], w_p2, b_p2)
        side = holder["side"]["dgrad"]
        ctx.t_cat, ctx.t_l2, ctx.t_p1, ctx.t_p2 = side["wcat"], side["linear.l2"], side["projection.2"], side["projection2.2"]
        ctx.sg_linear, ctx.hid, ctx.n_out = sg_linear, hid, b_l2.numel()
        ctx.save_for_backward(feat, H)
        return d32[:, :ctx.n_out] + b_l2, p1, p2

    @staticmethod
    def backward(ctx, dd, dp1, dp2):
        feat, H = ctx.saved_tensors
        hid, B = ctx.hid, feat.shape[0]
        dev = feat.device
        ng = ctx.needs_input_grad
        dH = torch.empty_like(H)
        db_cat = torch.empty(3 * hid, device=dev, dtype=torch.float32) if (ng[4] and _FUSE_COLSUM) else None    # = colsum(dH)
        parts = ((dd, ctx.t_l2, 0, 32), (dp1, ctx.t_p1, 1, None), (dp2, ctx.t_p2, 2, None))
        used = [False, False, False]
        padded = [None, None, None]
        present = [i for i, (g, _, _, _) in enumerate(parts) if g is not None]
        need_full = ng[3] or ng[4]                 # first-layer weight / bias gradients read every column of dH
        for g, wt, idx, pad in parts:
            sl = slice(idx * hid, (idx + 1) * hid)
            if g is None:
                # columns of a head without gradient: zero them only where somebody reads them
                if need_full or (present and present[0] < idx < present[-1]):
                    dH[:, sl].zero_()
                if db_cat is not None:
                    db_cat[sl].zero_()
                continue
            used[idx] = True
            g = _c(g)
            if pad is not None and g.shape[1] != pad:
                # zero-extend [B, 1] to 128 columns (also reused by the wgrad below) with ONE broadcast multiply
                gp = g * _unit_row(128, g.shape[1], dev)
                padded[idx] = gp
                g = gp[:, :pad]
            else:
                padded[idx] = g
            if ctx.strict:            # weights compensated: A = g | g, B = [w_hi | w_lo]
                g, wt = K.split_tf32(g, 1), _split_weight_bwd(wt, 1)
            K.gemm_nt(g, wt, None, slope=HeadsFn.SLOPE, round_out=True, out=dH[:, sl], dact=H[:, sl],
                      colsum=None if db_cat is None else db_cat[sl])
        grads = [None] * 11
        # second-layer weight / bias gradients
        for (g, wt, idx, pad), wpos, bpos in zip(parts, (5, 7, 9), (6, 8, 10)):
            if g is None:
                continue
            sl = slice(idx * hid, (idx + 1) * hid)
            gfull = padded[idx]
            if ng[wpos]:
                if ctx.strict:
                    dw = K.gemm_tn_wgrad(torch.cat([gfull, gfull], dim=0), K.split_tf32(H[:, sl], 2).view(2 * B, hid))
                else:
                    dw = K.gemm_tn_wgrad(gfull if gfull.shape[1] % 128 == 0 else gfull, H[:, sl])
                rows = 32 if pad is not None else dw.shape[0]
                grads[wpos] = dw[:rows].contiguous() if rows != dw.shape[0] else dw
            if ng[bpos]:
                grads[bpos] = K.colsum(_c(g))
        if ng[3]:
            if ctx.strict:
                grads[3] = K.gemm_tn_wgrad(torch.cat([dH, dH], dim=0), K.split_tf32(feat, 2).view(2 * B, -1))
            else:
                grads[3] = K.gemm_tn_wgrad(dH, feat)
        if ng[4]:
            grads[4] = db_cat if db_cat is not None else K.colsum(dH)
        if ng[2]:
            # only heads that actually received a gradient contribute (dH of the others is zero): contract over the
            # smallest contiguous range of hidden columns that covers them (G step: the `linear` head alone, K = hid)
            live = [i for i in range(3) if used[i] and not (i == 0 and ctx.sg_linear)]
            if live:
                lo, hi = live[0] * hid, (live[-1] + 1) * hid
                if ctx.strict:
                    grads[2] = K.gemm_nt(K.split_tf32(dH[:, lo:hi], 1), _split_weight_bwd(ctx.t_cat[:, lo:hi], 1), None,
                                         slope=1.0, round_out=False)
                else:
                    grads[2] = K.gemm_nt(dH[:, lo:hi], ctx.t_cat[:, lo:hi], None, slope=1.0, round_out=False)
            else:
                grads[2] = torch.zeros_like(feat)
        return tuple(grads)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
class RowNormalizeFn(Function):
    """F.normalize(x, dim=1, eps=1e-12) (training/gan/contrad.py:43,48)."""

    @staticmethod
    def forward(ctx, x):
        y, inv = K.rownorm_fwd(x if x.stride(1) == 1 else x.contiguous())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return K.rownorm_bwd(_c(dy), y, inv)


class ContrastiveFn(Function):
    """mode 0: nt_xent on z = [out1; out2] (training/criterion.py:24-45);
    mode 1: supcon_fake on z = [out1; out2; others] (training/gan/contrad.py:8-32)."""

    @staticmethod
    def forward(ctx, z, n, mode, temperature):
        z = _c(z)
        loss, lse = K.contrastive_fwd(z, n, mode, temperature)
        ctx.save_for_backward(z, lse)
        ctx.cfg = (n, mode, temperature)
        return loss.view(())

    @staticmethod
    def backward(ctx, gout):
        z, lse = ctx.saved_tensors
        n, mode, temperature = ctx.cfg
        return K.contrastive_bwd(z, n, mode, temperature, lse, _c(gout).float()), None, None, None


class ContrastiveTCFn(Function):
    """The same two losses in north_star's tensor-core formulation: the similarity matrix of the loss rows against all
    rows is ONE tcgen05 GEMM (cb200_gemm_nt_tf32, operands error-compensated - the logits are sim / tau with tau = 0.1),
    softmax / cross-entropy are warp-shuffle row reductions over it (cb200_sim_rows_fwd / _bwd), and the embedding
    gradient is two more GEMMs: dZ_A += G Z, dZ += G^T Z_A.  Rows are zero-padded to the GEMM tile (128); padding
    columns are masked by the row kernels.  Memory: Ra x R fp32 twice (S and G): 9 MB at N = 512, 0.5 GB at N = 4096."""

    @staticmethod
    def forward(ctx, z, n, mode, temperature):
        z = _c(z)
        R, d = z.shape
        row0 = 2 * n if mode else 0
        Ra = R - row0
        Rp, Rap = (R + 127) // 128 * 128, (Ra + 127) // 128 * 128
        zp = z if Rp == R else torch.cat([z, z.new_zeros(Rp - R, d)])
        za = zp[row0:row0 + Rap] if row0 + Rap <= Rp and Rap == Ra else torch.cat([z[row0:], z.new_zeros(Rap - Ra, d)])
        hi, lo = _hi_lo(zp)
        S = K.gemm_nt(K.split_tf32(za, 0), torch.cat([hi, hi, lo], dim=1).contiguous())           # [Rap, Rp]
        lse, rowloss = K.sim_rows_fwd(S, Ra, R, n, mode, row0, temperature)
        ctx.save_for_backward(zp, za, S, lse)
        ctx.cfg = (n, mode, temperature, R, Ra, row0)
        return rowloss.sum()

    @staticmethod
    def backward(ctx, gout):
        zp, za, S, lse = ctx.saved_tensors
        n, mode, temperature, R, Ra, row0 = ctx.cfg
        Rap, Rp = S.shape
        G = K.sim_rows_bwd(S, Ra, R, n, mode, row0, temperature, lse, _c(gout).float(), Rap, Rp)      # [Rap, Rp]
        hi, lo = _hi_lo(zp.t().contiguous())
        t1 = K.gemm_nt(K.split_tf32(G, 0), torch.cat([hi, hi, lo], dim=1).contiguous())                # G Z      [Rap, d]
        gs, zs = K.split_tf32(G, 2), K.split_tf32(za, 2)
        t2 = K.gemm_tn_wgrad(torch.cat([gs[0], gs[1], gs[0]]), torch.cat([zs[0], zs[0], zs[1]]))       # G^T Z_A  [Rp, d]
        dz = t2[:R].clone() if Rp != R else t2
        dz[row0:row0 + Ra] += t1[:Ra]
        return dz, None, None, None


_CONTRASTIVE_PATH = os.environ.get("CB200_CONTRASTIVE", "auto").strip().lower()      # simt | tc | auto


def contrastive_loss(z, n, mode, temperature):
    """NT-Xent (mode 0) / supcon-fake (mode 1) of the stacked embeddings z.  Two implementations of the same function:
    the fused fp32 SIMT kernels (no similarity matrix in memory; fastest at the benchmark's R = 1-1.5 k rows, where the
    tensor-core path is ~15 launches against 2) and the tensor-core formulation (ContrastiveTCFn; its GEMMs win once the
    R^2 x 128 FLOPs dominate).  CB200_CONTRASTIVE = simt | tc | auto (default: tc from R >= 4096)."""
    use_tc = _CONTRASTIVE_PATH == "tc" or (_CONTRASTIVE_PATH == "auto" and z.shape[0] >= 4096)
    return (ContrastiveTCFn if use_tc else ContrastiveFn).apply(z, n, mode, temperature)


class GanDLossFn(Function):
    """d_all [3N,1] -> (L_dis, mean d_real, mean d_gen) with d_real = d_all[:N], d_gen = d_all[2N:]
    (training/gan/contrad.py:52-70).  The two means are logging values (non-differentiable here)."""

    @staticmethod
    def forward(ctx, d_all, n, kind, gen_offset=None):
        d_all = _c(d_all)
        flat = d_all.view(-1)
        off = 2 * n if gen_offset is None else int(gen_offset)      # [N real | N real view 2 | N fake] by default
        # d L_dis / d d_all as ONE vector: the kernel fills the real and the fake rows, the rest stays zero
        g_all = torch.zeros(d_all.shape[0], device=d_all.device)
        out = K.gan_d_loss(flat[:n], flat[off:off + n], kind, g_real=g_all[:n], g_gen=g_all[off:off + n])[0]
        ctx.save_for_backward(g_all)
        ctx.shape = tuple(d_all.shape)
        loss, means = out[0], out[1:]              # views of a fresh buffer (no copies)
        ctx.mark_non_differentiable(means)
        return loss, means

    @staticmethod
    def backward(ctx, gl, _gm):
        (g_all,) = ctx.saved_tensors
        return (g_all * gl).view(ctx.shape), None, None, None


class GanGLossFn(Function):
    """training/gan/contrad.py:75-80."""

    @staticmethod
    def forward(ctx, d_gen, kind):
        d_gen = _c(d_gen)
        out, g = K.gan_g_loss(d_gen.view(-1), kind)
        ctx.save_for_backward(g)
        ctx.shape = d_gen.shape
        return out.view(())

    @staticmethod
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return (g * gl).view(ctx.shape), None
