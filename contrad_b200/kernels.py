"""Raw (non-autograd) Python bindings of the C ABI: shape checks, output allocation, stream plumbing.
The autograd Functions in ``contrad_b200.functional`` are built on these."""
import ctypes

import torch

from . import _capi
from ._capi import check, f32, i32, i64, lib, ptr, stream_ptr

PARAM_FIELDS = ("sx", "sy", "bx", "by", "flip", "cj_on", "contrast", "hue", "sat", "val", "gray_on")


_PROFILE = None


def set_profile_hook(records):
    """records: a list that receives (kernel family, start event, end event, algorithmic FLOPs, bytes) for
    every C-ABI call made while set (bench.py's instrumented steps); None disables."""
    global _PROFILE
    _PROFILE = records


def _call(name, flops, nbytes, fn, *args):
    if _PROFILE is None:
        check(fn(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    _PROFILE.append((name, e0, e1, float(flops), float(nbytes)))


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------ augmentation
def augment_simclr_fwd(x, params, order):
    x = _f32c(x, "x")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    # order -1: per-image order in an extra row 11 (CUDA-graph replay, include/contrad_b200.h)
    assert C == 3 and params.shape == (len(PARAM_FIELDS) + (1 if order < 0 else 0), B), (x.shape, params.shape, order)
    y = torch.empty_like(x)
    _call("augment_simclr_fwd", 0, 8 * x.numel(), lib().cb200_augment_simclr_fwd, ptr(x), ptr(y), ptr(params), i32(B), i32(H), i32(W), i32(order),
                                         stream_ptr())
    return y


def augment_simclr_params(boxes, u, order_src, rows, cfg11):
    """[rows, B] parameter block of the fused chain from the host crop draws `boxes` [4,B], device uniforms `u` [7,B]
    and (rows == 12) the device scalar `order_src`; cfg11: 11 python floats (see include/contrad_b200.h)."""
    B = boxes.shape[1]
    assert boxes.shape == (4, B) and u.shape == (7, B) and boxes.is_contiguous() and u.is_contiguous()
    params = torch.empty(rows, B, device=u.device, dtype=torch.float32)
    cfg = (ctypes.c_float * 11)(*[float(v) for v in cfg11])
    _call("augment_simclr_params", 0, 0, lib().cb200_augment_simclr_params, ptr(boxes), ptr(u), ptr(order_src), ptr(params),
          i32(B), i32(rows), cfg, stream_ptr())
    return params


def augment_simclr_bwd(x, dy, params, order):
    x = _f32c(x, "x")
    dy = _f32c(dy, "dy")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    _call("augment_simclr_bwd", 0, 12 * x.numel(), lib().cb200_augment_simclr_bwd, ptr(x), ptr(dy), ptr(dx), ptr(params), i32(B), i32(H), i32(W), i32(order),
                                         stream_ptr())
    return dx


def augment_needs_large_path(H, W):
    """The one-image-per-CTA kernels hold H*W <= 4096 pixels in shared memory; beyond that the global-memory path."""
    return H * W > 4096 or W % 4 != 0


def augment_simclr_large_fwd(x, params, order):
    """Any image size.  Returns (y, means[B,3]); `means` must be passed to augment_simclr_large_bwd."""
    x = _f32c(x, "x")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    assert C == 3 and params.shape == (len(PARAM_FIELDS) + (1 if order < 0 else 0), B), (x.shape, params.shape, order)
    y = torch.empty_like(x)
    means = torch.empty(B, 3, device=x.device, dtype=torch.float32)
    _call("augment_simclr_fwd", 0, 12 * x.numel(), lib().cb200_augment_simclr_large_fwd, ptr(x), ptr(y), ptr(params), ptr(means),
          i32(B), i32(H), i32(W), i32(order), stream_ptr())
    return y, means


def augment_simclr_large_bwd(x, dy, params, order, means):
    x = _f32c(x, "x")
    dy = _f32c(dy, "dy")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    gsums = torch.empty(B, 3, device=x.device, dtype=torch.float32)
    _call("augment_simclr_bwd", 0, 20 * x.numel(), lib().cb200_augment_simclr_large_bwd, ptr(x), ptr(dy), ptr(dx), ptr(params),
          ptr(means), ptr(gsums), i32(B), i32(H), i32(W), i32(order), stream_ptr())
    return dx


def augment_simclr_mixed_fwd(x_u8, n_u8_views, x_f32, params, order):
    """Row f3: B = n_u8_views + len(x_f32) views in ONE launch; view b < n_u8_views reads uint8 image b % len(x_u8)
    (ToTensor's / 255 inside the kernel), the others the fp32 images.  Returns (y [B,3,H,W], means [B,3])."""
    params = _f32c(params, "params")
    if x_u8 is not None:
        if x_u8.dtype != torch.uint8:
            raise TypeError("x_u8 must be uint8 (got %s)" % x_u8.dtype)
        x_u8 = x_u8 if x_u8.is_contiguous() else x_u8.contiguous()
    if x_f32 is not None:
        x_f32 = _f32c(x_f32, "x_f32")
    ref = x_u8 if x_u8 is not None else x_f32
    n_u8 = 0 if x_u8 is None else x_u8.shape[0]
    n_f32 = 0 if x_f32 is None else x_f32.shape[0]
    _, C, H, W = ref.shape
    if n_u8_views and not n_u8:
        raise ValueError("n_u8_views = %d without uint8 images" % n_u8_views)
    if x_u8 is not None and x_f32 is not None and tuple(x_u8.shape[1:]) != tuple(x_f32.shape[1:]):
        raise ValueError("uint8 and fp32 images differ in shape: %s vs %s" % (tuple(x_u8.shape), tuple(x_f32.shape)))
    B = n_u8_views + n_f32
    assert C == 3 and params.shape == (len(PARAM_FIELDS) + (1 if order < 0 else 0), B), (ref.shape, params.shape, order)
    y = torch.empty(B, 3, H, W, device=ref.device, dtype=torch.float32)
    means = torch.empty(B, 3, device=ref.device, dtype=torch.float32)
    nbytes = n_u8_views * 3 * H * W * 5 + n_f32 * 3 * H * W * 8
    _call("augment_simclr_fwd", 0, nbytes, lib().cb200_augment_simclr_mixed_fwd, ptr(x_u8), i32(n_u8), i32(n_u8_views),
          ptr(x_f32), ptr(y), ptr(params), ptr(means), i32(B), i32(H), i32(W), i32(order), stream_ptr())
    return y, means


def gaussian_blur(x, taps, on, adjoint=False):
    """x [B,C,H,W]; taps [k] normalised 1-D Gaussian (device); on [B] 0/1 mask.  y = on ? blur(x) : x (or its adjoint)."""
    x = _f32c(x, "x")
    taps = _f32c(taps, "taps")
    on = _f32c(on, "on")
    B, C, H, W = x.shape
    assert on.numel() == B
    tmp = torch.empty_like(x)
    y = torch.empty_like(x)
    _call("gaussian_blur", 0, 16 * x.numel(), lib().cb200_gaussian_blur, ptr(x), ptr(tmp), ptr(y), ptr(taps), ptr(on), i32(B), i32(C),
          i32(H), i32(W), i32(taps.numel()), i32(1 if adjoint else 0), stream_ptr())
    return y


def cutout(x, params, length):
    """x [B,C,H,W]; params [3,B] = {on, h centre, w centre}; zeroes the clipped length x length square."""
    x = _f32c(x, "x")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    assert params.shape == (3, B)
    y = torch.empty_like(x)
    _call("cutout", 0, 8 * x.numel(), lib().cb200_cutout, ptr(x), ptr(y), ptr(params), i32(B), i32(C), i32(H), i32(W), i32(length),
          stream_ptr())
    return y


PADDING_MODES = {"zeros": 0, "border": 1, "reflection": 2}


def shift_flip(x, params, padding_mode, adjoint=False):
    """x [B,P,H,W]; params [3,B] = {sign, bias_x, bias_y}; nearest-neighbour mirror + translation (row f4).
    adjoint=True applies the transpose (x = dy, returns dx)."""
    x = _f32c(x, "x")
    params = _f32c(params, "params")
    B, P, H, W = x.shape
    assert params.shape == (3, B), (params.shape, B)
    mode = PADDING_MODES[padding_mode]
    y = torch.empty_like(x)
    fn = lib().cb200_shift_flip_bwd if adjoint else lib().cb200_shift_flip_fwd
    _call("shift_flip", 0, 8 * x.numel(), fn, ptr(x), ptr(y), ptr(params), i32(B), i32(P), i32(H), i32(W), i32(mode),
          stream_ptr())
    return y


def noise_clamp_fwd(x, noise, sigma):
    x = _f32c(x, "x")
    noise = _f32c(noise, "noise")
    assert x.shape == noise.shape
    y = torch.empty_like(x)
    _call("noise_clamp", 0, 12 * x.numel(), lib().cb200_noise_clamp_fwd, ptr(x), ptr(noise), ptr(y), f32(sigma), i64(x.numel()),
          stream_ptr())
    return y


def noise_clamp_bwd(x, noise, dy, sigma):
    x = _f32c(x, "x")
    noise = _f32c(noise, "noise")
    dy = _f32c(dy, "dy")
    dx = torch.empty_like(x)
    _call("noise_clamp", 0, 16 * x.numel(), lib().cb200_noise_clamp_bwd, ptr(x), ptr(noise), ptr(dy), ptr(dx), f32(sigma),
          i64(x.numel()), stream_ptr())
    return dx


DIFFAUG_FLAGS = {"color": 1, "translation": 2, "cutout": 4}


def diffaug(x, params, flags, adjoint=False):
    """DiffAugment on explicit draws: x [B,3,H,W], params [7,B] (include/contrad_b200.h).  adjoint=True: x = dy -> dx."""
    x = _f32c(x, "x")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    assert C == 3 and params.shape == (7, B), (x.shape, params.shape)
    y = torch.empty_like(x)
    scratch = torch.empty(B, device=x.device, dtype=torch.float32)
    fn = lib().cb200_diffaug_bwd if adjoint else lib().cb200_diffaug_fwd
    _call("diffaug", 0, 12 * x.numel(), fn, ptr(x), ptr(y), ptr(params), ptr(scratch), i32(B), i32(H), i32(W), i32(flags),
          stream_ptr())
    return y


# ------------------------------------------------------------------ tensor-core GEMM / conv
def _colsum_buf(colsum, n):
    if colsum is not None:
        assert colsum.is_cuda and colsum.dtype == torch.float32 and colsum.is_contiguous() and colsum.numel() == n
    return colsum


_SPLITK_COUNTERS = {}
_SPLITK_WS_BYTES = 20 << 20            # 296 CTAs x 128 x 128 fp32 partials
_SPLITK_MAX_UNITS = 256


def _offer_splitk_workspace(rows, n, classes, device):
    """Hand the next tap-GEMM call a split-K scratch when its tile list (128 x {128,64,32} tiles) cannot fill the GPU
    (include/contrad_b200.h: cb200_tapgemm_workspace).  Returns the scratch tensor (keep it alive across the call)."""
    if (rows + 127) // 128 >= 64:          # long tile lists take the persistent CTA-pair kernels: nothing to split
        return None
    key = (device.index, _capi.stream_ptr().value)
    counters = _SPLITK_COUNTERS.get(key)
    if counters is None:
        if torch.cuda.is_current_stream_capturing():
            return None                  # never allocate persistent state inside a capture: run unsplit
        counters = _SPLITK_COUNTERS[key] = torch.zeros(_SPLITK_MAX_UNITS, dtype=torch.int32, device=device)
    ws = torch.empty(_SPLITK_WS_BYTES // 4, dtype=torch.float32, device=device)
    check(lib().cb200_tapgemm_workspace(ptr(ws), i64(_SPLITK_WS_BYTES), ptr(counters), i32(_SPLITK_MAX_UNITS)),
          "tapgemm_workspace")
    return ws


def gemm_nt(a, bw, bias=None, slope=1.0, round_out=False, out=None, dact=None, colsum=None):
    """out[M,N] = lrelu_slope(a[M,K] @ bw[N,K]^T + bias), or (...) * lrelu'(dact) when dact is given.
    `a`, `bw`, `out`, `dact` may be row-strided 2-D views (dact must share out's row stride).
    colsum: optional contiguous [N] tensor that receives the column sums of `out` (fused into the epilogue)."""
    assert a.dim() == 2 and bw.dim() == 2 and a.shape[1] == bw.shape[1]
    assert a.stride(1) == 1 and bw.stride(1) == 1
    M, K = a.shape
    N = bw.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    assert out.shape == (M, N) and out.stride(1) == 1
    if dact is not None:
        assert dact.shape == (M, N) and dact.stride(1) == 1 and dact.stride(0) == out.stride(0)
    ws = _offer_splitk_workspace(M, N, 1, a.device)      # noqa: F841 - kept alive until the launch is enqueued
    _call("gemm_nt_tf32", 2.0 * M * N * K, 4 * (M * K + N * K + M * N), lib().cb200_gemm_nt_tf32, ptr(a), i64(a.stride(0)), ptr(bw), i64(bw.stride(0)), ptr(bias), ptr(dact), ptr(out),
                                   i64(out.stride(0)), i32(M), i32(N), i32(K), f32(slope), i32(1 if round_out else 0),
                                   ptr(_colsum_buf(colsum, N)), stream_ptr())
    return out


def conv2d_nhwc_fwd(x, wmat, bias, ks, stride, slope=1.0, round_out=False):
    """x [B,H,W,Cin] NHWC -> y [B,Ho,Wo,Cout]; wmat [Cout, ks*ks*Cin]."""
    x = _f32c(x, "x")
    B, H, W, Cin = x.shape
    Cout = wmat.shape[0]
    assert wmat.shape[1] == ks * ks * Cin and wmat.is_contiguous()
    Ho, Wo = H // stride, W // stride
    y = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.float32)
    ws = _offer_splitk_workspace(B * Ho * Wo, Cout, 1, x.device)      # noqa: F841
    _call("conv2d_nhwc_fwd", 2.0 * B * Ho * Wo * Cout * Cin * ks * ks, 4 * (x.numel() + wmat.numel() + y.numel()), lib().cb200_conv2d_nhwc_fwd, ptr(x), ptr(wmat), ptr(bias), ptr(y), i32(B), i32(H), i32(W), i32(Cin),
                                      i32(Cout), i32(ks), i32(stride), f32(slope), i32(1 if round_out else 0),
                                      stream_ptr())
    return y


def conv2d_nhwc_dgrad(dy, wmat_t, in_shape, ks, stride, act_in=None, bias_out=None, slope=1.0, round_out=False,
                      colsum=None):
    """dy [B,Ho,Wo,Cout] -> dx [B,H,W,Cin] (in_shape).  wmat_t: see pack_dgrad_weight.
    colsum: optional contiguous [Cin] tensor that receives sum over pixels of dx (= the bias gradient of the layer
    that produced the activation `act_in`)."""
    dy = _f32c(dy, "dy")
    B, H, W, Cin = in_shape
    Cout = dy.shape[3]
    dx = torch.empty(B, H, W, Cin, device=dy.device, dtype=torch.float32)
    ws = _offer_splitk_workspace(B * (H // stride) * (W // stride), Cin, stride * stride, dy.device)      # noqa: F841
    _call("conv2d_nhwc_dgrad", 2.0 * B * (H // stride) * (W // stride) * Cout * Cin * ks * ks, 4 * (dy.numel() + wmat_t.numel() + dx.numel()), lib().cb200_conv2d_nhwc_dgrad, ptr(dy), ptr(wmat_t), ptr(act_in), ptr(bias_out), ptr(dx), i32(B), i32(H),
                                        i32(W), i32(Cin), i32(Cout), i32(ks), i32(stride), f32(slope),
                                        i32(1 if round_out else 0), ptr(_colsum_buf(colsum, Cin)), stream_ptr())
    return dx


def conv2d_nhwc_wgrad(x, dy, ks, stride):
    """x [B,H,W,Cin], dy [B,Ho,Wo,Cout] -> dW_hat [Cout, ks*ks*Cin] (forward-pack layout)."""
    x = _f32c(x, "x")
    dy = _f32c(dy, "dy")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    dw = torch.empty(Cout, ks * ks * Cin, device=x.device, dtype=torch.float32)
    _call("conv2d_nhwc_wgrad", 2.0 * B * (H // stride) * (W // stride) * Cout * Cin * ks * ks, 4 * (x.numel() + dy.numel() + dw.numel()), lib().cb200_conv2d_nhwc_wgrad, ptr(x), ptr(dy), ptr(dw), i32(B), i32(H), i32(W), i32(Cin), i32(Cout),
                                        i32(ks), i32(stride), stream_ptr())
    return dw


def gemm_tn_wgrad(dy, x, out=None):
    """dW[N,K] = dy[M,N]^T @ x[M,K]; dy / x / out may be row-strided 2-D views."""
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0]
    assert dy.stride(1) == 1 and x.stride(1) == 1
    M, N = dy.shape
    Kd = x.shape[1]
    if out is None:
        out = torch.empty(N, Kd, device=x.device, dtype=torch.float32)
    assert out.shape == (N, Kd) and out.stride(1) == 1
    _call("gemm_tn_wgrad", 2.0 * M * N * Kd, 4 * (M * N + M * Kd + N * Kd), lib().cb200_gemm_tn_wgrad, ptr(dy), i64(dy.stride(0)), ptr(x), i64(x.stride(0)), ptr(out), i64(out.stride(0)),
                                    i32(M), i32(N), i32(Kd), stream_ptr())
    return out


# ------------------------------------------------------------------ spectral norm / packing
def sn_power_iter(w, u, v, sigma, training=True, eps=1e-12):
    """In place: u, v <- one power iteration (training) ; sigma[2] <- {sigma, 1/sigma}."""
    Cout = w.shape[0]
    Fdim = w.numel() // Cout
    t = torch.empty(Fdim, device=w.device, dtype=torch.float32)
    s = torch.empty(Cout, device=w.device, dtype=torch.float32)
    _call("sn_power_iter", 0, 0, lib().cb200_sn_power_iter, ptr(w), ptr(u), ptr(v), ptr(sigma), ptr(t), ptr(s), i32(Cout), i32(Fdim),
                                    f32(eps), i32(1 if training else 0), stream_ptr())


def sn_pack_weights(w4, sigma, fwd=None, ld_fwd=0, dgrad=None, dgrad_mode=0, ldt=0, col0=0, round_out=True):
    """w4: weight viewed as [Cout, Cin, KH, KW] (contiguous)."""
    Cout, Cin, KH, KW = w4.shape
    _call("sn_pack_weights", 0, 0, lib().cb200_sn_pack_weights, ptr(w4), ptr(sigma), ptr(fwd), i64(ld_fwd), ptr(dgrad), i32(dgrad_mode),
                                      i64(ldt), i32(col0), i32(Cout), i32(Cin), i32(KH), i32(KW),
                                      i32(1 if round_out else 0), stream_ptr())


def sn_weight_bwd(dw_hat_packed, ld_fwd, w4, u, v, sigma, dw_out, accumulate=False):
    Cout, Cin, KH, KW = w4.shape
    acc = torch.empty(1, device=w4.device, dtype=torch.float32)
    _call("sn_weight_bwd", 0, 0, lib().cb200_sn_weight_bwd, ptr(dw_hat_packed), i64(ld_fwd), ptr(w4), ptr(u), ptr(v), ptr(sigma), ptr(acc),
                                    ptr(dw_out), i32(1 if accumulate else 0), i32(Cout), i32(Cin), i32(KH), i32(KW),
                                    stream_ptr())
    return dw_out


class _SnLayer(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("u", ctypes.c_void_p), ("v", ctypes.c_void_p), ("sigma", ctypes.c_void_p),
                ("t", ctypes.c_void_p), ("s", ctypes.c_void_p), ("cout", ctypes.c_int), ("f", ctypes.c_int)]


class _SnPackJob(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("sigma", ctypes.c_void_p), ("fwd", ctypes.c_void_p), ("dgrad", ctypes.c_void_p),
                ("ld_fwd", ctypes.c_longlong), ("ldt", ctypes.c_longlong), ("cout", ctypes.c_int), ("cin", ctypes.c_int),
                ("kh", ctypes.c_int), ("kw", ctypes.c_int), ("dgrad_mode", ctypes.c_int), ("col0", ctypes.c_int),
                ("round_out", ctypes.c_int)]


class _SnBwdJob(ctypes.Structure):
    _fields_ = [("dw_hat_packed", ctypes.c_void_p), ("w", ctypes.c_void_p), ("u", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("sigma", ctypes.c_void_p), ("acc", ctypes.c_void_p), ("dw", ctypes.c_void_p),
                ("ld_fwd", ctypes.c_longlong), ("cout", ctypes.c_int), ("cin", ctypes.c_int), ("kh", ctypes.c_int),
                ("kw", ctypes.c_int)]


def _p(t):
    """Raw device address for the ctypes structs of the batched launches (ptr() refuses CPU tensors; None -> 0)."""
    return ptr(t).value or 0


def sn_power_iter_batched(layers, training=True, eps=1e-12):
    """layers: list of (w, u, v, sigma[2]).  One launch per phase for all of them (u, v, sigma updated in place)."""
    dev = layers[0][0].device
    fs = [w.numel() // w.shape[0] for w, _, _, _ in layers]
    couts = [w.shape[0] for w, _, _, _ in layers]
    pad4 = lambda n: (n + 3) // 4 * 4                      # keep every layer's scratch 16-byte aligned (float4 loads)
    t_all = torch.zeros(sum(pad4(f) for f in fs), device=dev, dtype=torch.float32) if training else torch.empty(1, device=dev)
    s_all = torch.empty(sum(couts), device=dev, dtype=torch.float32)
    arr = (_SnLayer * len(layers))()
    to, so = 0, 0
    for i, (w, u, v, sigma) in enumerate(layers):
        assert w.is_contiguous()
        arr[i] = _SnLayer(w.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(),
                          t_all.data_ptr() + 4 * to if training else 0, s_all.data_ptr() + 4 * so, couts[i], fs[i])
        to += pad4(fs[i])
        so += couts[i]
    _call("sn_power_iter", 0, 0, lib().cb200_sn_power_iter_batched, arr, i32(len(layers)), f32(eps),
          i32(1 if training else 0), stream_ptr())


def sn_pack_batched(jobs):
    """jobs: list of dicts(w4, sigma, fwd, ld_fwd, dgrad, dgrad_mode, ldt, col0, round_out)."""
    arr = (_SnPackJob * len(jobs))()
    for i, j in enumerate(jobs):
        cout, cin, kh, kw = j["w4"].shape
        arr[i] = _SnPackJob(j["w4"].data_ptr(), _p(j.get("sigma")), _p(j.get("fwd")), _p(j.get("dgrad")),
                            int(j.get("ld_fwd", 0)), int(j.get("ldt", 0)), cout, cin, kh, kw, int(j.get("dgrad_mode", 0)),
                            int(j.get("col0", 0)), 1 if j.get("round_out", True) else 0)
    _call("sn_pack_weights", 0, 0, lib().cb200_sn_pack_batched, arr, i32(len(jobs)), stream_ptr())


def sn_weight_bwd_batched(jobs):
    """jobs: list of dicts(dw_hat_packed, ld_fwd, w4, u, v, sigma, dw)."""
    dev = jobs[0]["w4"].device
    acc = torch.zeros(len(jobs), device=dev, dtype=torch.float32)
    arr = (_SnBwdJob * len(jobs))()
    for i, j in enumerate(jobs):
        cout, cin, kh, kw = j["w4"].shape
        arr[i] = _SnBwdJob(j["dw_hat_packed"].data_ptr(), j["w4"].data_ptr(), j["u"].data_ptr(), j["v"].data_ptr(),
                           j["sigma"].data_ptr(), acc.data_ptr() + 4 * i, j["dw"].data_ptr(), int(j["ld_fwd"]), cout, cin,
                           kh, kw)
    _call("sn_weight_bwd", 0, 0, lib().cb200_sn_weight_bwd_batched, arr, i32(len(jobs)), stream_ptr())


# ------------------------------------------------------------------ first conv layer
def conv_first_fwd(x, w, sigma, bias, slope=0.1, round_out=True, in_scale=2.0, in_shift=-1.0):
    x = _f32c(x, "x")
    B, C, H, W = x.shape
    assert C == 3 and w.shape == (64, 3, 3, 3) and w.is_contiguous()
    y = torch.empty(B, H, W, 64, device=x.device, dtype=torch.float32)
    _call("conv_first_fwd", 0, 0, lib().cb200_conv_first_fwd, ptr(x), ptr(w), ptr(sigma), ptr(bias), ptr(y), i32(B), i32(H), i32(W),
                                     f32(slope), i32(1 if round_out else 0), f32(in_scale), f32(in_shift), stream_ptr())
    return y


def conv_first_wgrad(x, dy, in_scale=2.0, in_shift=-1.0):
    x = _f32c(x, "x")
    dy = _f32c(dy, "dy")
    B, C, H, W = x.shape
    dw = torch.zeros(64, 27, device=x.device, dtype=torch.float32)
    db = torch.zeros(64, device=x.device, dtype=torch.float32)
    _call("conv_first_wgrad", 0, 0, lib().cb200_conv_first_wgrad, ptr(x), ptr(dy), ptr(dw), ptr(db), i32(B), i32(H), i32(W), f32(in_scale), f32(in_shift), stream_ptr())
    return dw, db


def conv_first_dgrad_finish(dpad):
    dpad = _f32c(dpad, "dpad")
    B, H, W, cpad = dpad.shape
    dx = torch.empty(B, 3, H, W, device=dpad.device, dtype=torch.float32)
    _call("conv_first_dgrad_finish", 0, 0, lib().cb200_conv_first_dgrad_finish, ptr(dpad), ptr(dx), i32(B), i32(H), i32(W), i32(cpad), stream_ptr())
    return dx


# ------------------------------------------------------------------ losses
LOSS_KINDS = {"nonsat": 0, "hinge": 1, "wgan": 2, "lsgan": 3}


def rownorm_fwd(x, eps=1e-12):
    """x: [rows, d] (row-strided view allowed) -> (y contiguous, inv_norm)."""
    assert x.dim() == 2 and x.stride(1) == 1
    rows, d = x.shape
    y = torch.empty(rows, d, device=x.device, dtype=torch.float32)
    inv = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("rownorm_fwd", 0, 0, lib().cb200_rownorm_fwd, ptr(x), i64(x.stride(0)), ptr(y), ptr(inv), i32(rows), i32(d), f32(eps),
                                  stream_ptr())
    return y, inv


def rownorm_bwd(dy, y, inv, out=None, round_out=False):
    dy = _f32c(dy, "dy")
    rows, d = y.shape
    if out is None:
        out = torch.empty(rows, d, device=y.device, dtype=torch.float32)
    assert out.stride(1) == 1
    _call("rownorm_bwd", 0, 0, lib().cb200_rownorm_bwd, ptr(dy), ptr(y), ptr(inv), ptr(out), i64(out.stride(0)), i32(rows), i32(d),
                                  i32(1 if round_out else 0), stream_ptr())
    return out


def contrastive_fwd(z, n, mode, temperature):
    """z: [2n or 3n, 128] normalised rows.  Returns (loss[1], lse)."""
    z = _f32c(z, "z")
    active = n if mode else 2 * n
    assert z.shape == ((3 if mode else 2) * n, 128), z.shape
    lse = torch.empty(active, device=z.device, dtype=torch.float32)
    row_loss = torch.empty(48 * active, device=z.device, dtype=torch.float32)      # column-split partials
    loss = torch.empty(1, device=z.device, dtype=torch.float32)
    _call("contrastive_fwd", 0, 0, lib().cb200_contrastive_fwd, ptr(z), i32(n), i32(z.shape[1]), i32(mode), f32(temperature), ptr(lse),
                                      ptr(row_loss), ptr(loss), stream_ptr())
    return loss, lse


def contrastive_bwd(z, n, mode, temperature, lse, gscale):
    z = _f32c(z, "z")
    dz = torch.empty_like(z)
    gscale = _f32c(gscale.reshape(1), "gscale")
    _call("contrastive_bwd", 0, 0, lib().cb200_contrastive_bwd, ptr(z), i32(n), i32(z.shape[1]), i32(mode), f32(temperature), ptr(lse),
                                      ptr(gscale), ptr(dz), stream_ptr())
    return dz


def sim_rows_fwd(S, Ra, R, n, mode, row0, temperature):
    """S [>=Ra, lds] similarity rows -> (lse [Ra], rowloss [Ra]); include/contrad_b200.h cb200_sim_rows_fwd."""
    lse = torch.empty(Ra, device=S.device, dtype=torch.float32)
    rowloss = torch.empty(Ra, device=S.device, dtype=torch.float32)
    _call("sim_rows_fwd", 0, 4 * Ra * R, lib().cb200_sim_rows_fwd, ptr(S), i64(S.stride(0)), i32(Ra), i32(R), i32(n), i32(mode),
          i32(row0), f32(temperature), ptr(lse), ptr(rowloss), stream_ptr())
    return lse, rowloss


def sim_rows_bwd(S, Ra, R, n, mode, row0, temperature, lse, gscale, rows_pad, cols_pad):
    """-> G [rows_pad, cols_pad] (rows >= Ra zero) = gscale * dLoss/dS."""
    G = torch.zeros(rows_pad, cols_pad, device=S.device, dtype=torch.float32) if rows_pad != Ra else \
        torch.empty(rows_pad, cols_pad, device=S.device, dtype=torch.float32)
    gscale = _f32c(gscale.reshape(1), "gscale")
    _call("sim_rows_bwd", 0, 8 * Ra * R, lib().cb200_sim_rows_bwd, ptr(S), i64(S.stride(0)), i32(Ra), i32(R), i32(n), i32(mode),
          i32(row0), f32(temperature), ptr(lse), ptr(gscale), ptr(G), i64(cols_pad), i32(cols_pad), stream_ptr())
    return G


def gan_d_loss(d_real, d_gen, kind, g_real=None, g_gen=None):
    """d_real, d_gen: 1-D views (any element stride, equal).  Returns (out3, g_real, g_gen); g_real / g_gen may be
    given as contiguous [n] destinations (e.g. slices of one gradient vector)."""
    n = d_real.numel()
    assert d_real.dim() == 1 and d_gen.dim() == 1 and d_real.stride(0) == d_gen.stride(0)
    out = torch.empty(3, device=d_real.device, dtype=torch.float32)
    g_r = torch.empty(n, device=d_real.device, dtype=torch.float32) if g_real is None else g_real
    g_g = torch.empty(n, device=d_real.device, dtype=torch.float32) if g_gen is None else g_gen
    assert g_r.numel() == n and g_g.numel() == n and g_r.is_contiguous() and g_g.is_contiguous()
    _call("gan_d_loss", 0, 0, lib().cb200_gan_d_loss, ptr(d_real), ptr(d_gen), i64(d_real.stride(0)), i32(n), i32(LOSS_KINDS[kind]),
                                 ptr(out), ptr(g_r), ptr(g_g), stream_ptr())
    return out, g_r, g_g


def gan_g_loss(d_gen, kind):
    n = d_gen.numel()
    assert d_gen.dim() == 1
    out = torch.empty(1, device=d_gen.device, dtype=torch.float32)
    g = torch.empty(n, device=d_gen.device, dtype=torch.float32)
    k = LOSS_KINDS.get(kind, 2)
    _call("gan_g_loss", 0, 0, lib().cb200_gan_g_loss, ptr(d_gen), i64(d_gen.stride(0)), i32(n), i32(k), ptr(out), ptr(g), stream_ptr())
    return out, g


def lrelu_bwd(dy, act, slope, round_out=False):
    dy = _f32c(dy, "dy")
    act = _f32c(act, "act")
    out = torch.empty_like(act)
    _call("lrelu_bwd", 0, 0, lib().cb200_lrelu_bwd, ptr(dy), ptr(act), ptr(out), i64(act.numel()), f32(slope), i32(1 if round_out else 0),
                                stream_ptr())
    return out


def colsum(x2d):
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    M, N = x2d.shape
    out = torch.empty(N, device=x2d.device, dtype=torch.float32)
    _call("colsum", 0, 0, lib().cb200_colsum, ptr(x2d), i64(x2d.stride(0)), i32(M), i32(N), ptr(out), stream_ptr())
    return out


# ------------------------------------------------------------------ generator-side kernels
def bn_stats(x2d):
    M, C = x2d.shape
    sums = torch.empty(2, C, device=x2d.device, dtype=torch.float32)
    _call("bn_stats", 0, 8 * x2d.numel(), lib().cb200_bn_stats, ptr(x2d), i32(M), i32(C), ptr(sums), stream_ptr())
    return sums


def bn_finalize(sums, count, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    C = sums.shape[1]
    stats = torch.empty(2, C, device=sums.device, dtype=torch.float32)
    _call("bn_finalize", 0, 0, lib().cb200_bn_finalize, ptr(sums), f32(count), i32(C), f32(eps), f32(momentum), ptr(stats),
          ptr(running_mean), ptr(running_var), stream_ptr())
    return stats


def bn_apply_relu(x2d, stats, gamma, beta, remap_s=0, round_out=True):
    M, C = x2d.shape
    y = torch.empty(M, C, device=x2d.device, dtype=torch.float32)
    _call("bn_apply_relu", 0, 8 * x2d.numel(), lib().cb200_bn_apply_relu, ptr(x2d), ptr(stats), ptr(gamma), ptr(beta), ptr(y),
          i32(M), i32(C), i32(remap_s), i32(1 if round_out else 0), stream_ptr())
    return y


def bn_bwd_reduce(dy2d, y2d, x2d, stats, remap_s=0):
    M, C = x2d.shape
    sums = torch.empty(2, C, device=x2d.device, dtype=torch.float32)
    _call("bn_bwd_reduce", 0, 12 * x2d.numel(), lib().cb200_bn_bwd_reduce, ptr(dy2d), ptr(y2d), ptr(x2d), ptr(stats), i32(M),
          i32(C), i32(remap_s), ptr(sums), stream_ptr())
    return sums


def bn_bwd_apply(dy2d, y2d, x2d, stats, gamma, sums, count, remap_s=0, round_out=True):
    M, C = x2d.shape
    dx = torch.empty(M, C, device=x2d.device, dtype=torch.float32)
    _call("bn_bwd_apply", 0, 16 * x2d.numel(), lib().cb200_bn_bwd_apply, ptr(dy2d), ptr(y2d), ptr(x2d), ptr(stats), ptr(gamma),
          ptr(sums), f32(count), ptr(dx), i32(M), i32(C), i32(remap_s), i32(1 if round_out else 0), stream_ptr())
    return dx


def g_final_fwd(pre_nhwc, bias):
    B, H, W, cpad = pre_nhwc.shape
    out = torch.empty(B, 3, H, W, device=pre_nhwc.device, dtype=torch.float32)
    _call("g_final_fwd", 0, 0, lib().cb200_g_final_fwd, ptr(pre_nhwc), ptr(bias), ptr(out), i32(B), i32(H), i32(W), i32(cpad),
          stream_ptr())
    return out


def g_final_bwd(dout, out, need_bias=True):
    dout = _f32c(dout, "dout")
    B, _, H, W = out.shape
    dpre = torch.empty_like(out)
    dbias = torch.empty(3, device=out.device, dtype=torch.float32) if need_bias else None
    _call("g_final_bwd", 0, 0, lib().cb200_g_final_bwd, ptr(dout), ptr(out), ptr(dpre), ptr(dbias), i32(B), i32(H), i32(W),
          stream_ptr())
    return dpre, dbias


def round_tf32_(x):
    """Device-side RNA rounding to TF32 into a new tensor (kernel, not the torch bit trick below)."""
    x = _f32c(x, "x")
    y = torch.empty_like(x)
    _call("round_tf32", 0, 0, lib().cb200_round_tf32, ptr(x), ptr(y), i64(x.numel()), stream_ptr())
    return y


def split_tf32(x, mode):
    """x [..., C] -> error-compensated TF32 operands (include/contrad_b200.h: cb200_split_tf32).
    mode 0: [..., 3C] = hi | lo | hi;  mode 1: [..., 2C] = hi | hi;  mode 2: [2, ..., C] = hi ; lo."""
    x = _f32c(x, "x")
    C = x.shape[-1]
    rows = x.numel() // C
    if mode == 2:
        out = torch.empty((2,) + tuple(x.shape), device=x.device, dtype=torch.float32)
    else:
        out = torch.empty(tuple(x.shape[:-1]) + ((3 if mode == 0 else 2) * C,), device=x.device, dtype=torch.float32)
    _call("split_tf32", 0, 4 * x.numel() * (4 if mode == 0 else 3), lib().cb200_split_tf32, ptr(x), ptr(out), i64(rows),
          i32(C), i32(mode), stream_ptr())
    return out


# ------------------------------------------------------------------ fused Adam
class _AdamTensor(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("numel", ctypes.c_longlong)]


def adam_step(entries, lr, beta1, beta2, eps, step):
    """entries: list of (param, grad, exp_avg, exp_avg_sq) contiguous fp32 CUDA tensors; in-place update."""
    arr = (_AdamTensor * len(entries))()
    nbytes = 0
    for i, (p, g, m, v) in enumerate(entries):
        assert p.is_contiguous() and g.is_contiguous()
        arr[i] = _AdamTensor(ptr(p).value, ptr(g).value, ptr(m).value, ptr(v).value, p.numel())      # ptr() refuses CPU tensors
        nbytes += 28 * p.numel()
    if isinstance(lr, torch.Tensor):
        # device-resident {lr, 1-b1^t, sqrt(1-b2^t)} (CUDA-graph replay); `step` is ignored
        assert lr.is_cuda and lr.dtype == torch.float32 and lr.numel() >= 3 and lr.is_contiguous()
        _call("adam_step", 0, nbytes, lib().cb200_adam_step_dev, arr, i32(len(entries)), ptr(lr), f32(beta1), f32(beta2),
              f32(eps), stream_ptr())
        return
    _call("adam_step", 0, nbytes, lib().cb200_adam_step, arr, i32(len(entries)), f32(lr), f32(beta1), f32(beta2), f32(eps),
          i32(step), stream_ptr())


# ------------------------------------------------------------------ layout helpers (torch ops; test/reference use)
def pack_fwd_weight(w):
    """OIHW -> [Cout, kh*kw*Cin] (tap-major, channel-minor)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


_KSEL = ((1, 3), (0, 2))      # output parity -> the two kernel rows/cols that hit it (stride 2, pad 1, k 4)


def pack_dgrad_weight(w, stride):
    """OIHW -> data-gradient GEMM matrix.
    stride 1 (3x3): [Cin, 9*Cout], column (kh*3+kw)*Cout+co = w[co,ci,kh,kw]
    stride 2 (4x4): [4 (ph*2+pw), Cin, 4 (jh*2+jw), Cout] flattened to [4*Cin, 4*Cout]."""
    if stride == 1:
        return w.permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous()
    cout, cin = w.shape[:2]
    out = w.new_empty(2, 2, cin, 2, 2, cout)
    for ph in range(2):
        for pw in range(2):
            for jh in range(2):
                for jw in range(2):
                    out[ph, pw, :, jh, jw, :] = w[:, :, _KSEL[ph][jh], _KSEL[pw][jw]].t()
    return out.reshape(4 * cin, 4 * cout).contiguous()


def round_tf32(t):
    """Round-to-nearest (ties away) to TF32, the rounding the producer epilogues apply (cvt.rna.tf32.f32)."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
