"""Raw (non-autograd) Python bindings of the C ABI: shape checks, output allocation, stream plumbing.
The autograd Functions in ``contrad_b200.functional`` are built on these."""
import torch

from . import _capi
from ._capi import check, f32, i32, i64, lib, ptr, stream_ptr

PARAM_FIELDS = ("sx", "sy", "bx", "by", "flip", "cj_on", "contrast", "hue", "sat", "val", "gray_on")


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------ augmentation
def augment_simclr_fwd(x, params, order):
    x = _f32c(x, "x")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    assert C == 3 and params.shape == (len(PARAM_FIELDS), B), (x.shape, params.shape)
    y = torch.empty_like(x)
    check(lib().cb200_augment_simclr_fwd(ptr(x), ptr(y), ptr(params), i32(B), i32(H), i32(W), i32(order),
                                         stream_ptr()), "cb200_augment_simclr_fwd")
    return y


def augment_simclr_bwd(x, dy, params, order):
    x = _f32c(x, "x")
    dy = _f32c(dy, "dy")
    params = _f32c(params, "params")
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    check(lib().cb200_augment_simclr_bwd(ptr(x), ptr(dy), ptr(dx), ptr(params), i32(B), i32(H), i32(W), i32(order),
                                         stream_ptr()), "cb200_augment_simclr_bwd")
    return dx


# ------------------------------------------------------------------ tensor-core GEMM / conv
def gemm_nt(a, bw, bias=None, slope=1.0, round_out=False, out=None):
    """out[M,N] = lrelu_slope(a[M,K] @ bw[N,K]^T + bias).  `a` / `out` may be row-strided 2-D views."""
    assert a.dim() == 2 and bw.dim() == 2 and a.shape[1] == bw.shape[1]
    assert a.stride(1) == 1 and bw.is_contiguous()
    M, K = a.shape
    N = bw.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    assert out.shape == (M, N) and out.stride(1) == 1
    check(lib().cb200_gemm_nt_tf32(ptr(a), i64(a.stride(0)), ptr(bw), ptr(bias), ptr(out), i64(out.stride(0)),
                                   i32(M), i32(N), i32(K), f32(slope), i32(1 if round_out else 0), stream_ptr()),
          "cb200_gemm_nt_tf32")
    return out


def conv2d_nhwc_fwd(x, wmat, bias, ks, stride, slope=1.0, round_out=False):
    """x [B,H,W,Cin] NHWC -> y [B,Ho,Wo,Cout]; wmat [Cout, ks*ks*Cin]."""
    x = _f32c(x, "x")
    B, H, W, Cin = x.shape
    Cout = wmat.shape[0]
    assert wmat.shape[1] == ks * ks * Cin and wmat.is_contiguous()
    Ho, Wo = H // stride, W // stride
    y = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.float32)
    check(lib().cb200_conv2d_nhwc_fwd(ptr(x), ptr(wmat), ptr(bias), ptr(y), i32(B), i32(H), i32(W), i32(Cin),
                                      i32(Cout), i32(ks), i32(stride), f32(slope), i32(1 if round_out else 0),
                                      stream_ptr()), "cb200_conv2d_nhwc_fwd")
    return y


def conv2d_nhwc_dgrad(dy, wmat_t, in_shape, ks, stride, act_in=None, bias_out=None, slope=1.0, round_out=False):
    """dy [B,Ho,Wo,Cout] -> dx [B,H,W,Cin] (in_shape).  wmat_t: see pack_dgrad_weight."""
    dy = _f32c(dy, "dy")
    B, H, W, Cin = in_shape
    Cout = dy.shape[3]
    dx = torch.empty(B, H, W, Cin, device=dy.device, dtype=torch.float32)
    check(lib().cb200_conv2d_nhwc_dgrad(ptr(dy), ptr(wmat_t), ptr(act_in), ptr(bias_out), ptr(dx), i32(B), i32(H),
                                        i32(W), i32(Cin), i32(Cout), i32(ks), i32(stride), f32(slope),
                                        i32(1 if round_out else 0), stream_ptr()), "cb200_conv2d_nhwc_dgrad")
    return dx


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
