"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the tcgen05 entry points of contrad_b200.kernels with their COMPLETE
epilogue semantics (include/contrad_b200.h: bias, LeakyReLU, lrelu' masks, TF32 rounding of outputs, fused column sums,
caller-provided strided outputs).  Together with the CUDA emulator for the SIMT kernels (tests/emu) they let the
product's SNDCGAN train step run on the CPU, so its host logic is checked without a GPU.  Written from the header's
definitions, independently of the CUDA sources.  (tests/cpu_kernels.py holds the reduced set the StyleGAN2 host tests
need.)  The product never imports this file."""
import contextlib

import torch
import torch.nn.functional as F

_KSEL = ((1, 3), (0, 2))      # output parity -> kernel rows / cols (4x4, stride 2, pad 1); see kernels.pack_dgrad_weight


def _round(t, on):
    from contrad_b200 import kernels as K
    return K.round_tf32(t) if on else t


def _lrelu(x, slope):
    return x if slope == 1.0 else torch.where(x > 0, x, x * slope)


def _dlrelu(act, slope):
    return torch.where(act > 0, torch.ones_like(act), torch.full_like(act, slope))


def gemm_nt(a, bw, bias=None, slope=1.0, round_out=False, out=None, dact=None, colsum=None):
    y = a @ bw.t()
    if bias is not None:
        y = y + bias
    y = y * _dlrelu(dact, slope) if dact is not None else _lrelu(y, slope)
    y = _round(y, round_out)
    if colsum is not None:
        colsum.copy_(y.sum(0))
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


def gemm_tn_wgrad(dy, x, out=None):
    r = dy.t() @ x
    if out is not None:
        out.copy_(r)
        return out
    return r


def _oihw_from_fwd(wmat, ks, cin):
    return wmat.view(wmat.shape[0], ks, ks, cin).permute(0, 3, 1, 2)


def _oihw_from_dgrad(wmat_t, ks, stride, cin, cout):
    if stride == 1:                                       # [Cin, 9*Cout], column (kh*3+kw)*Cout+co
        return wmat_t.view(cin, ks, ks, cout).permute(3, 0, 1, 2)
    packed = wmat_t.view(2, 2, cin, 2, 2, cout)           # [ph][pw][Cin][jh][jw][Cout]
    w = wmat_t.new_zeros(cout, cin, 4, 4)
    for ph in range(2):
        for pw in range(2):
            for jh in range(2):
                for jw in range(2):
                    w[:, :, _KSEL[ph][jh], _KSEL[pw][jw]] = packed[ph, pw, :, jh, jw, :].t()
    return w


def conv2d_nhwc_fwd(x, wmat, bias, ks, stride, slope=1.0, round_out=False):
    y = F.conv2d(x.permute(0, 3, 1, 2), _oihw_from_fwd(wmat, ks, x.shape[3]), bias, stride=stride, padding=1)
    return _round(_lrelu(y, slope), round_out).permute(0, 2, 3, 1).contiguous()


def conv2d_nhwc_dgrad(dy, wmat_t, in_shape, ks, stride, act_in=None, bias_out=None, slope=1.0, round_out=False,
                      colsum=None):
    B, H, W, cin = in_shape
    w = _oihw_from_dgrad(wmat_t, ks, stride, cin, dy.shape[3])
    dx = F.conv_transpose2d(dy.permute(0, 3, 1, 2), w, stride=stride, padding=1).permute(0, 2, 3, 1)
    if act_in is not None:
        dx = dx * _dlrelu(act_in, slope)
    else:
        if bias_out is not None:
            dx = dx + bias_out
        dx = _lrelu(dx, slope)
    dx = _round(dx, round_out).contiguous()
    if colsum is not None:
        colsum.copy_(dx.reshape(-1, cin).sum(0))
    return dx


def conv2d_nhwc_wgrad(x, dy, ks, stride):
    cin, cout = x.shape[3], dy.shape[3]
    dw = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (cout, cin, ks, ks), dy.permute(0, 3, 1, 2), stride=stride,
                                     padding=1)
    return dw.permute(0, 2, 3, 1).reshape(cout, ks * ks * cin).contiguous()


NAMES = ("gemm_nt", "gemm_tn_wgrad", "conv2d_nhwc_fwd", "conv2d_nhwc_dgrad", "conv2d_nhwc_wgrad")


@contextlib.contextmanager
def patched():
    import sys
    from contrad_b200 import kernels as K
    me = sys.modules[__name__]
    saved = [(n, getattr(K, n)) for n in NAMES]
    try:
        for n, _ in saved:
            setattr(K, n, getattr(me, n))
        yield
    finally:
        for n, fn in saved:
            setattr(K, n, fn)
