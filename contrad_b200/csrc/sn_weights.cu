// Spectral normalisation + weight packing for the tensor-core kernels (sm_100a, HBM-bound SIMT).
//
// Reference: torch.nn.utils.spectral_norm as applied by models/gan/sndcgan.py:111-118 to every
// Conv2d / Linear of D_SNDCGAN (arithmetic: torch/nn/utils/spectral_norm.py:92-114): per forward in
// train mode ONE power iteration, in place and without grad,
//        v <- normalize(W^T u),  u <- normalize(W v),  sigma = u^T W v,  W_hat = W / sigma,
// ~200 tiny ATen ops per D forward in the reference.  Here:
//   sn_wtu   : t  = W^T u            (column sums, coalesced along F, atomics across row chunks)
//   sn_wv    : s' = W t              (one CTA per row) - linear in t, so no normalisation pass in between
//   sn_final : v = t/|t|, s = s'/|t|, u = s/|s|, sigma = u.s          (one CTA)
//   sn_pack  : W/sigma -> GEMM layouts ([Cout][kh][kw][Cin] forward, transposed / parity-class
//              layouts for the data-gradient), rounded to TF32 because they feed tcgen05.mma
//   sn_bwd   : dW = (dW_hat - <dW_hat, W_hat> u v^T) / sigma, un-packing the wgrad layout
// sigma never leaves the device.
#include "common.cuh"

namespace {

constexpr int kT = 256;

// t[f] += sum_{o in chunk} u[o] * W[o, f]
__global__ void __launch_bounds__(kT) sn_wtu_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                                    float* __restrict__ t, int Cout, int F, int rows_per_cta) {
    const int f = (blockIdx.x * kT + threadIdx.x) * 4;
    if (f >= F) return;
    const int o0 = blockIdx.y * rows_per_cta;
    const int o1 = min(Cout, o0 + rows_per_cta);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (f + 3 < F) && ((F & 3) == 0);
    for (int o = o0; o < o1; ++o) {
        const float uo = __ldg(u + o);
        const float* row = w + (size_t)o * F + f;
        if (vec) {
            float4 x = __ldg(reinterpret_cast<const float4*>(row));
            acc[0] += uo * x.x; acc[1] += uo * x.y; acc[2] += uo * x.z; acc[3] += uo * x.w;
        } else {
            for (int e = 0; e < 4 && f + e < F; ++e) acc[e] += uo * __ldg(row + e);
        }
    }
    for (int e = 0; e < 4 && f + e < F; ++e) atomicAdd(t + f + e, acc[e]);
}

// s[o] = dot(W[o,:], t)
__global__ void __launch_bounds__(kT) sn_wv_kernel(const float* __restrict__ w, const float* __restrict__ t,
                                                   float* __restrict__ s, int F) {
    __shared__ float red[32];
    const int o = blockIdx.x;
    const float* row = w + (size_t)o * F;
    float acc[1] = {0.f};
    if ((F & 3) == 0) {
        for (int f = threadIdx.x * 4; f < F; f += kT * 4) {
            float4 x = __ldg(reinterpret_cast<const float4*>(row + f));
            float4 y = __ldg(reinterpret_cast<const float4*>(t + f));
            acc[0] += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
    } else {
        for (int f = threadIdx.x; f < F; f += kT) acc[0] += __ldg(row + f) * __ldg(t + f);
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) s[o] = acc[0];
}

// training: t = W^T u (unnormalised), s = W t.   eval: t = v (given), s = W v.
// out: v, u updated (training only), sigma[0] = u . (W v), sigma[1] = 1/sigma
__global__ void __launch_bounds__(kT) sn_final_kernel(const float* __restrict__ t, const float* __restrict__ s,
                                                      float* __restrict__ u, float* __restrict__ v,
                                                      float* __restrict__ sigma, int Cout, int F, float eps,
                                                      int training) {
    __shared__ float red[32];
    float a[1];
    if (training) {
        a[0] = 0.f;
        for (int f = threadIdx.x; f < F; f += kT) { float x = t[f]; a[0] += x * x; }
        block_sum<1>(a, red);
        const float tn = fmaxf(sqrtf(a[0]), eps);
        for (int f = threadIdx.x; f < F; f += kT) v[f] = t[f] / tn;
        a[0] = 0.f;
        for (int o = threadIdx.x; o < Cout; o += kT) { float x = s[o] / tn; a[0] += x * x; }
        block_sum<1>(a, red);
        const float sn = fmaxf(sqrtf(a[0]), eps);
        a[0] = 0.f;
        for (int o = threadIdx.x; o < Cout; o += kT) {
            float wv = s[o] / tn;
            float un = wv / sn;
            u[o] = un;
            a[0] += un * wv;
        }
        block_sum<1>(a, red);
    } else {
        a[0] = 0.f;
        for (int o = threadIdx.x; o < Cout; o += kT) a[0] += u[o] * s[o];
        block_sum<1>(a, red);
    }
    if (threadIdx.x == 0) {
        sigma[0] = a[0];
        sigma[1] = 1.f / a[0];
    }
}

// One thread per source element of W[Cout][Cin][KH][KW]; writes up to two packed copies scaled by 1/sigma.
//   fwd   : [Cout][KH][KW][Cin] with row stride ld_fwd (lets several heads share one concatenated matrix)
//   dgrad : mode 1 -> [Cin][KH][KW][Cout]            (3x3 stride 1, linear layers with KH=KW=1)
//           mode 2 -> [ph][pw][Cin][jh][jw][Cout]    (4x4 stride 2; kh = {1,3}/{0,2}[ph][jh])
//           mode 3 -> [KH][KW][Cin][ldt]             (transposed matrix of a flattened-feature linear layer:
//                                                      row (kh,kw,ci), column offset col0 + co)
__global__ void __launch_bounds__(kT) sn_pack_kernel(const float* __restrict__ w, const float* __restrict__ sigma,
                                                     float* __restrict__ fwd, long long ld_fwd,
                                                     float* __restrict__ dg, int dg_mode, long long ldt, int col0,
                                                     int Cout, int Cin, int KH, int KW, int do_round) {
    const long long total = (long long)Cout * Cin * KH * KW;
    const long long idx = (long long)blockIdx.x * kT + threadIdx.x;
    if (idx >= total) return;
    const float inv = sigma ? sigma[1] : 1.f;
    int kw = (int)(idx % KW);
    long long r = idx / KW;
    int kh = (int)(r % KH); r /= KH;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    float val = w[idx] * inv;
    if (do_round) val = round_tf32(val);
    if (fwd) fwd[(long long)co * ld_fwd + ((long long)(kh * KW + kw)) * Cin + ci] = val;
    if (dg) {
        if (dg_mode == 1) {
            dg[(((long long)ci * KH + kh) * KW + kw) * Cout + co] = val;
        } else if (dg_mode == 2) {
            const int ph = (kh & 1) ? 0 : 1, jh = (kh == 1 || kh == 0) ? 0 : 1;
            const int pw = (kw & 1) ? 0 : 1, jw = (kw == 1 || kw == 0) ? 0 : 1;
            dg[((((long long)(ph * 2 + pw) * Cin + ci) * 2 + jh) * 2 + jw) * Cout + co] = val;
        } else {
            dg[(((long long)kh * KW + kw) * Cin + ci) * ldt + col0 + co] = val;
        }
    }
}

// acc[0] += <dW_hat(packed fwd layout), W> (elementwise over the layer)
__global__ void __launch_bounds__(kT) sn_bwd_dot_kernel(const float* __restrict__ dwp, long long ld_fwd,
                                                        const float* __restrict__ w, float* __restrict__ acc,
                                                        int Cout, int Cin, int KH, int KW) {
    __shared__ float red[32];
    const long long total = (long long)Cout * Cin * KH * KW;
    float a[1] = {0.f};
    for (long long idx = (long long)blockIdx.x * kT + threadIdx.x; idx < total; idx += (long long)gridDim.x * kT) {
        int kw = (int)(idx % KW);
        long long r = idx / KW;
        int kh = (int)(r % KH); r /= KH;
        int ci = (int)(r % Cin);
        int co = (int)(r / Cin);
        a[0] += dwp[(long long)co * ld_fwd + ((long long)(kh * KW + kw)) * Cin + ci] * w[idx];
    }
    block_sum<1>(a, red);
    if (threadIdx.x == 0) atomicAdd(acc, a[0]);
}

// dW[idx] (+)= (dW_hat - (acc/sigma^2) * u[co] * v[f]) / sigma ... with <dW_hat, W_hat> = acc / sigma
__global__ void __launch_bounds__(kT) sn_bwd_apply_kernel(const float* __restrict__ dwp, long long ld_fwd,
                                                          const float* __restrict__ u, const float* __restrict__ v,
                                                          const float* __restrict__ sigma,
                                                          const float* __restrict__ acc, float* __restrict__ dw,
                                                          int accumulate, int Cout, int Cin, int KH, int KW) {
    const long long total = (long long)Cout * Cin * KH * KW;
    const long long idx = (long long)blockIdx.x * kT + threadIdx.x;
    if (idx >= total) return;
    int kw = (int)(idx % KW);
    long long r = idx / KW;
    int kh = (int)(r % KH); r /= KH;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    const int f = (int)(idx - (long long)co * Cin * KH * KW);
    float g = dwp[(long long)co * ld_fwd + ((long long)(kh * KW + kw)) * Cin + ci];
    float out;
    if (sigma) {
        const float inv = sigma[1];
        out = (g - acc[0] * inv * u[co] * v[f]) * inv;
    } else {
        out = g;
    }
    dw[idx] = accumulate ? dw[idx] + out : out;
}

}  // namespace

// One power iteration (training != 0) or sigma from the stored u, v (training == 0).
//   w [Cout, F] (any OIHW weight viewed as a matrix), u [Cout], v [F] updated in place,
//   sigma [2] <- {sigma, 1/sigma};  scratch: t [F] and s [Cout] floats.
extern "C" int cb200_sn_power_iter(const float* w, float* u, float* v, float* sigma, float* t_scratch,
                                   float* s_scratch, int Cout, int F, float eps, int training, void* stream) {
    CB200_CHECK_ARG(Cout > 0 && F > 0, "sn_power_iter: empty weight");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* tvec = v;
    if (training) {
        cudaError_t e = cudaMemsetAsync(t_scratch, 0, sizeof(float) * F, st);
        if (e != cudaSuccess) { cb200_set_error("sn_power_iter: memset: %s", cudaGetErrorString(e)); return (int)e; }
        const int rows_per_cta = 64;
        dim3 grid((F + kT * 4 - 1) / (kT * 4), (Cout + rows_per_cta - 1) / rows_per_cta);
        sn_wtu_kernel<<<grid, kT, 0, st>>>(w, u, t_scratch, Cout, F, rows_per_cta);
        CB200_COUNT_LAUNCH();
        tvec = t_scratch;
    }
    sn_wv_kernel<<<Cout, kT, 0, st>>>(w, tvec, s_scratch, F);
    CB200_COUNT_LAUNCH();
    sn_final_kernel<<<1, kT, 0, st>>>(t_scratch, s_scratch, u, v, sigma, Cout, F, eps, training);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sn_power_iter");
    return CB200_OK;
}

// W[Cout,Cin,KH,KW] * (1/sigma) -> packed GEMM layouts (see sn_pack_kernel).  sigma may be NULL (scale 1).
extern "C" int cb200_sn_pack_weights(const float* w, const float* sigma, float* fwd, long long ld_fwd, float* dgrad,
                                     int dgrad_mode, long long ldt, int col0, int Cout, int Cin, int KH, int KW,
                                     int round_out, void* stream) {
    CB200_CHECK_ARG(Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "sn_pack_weights: empty weight");
    CB200_CHECK_ARG(dgrad == nullptr || (dgrad_mode >= 1 && dgrad_mode <= 3), "sn_pack_weights: bad dgrad_mode");
    CB200_CHECK_ARG(dgrad_mode != 2 || (KH == 4 && KW == 4), "sn_pack_weights: parity-class layout needs a 4x4 kernel");
    const long long total = (long long)Cout * Cin * KH * KW;
    sn_pack_kernel<<<(unsigned)((total + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(
        w, sigma, fwd, ld_fwd, dgrad, dgrad_mode, ldt, col0, Cout, Cin, KH, KW, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sn_pack_weights");
    return CB200_OK;
}

// dW[Cout,Cin,KH,KW] (+)= d(W/sigma)/dW applied to dW_hat given in the forward-pack layout.
// acc_scratch: one float.  sigma == NULL -> plain un-packing (layers without spectral norm).
extern "C" int cb200_sn_weight_bwd(const float* dw_hat_packed, long long ld_fwd, const float* w, const float* u,
                                   const float* v, const float* sigma, float* acc_scratch, float* dw, int accumulate,
                                   int Cout, int Cin, int KH, int KW, void* stream) {
    CB200_CHECK_ARG(Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "sn_weight_bwd: empty weight");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)Cout * Cin * KH * KW;
    if (sigma) {
        cudaError_t e = cudaMemsetAsync(acc_scratch, 0, sizeof(float), st);
        if (e != cudaSuccess) { cb200_set_error("sn_weight_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
        long long blocks = (total + kT - 1) / kT;
        if (blocks > 592) blocks = 592;
        sn_bwd_dot_kernel<<<(unsigned)blocks, kT, 0, st>>>(dw_hat_packed, ld_fwd, w, acc_scratch, Cout, Cin, KH, KW);
        CB200_COUNT_LAUNCH();
    }
    sn_bwd_apply_kernel<<<(unsigned)((total + kT - 1) / kT), kT, 0, st>>>(dw_hat_packed, ld_fwd, u, v, sigma,
                                                                         acc_scratch, dw, accumulate, Cout, Cin, KH,
                                                                         KW);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sn_weight_bwd");
    return CB200_OK;
}
