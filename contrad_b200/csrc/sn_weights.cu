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
#include <stdlib.h>

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


// ================================================================================================
// Batched variants: every spectrally-normalised layer of the discriminator in ONE launch per phase
// (4 launches + 1 memset per D forward instead of ~70).  Layer tables travel as kernel parameters.
// ================================================================================================
constexpr int kMaxLayers = 16;

struct SnPowerBatch {
    const float* w[kMaxLayers];
    float* u[kMaxLayers];
    float* v[kMaxLayers];
    float* sigma[kMaxLayers];
    float* t[kMaxLayers];
    float* s[kMaxLayers];
    int cout[kMaxLayers];
    int f[kMaxLayers];
    int cta_begin[kMaxLayers + 1];   // prefix sums of CTAs per layer for the current phase
    int rows_per_cta;
    int n;
};

__device__ __forceinline__ int find_layer(const int* cta_begin, int n, int cta) {
    int l = 0;
#pragma unroll 1
    while (l + 1 < n && cta >= cta_begin[l + 1]) ++l;
    return l;
}

__global__ void __launch_bounds__(kT) sn_wtu_batched_kernel(const __grid_constant__ SnPowerBatch p) {
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int local = blockIdx.x - p.cta_begin[l];
    const int F = p.f[l], Cout = p.cout[l];
    const int col_ctas = (F + kT * 4 - 1) / (kT * 4);
    const int f = ((local % col_ctas) * kT + threadIdx.x) * 4;
    if (f >= F) return;
    const int o0 = (local / col_ctas) * p.rows_per_cta;
    const int o1 = min(Cout, o0 + p.rows_per_cta);
    const float* __restrict__ w = p.w[l];
    const float* __restrict__ u = p.u[l];
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
    float* t = p.t[l];
    for (int e = 0; e < 4 && f + e < F; ++e) atomicAdd(t + f + e, acc[e]);
}

__global__ void __launch_bounds__(kT) sn_wv_batched_kernel(const __grid_constant__ SnPowerBatch p, int training) {
    __shared__ float red[32];
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int o = blockIdx.x - p.cta_begin[l];
    const int F = p.f[l];
    const float* __restrict__ row = p.w[l] + (size_t)o * F;
    const float* __restrict__ t = training ? p.t[l] : p.v[l];
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
    if (threadIdx.x == 0) p.s[l][o] = acc[0];
}

__global__ void __launch_bounds__(kT) sn_final_batched_kernel(const __grid_constant__ SnPowerBatch p, float eps,
                                                              int training) {
    __shared__ float red[32];
    const int l = blockIdx.x;
    const int F = p.f[l], Cout = p.cout[l];
    const float* t = p.t[l];
    const float* s = p.s[l];
    float* u = p.u[l];
    float* v = p.v[l];
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
        p.sigma[l][0] = a[0];
        p.sigma[l][1] = 1.f / a[0];
    }
}

struct SnPackBatch {
    const float* w[kMaxLayers];
    const float* sigma[kMaxLayers];
    float* fwd[kMaxLayers];
    float* dg[kMaxLayers];
    long long ld_fwd[kMaxLayers];
    long long ldt[kMaxLayers];
    int cout[kMaxLayers], cin[kMaxLayers], kh[kMaxLayers], kw[kMaxLayers];
    int dg_mode[kMaxLayers], col0[kMaxLayers], do_round[kMaxLayers];
    int cta_begin[kMaxLayers + 1];
    int n;
};

__device__ __forceinline__ void pack_one(const float* __restrict__ w, const float* __restrict__ sigma,
                                         float* __restrict__ fwd, long long ld_fwd, float* __restrict__ dg, int dg_mode,
                                         long long ldt, int col0, int Cout, int Cin, int KH, int KW, int do_round,
                                         long long idx) {
    const long long total = (long long)Cout * Cin * KH * KW;
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

__global__ void __launch_bounds__(kT) sn_pack_batched_kernel(const __grid_constant__ SnPackBatch p) {
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const long long idx = ((long long)(blockIdx.x - p.cta_begin[l]) * kT + threadIdx.x) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        pack_one(p.w[l], p.sigma[l], p.fwd[l], p.ld_fwd[l], p.dg[l], p.dg_mode[l], p.ldt[l], p.col0[l], p.cout[l], p.cin[l],
                 p.kh[l], p.kw[l], p.do_round[l], idx + e);
}

struct SnBwdBatch {
    const float* dwp[kMaxLayers];
    const float* w[kMaxLayers];
    const float* u[kMaxLayers];
    const float* v[kMaxLayers];
    const float* sigma[kMaxLayers];
    float* acc[kMaxLayers];
    float* dw[kMaxLayers];
    long long ld_fwd[kMaxLayers];
    int cout[kMaxLayers], cin[kMaxLayers], kh[kMaxLayers], kw[kMaxLayers];
    int cta_begin[kMaxLayers + 1];
    int n;
};

__global__ void __launch_bounds__(kT) sn_bwd_dot_batched_kernel(const __grid_constant__ SnBwdBatch p) {
    __shared__ float red[32];
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int local = blockIdx.x - p.cta_begin[l];
    const int nctas = p.cta_begin[l + 1] - p.cta_begin[l];
    const int Cin = p.cin[l], KH = p.kh[l], KW = p.kw[l];
    const long long total = (long long)p.cout[l] * Cin * KH * KW;
    const float* __restrict__ dwp = p.dwp[l];
    const float* __restrict__ w = p.w[l];
    const long long ld = p.ld_fwd[l];
    float a[1] = {0.f};
    for (long long idx = (long long)local * kT + threadIdx.x; idx < total; idx += (long long)nctas * kT) {
        int kw = (int)(idx % KW);
        long long r = idx / KW;
        int kh = (int)(r % KH); r /= KH;
        int ci = (int)(r % Cin);
        int co = (int)(r / Cin);
        a[0] += dwp[(long long)co * ld + ((long long)(kh * KW + kw)) * Cin + ci] * w[idx];
    }
    block_sum<1>(a, red);
    if (threadIdx.x == 0) atomicAdd(p.acc[l], a[0]);
}

__global__ void __launch_bounds__(kT) sn_bwd_apply_batched_kernel(const __grid_constant__ SnBwdBatch p) {
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int Cin = p.cin[l], KH = p.kh[l], KW = p.kw[l];
    const long long total = (long long)p.cout[l] * Cin * KH * KW;
    const long long idx = (long long)(blockIdx.x - p.cta_begin[l]) * kT + threadIdx.x;
    if (idx >= total) return;
    int kw = (int)(idx % KW);
    long long r = idx / KW;
    int kh = (int)(r % KH); r /= KH;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    const int f = (int)(idx - (long long)co * Cin * KH * KW);
    const float g = p.dwp[l][(long long)co * p.ld_fwd[l] + ((long long)(kh * KW + kw)) * Cin + ci];
    const float inv = p.sigma[l][1];
    p.dw[l][idx] = (g - p.acc[l][0] * inv * p.u[l][co] * p.v[l][f]) * inv;
}

// ------------------------------------------------------------------------------------------------
// Tiled variants of the pack / backward kernels.  A CTA owns a 32 (Cout) x CT (Cin) x KK (taps) block of one layer
// and moves it through shared memory, so that BOTH sides of every layout change are coalesced:
//   OIHW weight   w[co][ci][kk]            - contiguous along (ci, kk) for a fixed co
//   forward pack  fwd[co][kk][ci]          - contiguous along ci
//   dgrad packs   dg[g(ci, kk)][co]        - contiguous along co (all three modes)
// The element-indexed kernels above issue one 4-byte scattered store (a 32-byte sector read-modify-write) per
// element on two of the three layouts; on the 17 M head weights that was 0.75 ms per step.
// Shared layout: tile[co_l][kk * (CT + 1) + ci_l], row stride RS odd - conflict-free for the ci-major and the
// co-major accesses, 2-way for the OIHW-ordered one.
// ------------------------------------------------------------------------------------------------
constexpr int kTileCo = 32;

struct TileGeom {
    int co0, ci0, n_co, n_ci, KK, CT, RS;
};

__device__ __forceinline__ TileGeom tile_geom(int local_cta, int Cout, int Cin, int KK) {
    TileGeom g;
    g.KK = KK;
    g.CT = (KK == 1) ? 256 : 32;
    const int ci_tiles = (Cin + g.CT - 1) / g.CT;
    g.co0 = (local_cta / ci_tiles) * kTileCo;
    g.ci0 = (local_cta % ci_tiles) * g.CT;
    g.n_co = min(kTileCo, Cout - g.co0);
    g.n_ci = min(g.CT, Cin - g.ci0);
    g.RS = (KK * (g.CT + 1)) | 1;
    return g;
}

__host__ __device__ __forceinline__ int tiles_of(int Cout, int Cin, int KK) {
    const int CT = (KK == 1) ? 256 : 32;
    return ((Cout + kTileCo - 1) / kTileCo) * ((Cin + CT - 1) / CT);
}

constexpr size_t kTileSmemBytes = (size_t)kTileCo * ((16 * 33) | 1) * sizeof(float);     // KK <= 16

// KKc > 0: the tap count as a compile-time constant (the divisions by it become multiplies); 0 = generic.
template <int KKc>
__device__ __forceinline__ void pack_tile_body(const SnPackBatch& p, int l, const TileGeom& g, float* tile) {
    const int Cout = p.cout[l], Cin = p.cin[l], KW = p.kw[l], KK = KKc ? KKc : g.KK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long F = (long long)Cin * KK;
    const float inv = p.sigma[l] ? p.sigma[l][1] : 1.f;
    const float* __restrict__ w = p.w[l];
    const int ncols = g.n_ci * KK;
    const bool rnd = p.do_round[l] != 0;
    // Full tiles with a compile-time tap count issue ALL loads of a row (KK per lane) before the first shared-memory
    // store: with 4 in flight (the unroll of the generic loop) three resident CTAs kept ~12 KB per SM in flight, a
    // third of what the HBM latency-bandwidth product needs (1.9 TB/s in the round-1 launch list).
    const bool full = KKc > 1 && g.n_ci == 32 && g.CT == 32;
    constexpr int KKd = KKc > 1 ? KKc : 2;          // divisor on the full-tile path (never executed for KKc <= 1)
    for (int r = warp; r < g.n_co; r += kT / 32) {
        const float* src = w + (long long)(g.co0 + r) * F + (long long)g.ci0 * KK;
        float* row = tile + r * g.RS;
        if (full) {
            float val[KKc ? KKc : 1];
#pragma unroll
            for (int i = 0; i < KKc; ++i) val[i] = __ldg(src + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < KKc; ++i) {
                const int t = lane + 32 * i;
                float o = val[i] * inv;
                if (rnd) o = round_tf32(o);
                row[(t % KKd) * 33 + t / KKd] = o;
            }
            continue;
        }
#pragma unroll 4
        for (int t = lane; t < ncols; t += 32) {
            float val = __ldg(src + t) * inv;
            if (rnd) val = round_tf32(val);
            row[(t % KK) * (g.CT + 1) + t / KK] = val;
        }
    }
    __syncthreads();
    float* __restrict__ fwd = p.fwd[l];
    if (fwd) {
        for (int r = warp; r < g.n_co; r += kT / 32) {
            float* dst = fwd + (long long)(g.co0 + r) * p.ld_fwd[l] + g.ci0;
            const float* row = tile + r * g.RS;
            for (int kk = 0; kk < KK; ++kk)
#pragma unroll 4
                for (int ci = lane; ci < g.n_ci; ci += 32) dst[(long long)kk * Cin + ci] = row[kk * (g.CT + 1) + ci];
        }
    }
    float* __restrict__ dg = p.dg[l];
    if (dg && lane < g.n_co) {
        const int mode = p.dg_mode[l];
        const float* col = tile + lane * g.RS;
        for (int ci_l = warp; ci_l < g.n_ci; ci_l += kT / 32) {
            const int ci = g.ci0 + ci_l;
#pragma unroll 4
            for (int kk = 0; kk < KK; ++kk) {
                long long row;
                if (mode == 1) {
                    row = ((long long)ci * KK + kk) * Cout;
                } else if (mode == 2) {
                    const int kh = kk / KW, kw = kk % KW;
                    const int ph = (kh & 1) ? 0 : 1, jh = (kh <= 1) ? 0 : 1;
                    const int pw = (kw & 1) ? 0 : 1, jw = (kw <= 1) ? 0 : 1;
                    row = ((((long long)(ph * 2 + pw) * Cin + ci) * 2 + jh) * 2 + jw) * Cout;
                } else {
                    row = ((long long)kk * Cin + ci) * p.ldt[l] + p.col0[l];
                }
                dg[row + g.co0 + lane] = col[kk * (g.CT + 1) + ci_l];
            }
        }
    }
}

__global__ void __launch_bounds__(kT) sn_pack_tiled_kernel(const __grid_constant__ SnPackBatch p) {
    extern __shared__ float tile[];
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int KK = p.kh[l] * p.kw[l];
    const TileGeom g = tile_geom(blockIdx.x - p.cta_begin[l], p.cout[l], p.cin[l], KK);
    if (KK == 1) pack_tile_body<1>(p, l, g, tile);
    else if (KK == 9) pack_tile_body<9>(p, l, g, tile);
    else if (KK == 16) pack_tile_body<16>(p, l, g, tile);
    else pack_tile_body<0>(p, l, g, tile);
}

// Stage dW_hat (forward-pack order) of one tile into shared memory: coalesced along ci.  KKc > 1 on a full tile: the KK
// loads of a row are issued back to back (the generic loop has one load in flight per lane).
template <int KKc>
__device__ __forceinline__ void load_dwp_tile(const float* __restrict__ dwp, long long ld, int Cin, const TileGeom& g,
                                              float* tile) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool full = KKc > 1 && g.n_ci == 32 && g.CT == 32;
    for (int r = warp; r < g.n_co; r += kT / 32) {
        const float* src = dwp + (long long)(g.co0 + r) * ld + g.ci0;
        float* row = tile + r * g.RS;
        if (full) {
            float val[KKc ? KKc : 1];
#pragma unroll
            for (int kk = 0; kk < KKc; ++kk) val[kk] = __ldg(src + (long long)kk * Cin + lane);
#pragma unroll
            for (int kk = 0; kk < KKc; ++kk) row[kk * 33 + lane] = val[kk];
            continue;
        }
        for (int kk = 0; kk < g.KK; ++kk)
#pragma unroll 4
            for (int ci = lane; ci < g.n_ci; ci += 32) row[kk * (g.CT + 1) + ci] = __ldg(src + (long long)kk * Cin + ci);
    }
    __syncthreads();
}

template <int KKc>
__device__ __forceinline__ float dot_tile_body(const SnBwdBatch& p, int l, const TileGeom& g, const float* tile) {
    const int KK = KKc ? KKc : g.KK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long F = (long long)p.cin[l] * KK;
    const int ncols = g.n_ci * KK;
    float a = 0.f;
    const bool full = KKc > 1 && g.n_ci == 32 && g.CT == 32;
    constexpr int KKd = KKc > 1 ? KKc : 2;          // divisor on the full-tile path (never executed for KKc <= 1)
    for (int r = warp; r < g.n_co; r += kT / 32) {
        const float* src = p.w[l] + (long long)(g.co0 + r) * F + (long long)g.ci0 * KK;
        const float* row = tile + r * g.RS;
        if (full) {
            float val[KKc ? KKc : 1];
#pragma unroll
            for (int i = 0; i < KKc; ++i) val[i] = __ldg(src + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < KKc; ++i) {
                const int t = lane + 32 * i;
                a += val[i] * row[(t % KKd) * 33 + t / KKd];
            }
            continue;
        }
#pragma unroll 4
        for (int t = lane; t < ncols; t += 32) a += __ldg(src + t) * row[(t % KK) * (g.CT + 1) + t / KK];
    }
    return a;
}

__global__ void __launch_bounds__(kT) sn_bwd_dot_tiled_kernel(const __grid_constant__ SnBwdBatch p) {
    extern __shared__ float tile[];
    __shared__ float red[32];
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int Cin = p.cin[l], KK = p.kh[l] * p.kw[l];
    const TileGeom g = tile_geom(blockIdx.x - p.cta_begin[l], p.cout[l], Cin, KK);
    if (KK == 9) load_dwp_tile<9>(p.dwp[l], p.ld_fwd[l], Cin, g, tile);
    else if (KK == 16) load_dwp_tile<16>(p.dwp[l], p.ld_fwd[l], Cin, g, tile);
    else load_dwp_tile<0>(p.dwp[l], p.ld_fwd[l], Cin, g, tile);
    float a[1];
    if (KK == 1) a[0] = dot_tile_body<1>(p, l, g, tile);
    else if (KK == 9) a[0] = dot_tile_body<9>(p, l, g, tile);
    else if (KK == 16) a[0] = dot_tile_body<16>(p, l, g, tile);
    else a[0] = dot_tile_body<0>(p, l, g, tile);
    block_sum<1>(a, red);
    if (threadIdx.x == 0) atomicAdd(p.acc[l], a[0]);
}

template <int KKc>
__device__ __forceinline__ void apply_tile_body(const SnBwdBatch& p, int l, const TileGeom& g, const float* tile) {
    const int KK = KKc ? KKc : g.KK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long F = (long long)p.cin[l] * KK;
    const int ncols = g.n_ci * KK;
    const float inv = p.sigma[l][1];
    const float scale = p.acc[l][0] * inv;
    const float* __restrict__ v = p.v[l] + (long long)g.ci0 * KK;
    constexpr int KKd = KKc > 1 ? KKc : 2;
    for (int r = warp; r < g.n_co; r += kT / 32) {
        const float su = scale * p.u[l][g.co0 + r];
        float* dst = p.dw[l] + (long long)(g.co0 + r) * F + (long long)g.ci0 * KK;
        const float* row = tile + r * g.RS;
        if (KKc > 1 && g.n_ci == 32 && g.CT == 32) {
            float vv[KKc ? KKc : 1];
#pragma unroll
            for (int i = 0; i < KKc; ++i) vv[i] = __ldg(v + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < KKc; ++i) {
                const int t = lane + 32 * i;
                dst[t] = (row[(t % KKd) * 33 + t / KKd] - su * vv[i]) * inv;
            }
            continue;
        }
#pragma unroll 4
        for (int t = lane; t < ncols; t += 32) dst[t] = (row[(t % KK) * (g.CT + 1) + t / KK] - su * __ldg(v + t)) * inv;
    }
}

__global__ void __launch_bounds__(kT) sn_bwd_apply_tiled_kernel(const __grid_constant__ SnBwdBatch p) {
    extern __shared__ float tile[];
    const int l = find_layer(p.cta_begin, p.n, blockIdx.x);
    const int Cin = p.cin[l], KK = p.kh[l] * p.kw[l];
    const TileGeom g = tile_geom(blockIdx.x - p.cta_begin[l], p.cout[l], Cin, KK);
    if (KK == 9) load_dwp_tile<9>(p.dwp[l], p.ld_fwd[l], Cin, g, tile);
    else if (KK == 16) load_dwp_tile<16>(p.dwp[l], p.ld_fwd[l], Cin, g, tile);
    else load_dwp_tile<0>(p.dwp[l], p.ld_fwd[l], Cin, g, tile);
    if (KK == 1) apply_tile_body<1>(p, l, g, tile);
    else if (KK == 9) apply_tile_body<9>(p, l, g, tile);
    else if (KK == 16) apply_tile_body<16>(p, l, g, tile);
    else apply_tile_body<0>(p, l, g, tile);
}

bool sn_tiled_enabled() {
    static const bool on = []() { const char* e = getenv("CB200_SN_TILED"); return !(e && e[0] == '0'); }();
    return on;
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


// ------------------------------------------------------------------------------------------------
// Batched C ABI: plain-C descriptor arrays (host memory), one entry per layer / pack job.
// ------------------------------------------------------------------------------------------------
extern "C" {
struct cb200_sn_layer {       // one spectrally-normalised weight [cout, f]
    const float* w; float* u; float* v; float* sigma /* [2] */; float* t /* [f] scratch */; float* s /* [cout] scratch */;
    int cout, f;
};
struct cb200_sn_pack_job {    // one packing job (see cb200_sn_pack_weights)
    const float* w; const float* sigma; float* fwd; float* dgrad; long long ld_fwd, ldt;
    int cout, cin, kh, kw, dgrad_mode, col0, round_out;
};
struct cb200_sn_bwd_job {     // one weight-gradient un-packing job (see cb200_sn_weight_bwd)
    const float* dw_hat_packed; const float* w; const float* u; const float* v; const float* sigma; float* acc; float* dw;
    long long ld_fwd; int cout, cin, kh, kw;
};
}

// Power iteration (or sigma from stored u, v when training == 0) for n <= 16 layers: 1 memset-free phase each.
// The caller zeroes the `t` scratch of every layer (one contiguous memset) before the call when training.
extern "C" int cb200_sn_power_iter_batched(const cb200_sn_layer* layers, int n, float eps, int training, void* stream) {
    CB200_CHECK_ARG(n > 0 && n <= kMaxLayers, "sn_power_iter_batched: 1..16 layers");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SnPowerBatch p;
    p.n = n;
    p.rows_per_cta = 16;
    for (int l = 0; l < n; ++l) {
        p.w[l] = layers[l].w; p.u[l] = layers[l].u; p.v[l] = layers[l].v; p.sigma[l] = layers[l].sigma;
        p.t[l] = layers[l].t; p.s[l] = layers[l].s; p.cout[l] = layers[l].cout; p.f[l] = layers[l].f;
    }
    if (training) {
        int total = 0;
        for (int l = 0; l < n; ++l) {
            p.cta_begin[l] = total;
            total += ((p.f[l] + kT * 4 - 1) / (kT * 4)) * ((p.cout[l] + p.rows_per_cta - 1) / p.rows_per_cta);
        }
        p.cta_begin[n] = total;
        sn_wtu_batched_kernel<<<total, kT, 0, st>>>(p);
        CB200_COUNT_LAUNCH();
    }
    int rows = 0;
    for (int l = 0; l < n; ++l) { p.cta_begin[l] = rows; rows += p.cout[l]; }
    p.cta_begin[n] = rows;
    sn_wv_batched_kernel<<<rows, kT, 0, st>>>(p, training);
    CB200_COUNT_LAUNCH();
    sn_final_batched_kernel<<<n, kT, 0, st>>>(p, eps, training);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sn_power_iter_batched");
    return CB200_OK;
}

extern "C" int cb200_sn_pack_batched(const cb200_sn_pack_job* jobs, int n, void* stream) {
    CB200_CHECK_ARG(n > 0 && n <= kMaxLayers, "sn_pack_batched: 1..16 jobs");
    const bool tiled = sn_tiled_enabled();
    SnPackBatch p;
    p.n = n;
    int total = 0;
    for (int l = 0; l < n; ++l) {
        const cb200_sn_pack_job& j = jobs[l];
        CB200_CHECK_ARG(j.dgrad == nullptr || (j.dgrad_mode >= 1 && j.dgrad_mode <= 3), "sn_pack_batched: bad dgrad_mode");
        p.w[l] = j.w; p.sigma[l] = j.sigma; p.fwd[l] = j.fwd; p.dg[l] = j.dgrad; p.ld_fwd[l] = j.ld_fwd; p.ldt[l] = j.ldt;
        p.cout[l] = j.cout; p.cin[l] = j.cin; p.kh[l] = j.kh; p.kw[l] = j.kw;
        p.dg_mode[l] = j.dgrad_mode; p.col0[l] = j.col0; p.do_round[l] = j.round_out;
        p.cta_begin[l] = total;
        if (tiled) {
            CB200_CHECK_ARG(j.kh * j.kw <= 16, "sn_pack_batched: at most 16 taps");
            total += tiles_of(j.cout, j.cin, j.kh * j.kw);
        } else {
            const long long elems = (long long)j.cout * j.cin * j.kh * j.kw;
            total += (int)((elems + kT * 4 - 1) / (kT * 4));
        }
    }
    p.cta_begin[n] = total;
    if (tiled) {
        cudaFuncSetAttribute(sn_pack_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes);
        sn_pack_tiled_kernel<<<total, kT, kTileSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
    } else {
        sn_pack_batched_kernel<<<total, kT, 0, static_cast<cudaStream_t>(stream)>>>(p);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sn_pack_batched");
    return CB200_OK;
}

// The caller zeroes every job's `acc` scalar (one contiguous memset) before the call.
extern "C" int cb200_sn_weight_bwd_batched(const cb200_sn_bwd_job* jobs, int n, void* stream) {
    CB200_CHECK_ARG(n > 0 && n <= kMaxLayers, "sn_weight_bwd_batched: 1..16 jobs");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SnBwdBatch p;
    p.n = n;
    int total = 0;
    if (sn_tiled_enabled()) {
        for (int l = 0; l < n; ++l) {
            const cb200_sn_bwd_job& j = jobs[l];
            CB200_CHECK_ARG(j.kh * j.kw <= 16, "sn_weight_bwd_batched: at most 16 taps");
            p.dwp[l] = j.dw_hat_packed; p.w[l] = j.w; p.u[l] = j.u; p.v[l] = j.v; p.sigma[l] = j.sigma; p.acc[l] = j.acc;
            p.dw[l] = j.dw; p.ld_fwd[l] = j.ld_fwd; p.cout[l] = j.cout; p.cin[l] = j.cin; p.kh[l] = j.kh; p.kw[l] = j.kw;
            p.cta_begin[l] = total;
            total += tiles_of(j.cout, j.cin, j.kh * j.kw);
        }
        p.cta_begin[n] = total;
        cudaFuncSetAttribute(sn_bwd_dot_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes);
        cudaFuncSetAttribute(sn_bwd_apply_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes);
        sn_bwd_dot_tiled_kernel<<<total, kT, kTileSmemBytes, st>>>(p);
        CB200_COUNT_LAUNCH();
        sn_bwd_apply_tiled_kernel<<<total, kT, kTileSmemBytes, st>>>(p);
        CB200_COUNT_LAUNCH();
        CB200_CHECK_LAUNCH("sn_weight_bwd_batched");
        return CB200_OK;
    }
    for (int l = 0; l < n; ++l) {
        const cb200_sn_bwd_job& j = jobs[l];
        p.dwp[l] = j.dw_hat_packed; p.w[l] = j.w; p.u[l] = j.u; p.v[l] = j.v; p.sigma[l] = j.sigma; p.acc[l] = j.acc;
        p.dw[l] = j.dw; p.ld_fwd[l] = j.ld_fwd; p.cout[l] = j.cout; p.cin[l] = j.cin; p.kh[l] = j.kh; p.kw[l] = j.kw;
        p.cta_begin[l] = total;
        const long long elems = (long long)j.cout * j.cin * j.kh * j.kw;
        long long c = (elems + kT * 8 - 1) / (kT * 8);
        if (c > 64) c = 64;
        total += (int)c;
    }
    p.cta_begin[n] = total;
    sn_bwd_dot_batched_kernel<<<total, kT, 0, st>>>(p);
    CB200_COUNT_LAUNCH();
    total = 0;
    for (int l = 0; l < n; ++l) {
        p.cta_begin[l] = total;
        const long long elems = (long long)p.cout[l] * p.cin[l] * p.kh[l] * p.kw[l];
        total += (int)((elems + kT - 1) / kT);
    }
    p.cta_begin[n] = total;
    sn_bwd_apply_batched_kernel<<<total, kT, 0, st>>>(p);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sn_weight_bwd_batched");
    return CB200_OK;
}
