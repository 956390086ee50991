// Contrastive + GAN losses of ContraD for sm_100a (SIMT fp32, warp-shuffle reductions):
//
//   rownorm      F.normalize(x, dim=1, eps=1e-12)                      (training/gan/contrad.py:43,48)
//   contrastive  NT-Xent  (training/criterion.py:24-45)  and  supcon-fake (training/gan/contrad.py:8-32),
//                flash-style: the [2N,2N] / [3N,3N] similarity matrix, the -5e4 diagonal fill, the
//                log-softmax, the positive mask of the reference (25-45 ATen ops, two R x R fp32 matrices)
//                are never materialised - one pass computes row log-sum-exps and the loss, the
//                backward pass recomputes the similarities and accumulates dZ directly.
//                supcon only evaluates the N fake rows the loss keeps (the reference computes all 3N).
//   gan_loss     nonsat / hinge / wgan / lsgan D and G losses + their gradients (contrad.py:52-64,75-80)
//   colsum       bias gradients
//
// Similarities are computed in full fp32 (FFMA) on purpose: the logits are sim / tau with tau = 0.1,
// so TF32 rounding of the operands would be amplified 10x in the softmax.  At
// N = 512 the two losses are 0.27 + 0.20 GFLOP forward - ~1e-4 of the step's FLOPs.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kD = 128;            // embedding width (d_project)
constexpr int kWarps = 8;
constexpr int kRowsPerWarp = 4;     // rows per warp: each broadcast of a row vector feeds 4 FMAs per lane; the dot phase is
                                    // shared-memory bound (4 + R wavefronts per 4 R FMAs and k-step), R = 2 -> 4 cuts that by a third
constexpr int kRowsPerCta = kWarps * kRowsPerWarp;   // 32
constexpr int kJT = 32;            // columns per tile = one per lane
constexpr int kZStride = kD + 4;   // padded smem row (float4-aligned, conflict-free 128-bit reads)
constexpr int kMaxSplits = 16;

// ------------------------------------------------------------------------------------------ rownorm
__global__ void __launch_bounds__(256)
rownorm_fwd_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ y, float* __restrict__ inv_norm,
                   int rows, int d, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * ldx;
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) { float v = xr[k]; ss += v * v; }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    for (int k = lane; k < d; k += 32) y[(long long)row * d + k] = xr[k] * inv;
    if (lane == 0) inv_norm[row] = inv;
}

// dx = inv * (dy - y * <dy, y>)
__global__ void __launch_bounds__(256)
rownorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                   float* __restrict__ dx, long long lddx, int rows, int d, int round_out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* gr = dy + (long long)row * d;
    const float* yr = y + (long long)row * d;
    float dot = 0.f;
    for (int k = lane; k < d; k += 32) dot += gr[k] * yr[k];
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    for (int k = lane; k < d; k += 32) {
        float v = inv * (gr[k] - yr[k] * dot);
        dx[(long long)row * lddx + k] = round_out ? round_tf32(v) : v;
    }
}

// ------------------------------------------------------------------------------------------ contrastive
// mode 0: NT-Xent over R = 2N rows (all rows are loss rows);  mode 1: supcon-fake over R = 3N rows, loss rows
// 2N..3N-1.  Grid = (blocks of 16 rows, column splits): every CTA streams its slice of the columns through a
// shared-memory tile (lane-owns-column dot products against the warp's rows) so that ~2 x #SM CTAs are
// resident even though the problem has only 1-1.5 k rows.

__device__ __forceinline__ void load_ztile(float (*zt)[kZStride], const float* __restrict__ z, int j0, int R) {
    for (int e = threadIdx.x; e < kJT * (kD / 4); e += kWarps * 32) {
        const int r = e / (kD / 4), c4 = e % (kD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + r < R) v = __ldg(reinterpret_cast<const float4*>(z + (long long)(j0 + r) * kD) + c4);
        *reinterpret_cast<float4*>(&zt[r][c4 * 4]) = v;
    }
}

// s[r] = <z_i[r], z_j(lane)> for the warp's rows
__device__ __forceinline__ void dots_rows(const float (*zi)[kD], const float (*zt)[kZStride], int lane,
                                          float (&s)[kRowsPerWarp]) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) s[r] = 0.f;
#pragma unroll 8
    for (int k4 = 0; k4 < kD / 4; ++k4) {
        const float4 b = *reinterpret_cast<const float4*>(&zt[lane][k4 * 4]);
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&zi[r][k4 * 4]);
            s[r] = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s[r]))));
        }
    }
}

// partial[(local_row * splits + split) * 3 + {0,1,2}] = running max, sum of exp, positive-logit sum
__global__ void __launch_bounds__(kWarps * 32)
contrastive_fwd_kernel(const float* __restrict__ z, int R, int N, int mode, float inv_tau, int cols_per_split,
                       float* __restrict__ partial) {
    __shared__ __align__(16) float zt[kJT][kZStride];
    __shared__ __align__(16) float zi_all[kWarps][kRowsPerWarp][kD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = (mode == 1 ? 2 * N : 0);
    const int row0 = first + blockIdx.x * kRowsPerCta + warp * kRowsPerWarp;
    const int splits = gridDim.y, split = blockIdx.y;
    const int jbeg = split * cols_per_split, jend = min(R, jbeg + cols_per_split);
    float (*zi)[kD] = zi_all[warp];
    int gi[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        gi[r] = row0 + r;
        for (int k = lane; k < kD; k += 32) zi[r][k] = (gi[r] < R) ? __ldg(z + (long long)gi[r] * kD + k) : 0.f;
    }
    float m[kRowsPerWarp], l[kRowsPerWarp], pos[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) { m[r] = -INFINITY; l[r] = 0.f; pos[r] = 0.f; }
    for (int j0 = jbeg; j0 < jend; j0 += kJT) {
        __syncthreads();
        load_ztile(zt, z, j0, jend);
        __syncthreads();
        float s[kRowsPerWarp];
        dots_rows(zi, zt, lane, s);
        const int j = j0 + lane;
        const bool in = j < jend;
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            float v = s[r] * inv_tau;
            if (j == gi[r]) v = -5e4f;
            if (mode == 0) {
                const int pj = gi[r] < N ? gi[r] + N : gi[r] - N;
                if (in && j == pj) pos[r] += v;
            } else {
                if (in && j >= 2 * N && j != gi[r]) pos[r] += v;
            }
            const float tile_max = warp_max(in ? v : -INFINITY);
            const float new_m = fmaxf(m[r], tile_max);
            float e = in ? __expf(v - new_m) : 0.f;
            e = warp_sum(e);
            l[r] = l[r] * __expf(m[r] - new_m) + e;
            m[r] = new_m;
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        const float p = warp_sum(pos[r]);
        if (lane == 0 && gi[r] < R) {
            float* dst = partial + ((long long)(gi[r] - first) * splits + split) * 3;
            dst[0] = m[r]; dst[1] = l[r]; dst[2] = p;
        }
    }
}

// merge the column splits: lse[row], loss[0] (single CTA, fixed order => deterministic)
__global__ void __launch_bounds__(256)
contrastive_finalize_kernel(const float* __restrict__ partial, int rows, int splits, int N, int mode,
                            float* __restrict__ lse_out, float* __restrict__ loss) {
    __shared__ float red[32];
    float a[1] = {0.f};
    for (int row = threadIdx.x; row < rows; row += 256) {
        const float* p = partial + (long long)row * splits * 3;
        float M = -INFINITY;
        for (int s = 0; s < splits; ++s) M = fmaxf(M, p[s * 3]);
        float L = 0.f, pos = 0.f;
        for (int s = 0; s < splits; ++s) { L += p[s * 3 + 1] * __expf(p[s * 3] - M); pos += p[s * 3 + 2]; }
        const float lse = M + logf(L);
        lse_out[row] = lse;
        a[0] += (mode == 0) ? -(pos - lse) / (float)(2 * N) : -(pos / (float)(N - 1) - lse) / (float)N;
    }
    block_sum<1>(a, red);
    if (threadIdx.x == 0) loss[0] = a[0];
}

// dZ[i] += gscale * inv_tau * sum_{j in split} (c_ij + c_ji) z_j,   c_ij = dL/dS_ij
__global__ void __launch_bounds__(kWarps * 32)
contrastive_bwd_kernel(const float* __restrict__ z, int R, int N, int mode, float inv_tau, int cols_per_split,
                       const float* __restrict__ lse, const float* __restrict__ gscale, float* __restrict__ dz) {
    __shared__ __align__(16) float zt[kJT][kZStride];
    __shared__ __align__(16) float zi_all[kWarps][kRowsPerWarp][kD];
    __shared__ float coef_all[kWarps][kRowsPerWarp][kJT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kRowsPerCta + warp * kRowsPerWarp;     // all R rows receive gradient
    const int jbeg = blockIdx.y * cols_per_split, jend = min(R, jbeg + cols_per_split);
    float (*zi)[kD] = zi_all[warp];
    float (*coef)[kJT] = coef_all[warp];
    const int first_active = (mode == 1) ? 2 * N : 0;
    int gi[kRowsPerWarp];
    float lse_i[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        gi[r] = row0 + r;
        for (int k = lane; k < kD; k += 32) zi[r][k] = (gi[r] < R) ? __ldg(z + (long long)gi[r] * kD + k) : 0.f;
        lse_i[r] = (gi[r] < R && gi[r] >= first_active) ? __ldg(lse + gi[r] - first_active) : 0.f;
    }
    const float wpos = (mode == 0) ? 1.f / (float)(2 * N) : 1.f / (float)N;     // row weight
    const float wsup = 1.f / (float)(N - 1);
    float acc[kRowsPerWarp][4];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    for (int j0 = jbeg; j0 < jend; j0 += kJT) {
        __syncthreads();
        load_ztile(zt, z, j0, jend);
        __syncthreads();
        float s[kRowsPerWarp];
        dots_rows(zi, zt, lane, s);
        const int j = j0 + lane;
        const bool jin = j < jend;
        const bool j_active = jin && j >= first_active;
        const float lse_j = j_active ? __ldg(lse + j - first_active) : 0.f;
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int i = gi[r];
            float c = 0.f;
            if (jin && i < R && j != i) {
                const float v = s[r] * inv_tau;
                if (i >= first_active) {          // c_ij: row i is a loss row
                    float tgt;
                    if (mode == 0) tgt = (j == (i < N ? i + N : i - N)) ? 1.f : 0.f;
                    else tgt = (j >= 2 * N) ? wsup : 0.f;
                    c += wpos * (__expf(v - lse_i[r]) - tgt);
                }
                if (j_active) {                   // c_ji: row j is a loss row (S symmetric)
                    float tgt;
                    if (mode == 0) tgt = (i == (j < N ? j + N : j - N)) ? 1.f : 0.f;
                    else tgt = (i >= 2 * N) ? wsup : 0.f;
                    c += wpos * (__expf(v - lse_j) - tgt);
                }
            }
            coef[r][lane] = c;
        }
        __syncwarp();
        // acc[r][k-slice] += sum_j coef[r][j] * z_j[k-slice];  lane owns k = 4*lane .. 4*lane+3
#pragma unroll 4
        for (int jj = 0; jj < kJT; ++jj) {
            const float4 b = *reinterpret_cast<const float4*>(&zt[jj][lane * 4]);
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) {
                const float c = coef[r][jj];
                acc[r][0] += c * b.x; acc[r][1] += c * b.y; acc[r][2] += c * b.z; acc[r][3] += c * b.w;
            }
        }
        __syncwarp();
    }
    const float g = __ldg(gscale) * inv_tau;
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        if (gi[r] < R) {
            float* dst = dz + (long long)gi[r] * kD + lane * 4;
            if (gridDim.y == 1) {
                *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0] * g, acc[r][1] * g, acc[r][2] * g, acc[r][3] * g);
            } else {
                atomicAdd(dst + 0, acc[r][0] * g); atomicAdd(dst + 1, acc[r][1] * g);
                atomicAdd(dst + 2, acc[r][2] * g); atomicAdd(dst + 3, acc[r][3] * g);
            }
        }
    }
}

// pick the number of column splits so that ~2 CTAs per SM are in flight; columns per split multiple of 32
static int pick_splits(int row_blocks, int R, int* cols_per_split) {
    int splits = (2 * 148 + row_blocks - 1) / row_blocks;
    const int max_by_cols = (R + 4 * kJT - 1) / (4 * kJT);      // at least 4 tiles of columns per CTA
    if (splits > max_by_cols) splits = max_by_cols;
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
    int cps = (R + splits - 1) / splits;
    cps = ((cps + kJT - 1) / kJT) * kJT;
    *cols_per_split = cps;
    return (R + cps - 1) / cps;
}

// ------------------------------------------------------------------------------------------ GAN losses
// kind: 0 nonsat, 1 hinge, 2 wgan, 3 lsgan.   d_real / d_gen: [N] with element stride `stride`.
// out[0] = L_dis, out[1] = mean d_real, out[2] = mean d_gen;  grads (dL_dis/dd) written to g_real / g_gen.
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
gan_d_loss_kernel(const float* __restrict__ d_real, const float* __restrict__ d_gen, long long stride, int N,
                  int kind, float* __restrict__ out, float* __restrict__ g_real, float* __restrict__ g_gen) {
    __shared__ float red[4 * 32];
    float a[4] = {0.f, 0.f, 0.f, 0.f};     // loss_real, loss_gen, sum real, sum gen
    const float invN = 1.f / (float)N;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float r = d_real[i * stride], g = d_gen[i * stride];
        a[2] += r; a[3] += g;
        float gr, gg;
        if (kind == 0) { a[0] += softplus_f(-r); a[1] += softplus_f(g); gr = -sigmoid_f(-r); gg = sigmoid_f(g); }
        else if (kind == 1) { a[0] += fmaxf(1.f - r, 0.f); a[1] += fmaxf(1.f + g, 0.f); gr = (1.f - r > 0.f) ? -1.f : 0.f; gg = (1.f + g > 0.f) ? 1.f : 0.f; }
        else if (kind == 2) { a[0] += -r; a[1] += g; gr = -1.f; gg = 1.f; }
        else { a[0] += 0.5f * (r - 1.f) * (r - 1.f); a[1] += 0.5f * g * g; gr = (r - 1.f); gg = g; }
        g_real[i] = gr * invN;
        g_gen[i] = gg * invN;
    }
    block_sum<4>(a, red);
    if (threadIdx.x == 0) {
        out[0] = (a[0] + a[1]) * invN;
        out[1] = a[2] * invN;
        out[2] = a[3] * invN;
    }
}

// G loss: kind 0 nonsat softplus(-d).mean(); 3 lsgan 0.5*(d-1)^2.mean(); else -d.mean()
__global__ void __launch_bounds__(256)
gan_g_loss_kernel(const float* __restrict__ d_gen, long long stride, int N, int kind, float* __restrict__ out,
                  float* __restrict__ g_gen) {
    __shared__ float red[32];
    float a[1] = {0.f};
    const float invN = 1.f / (float)N;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float g = d_gen[i * stride];
        float gg;
        if (kind == 0) { a[0] += softplus_f(-g); gg = -sigmoid_f(-g); }
        else if (kind == 3) { a[0] += 0.5f * (g - 1.f) * (g - 1.f); gg = g - 1.f; }
        else { a[0] += -g; gg = -1.f; }
        g_gen[i] = gg * invN;
    }
    block_sum<1>(a, red);
    if (threadIdx.x == 0) out[0] = a[0] * invN;
}

// ------------------------------------------------------------------------------------------ column sums
// out[n] (+)= sum_m x[m, n]   x: [M, N] with row stride ld.  Threads: (256/cc) row lanes x cc column-slots;
// VEC = 4: every slot is a float4 (N and ld multiples of 4, 16-byte aligned base) -> 128-bit loads, 4 rows in flight.
template <int VEC>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long ld, int M, int N, int cc, int rows_per_cta, float* __restrict__ out) {
    __shared__ float scratch[VEC * 256];
    const int tx = threadIdx.x % cc, ty = threadIdx.x / cc, lanes = 256 / cc;
    const int n = (blockIdx.x * cc + tx) * VEC;
    const int m0 = blockIdx.y * rows_per_cta;
    const int m1 = min(M, m0 + rows_per_cta);
    float a[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) a[e] = 0.f;
    if (n < N) {
        float b1[VEC], b2[VEC], b3[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) { b1[e] = 0.f; b2[e] = 0.f; b3[e] = 0.f; }
        auto load = [&](int m, float (&acc)[VEC]) {
            if (VEC == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long long)m * ld + n));
                acc[0] += v.x; acc[1 % VEC] += v.y; acc[2 % VEC] += v.z; acc[3 % VEC] += v.w;
            } else {
                acc[0] += __ldg(x + (long long)m * ld + n);
            }
        };
        int m = m0 + ty;
        for (; m + 3 * lanes < m1; m += 4 * lanes) {          // 4 independent loads in flight
            load(m, a); load(m + lanes, b1); load(m + 2 * lanes, b2); load(m + 3 * lanes, b3);
        }
        for (; m < m1; m += lanes) load(m, a);
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] += (b1[e] + b2[e]) + b3[e];
    }
    // fold the row lanes: inside a warp by shuffles (cc < 32), across warps through shared memory (<= 8 terms)
    for (int o = 16; o >= cc; o >>= 1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) a[e] += __shfl_xor_sync(0xffffffffu, a[e], o);
    }
    const int span = cc < 32 ? 32 : cc, groups = 256 / span;      // one partial per (warp-group, column slot)
    if (groups > 1) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) scratch[e * 256 + threadIdx.x] = a[e];
        __syncthreads();
        if ((int)threadIdx.x < cc) {
            for (int r = 1; r < groups; ++r) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) a[e] += scratch[e * 256 + r * span + threadIdx.x];
            }
        }
    }
    if (ty == 0 && n < N) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) atomicAdd(out + n + e, a[e]);
    }
}

// out = dy * (act > 0 ? 1 : slope)   (LeakyReLU backward from the saved OUTPUT activation; slope > 0)
__global__ void __launch_bounds__(256)
lrelu_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ act, float4* __restrict__ out, long long n4,
                 float slope, int round_out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    float4 g = dy[i];
    const float4 a = act[i];
    g.x *= a.x > 0.f ? 1.f : slope; g.y *= a.y > 0.f ? 1.f : slope;
    g.z *= a.z > 0.f ? 1.f : slope; g.w *= a.w > 0.f ? 1.f : slope;
    if (round_out) { g.x = round_tf32(g.x); g.y = round_tf32(g.y); g.z = round_tf32(g.z); g.w = round_tf32(g.w); }
    out[i] = g;
}


// ------------------------------------------------------------------------------------------ tensor-core similarity path
// north_star's formulation of the two contrastive losses: the pairwise-cosine similarity matrix is ONE tcgen05 GEMM
// (cb200_gemm_nt_tf32 on error-compensated operands, see contrad_b200/functional.py ContrastiveTCFn), the
// temperature-softmax / cross-entropy below are warp-shuffle row reductions over that matrix.
//   S [Ra, lds]  dot products <z_i, z_j> of the loss rows i (global row index row0 + i) against all R rows (columns >= R
//                are padding), mode 0: NT-Xent (Ra = R = 2N), mode 1: supcon-fake (R = 3N, loss rows 2N .. 3N-1)
// fwd: lse[i] = logsumexp_j!=i(S_ij / tau);  rowloss[i] = -(c) * sum_j w_ij (S_ij / tau - lse_i)  with the reference's
//      weights (criterion.py:35-45: positive j = (i + N) mod 2N, c = 1/2N; contrad.py:13-32: the other fakes, 1/(N-1), c = 1/N)
// bwd: G_ij = gscale * c * (softmax_ij - w_ij) / tau, G_ii = 0   (d loss / d S_ij; padding columns zero)
__global__ void __launch_bounds__(256)
sim_rows_fwd_kernel(const float* __restrict__ S, long long lds, int Ra, int R, int N, int mode, int row0, float inv_tau,
                    float* __restrict__ lse, float* __restrict__ rowloss) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= Ra) return;
    const int gi = row0 + r;
    const float* s = S + (long long)r * lds;
    float m = -3.0e38f;
    for (int j = lane; j < R; j += 32) if (j != gi) m = fmaxf(m, s[j] * inv_tau);
    m = warp_max(m);
    float se = 0.f, pos = 0.f;
    const int pj = (gi + N) % (2 * N);
    for (int j = lane; j < R; j += 32) {
        if (j == gi) continue;
        const float l = s[j] * inv_tau;
        se += __expf(l - m);
        if (mode == 0) { if (j == pj) pos += l; }
        else if (j >= 2 * N) pos += l;
    }
    se = warp_sum(se);
    pos = warp_sum(pos);
    if (lane == 0) {
        const float L = m + logf(se);
        lse[r] = L;
        rowloss[r] = (mode == 0) ? -(pos - L) / (float)(2 * N) : -(pos / (float)(N - 1) - L) / (float)N;
    }
}

__global__ void __launch_bounds__(256)
sim_rows_bwd_kernel(const float* __restrict__ S, long long lds, int Ra, int R, int N, int mode, int row0, float inv_tau,
                    const float* __restrict__ lse, const float* __restrict__ gscale, float* __restrict__ G, long long ldg,
                    int cols) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= Ra) return;
    const int gi = row0 + r;
    const float* s = S + (long long)r * lds;
    float* g = G + (long long)r * ldg;
    const float L = lse[r];
    const float c = gscale[0] * inv_tau / (float)(mode == 0 ? 2 * N : N);
    const int pj = (gi + N) % (2 * N);
    const float wpos = (mode == 0) ? 1.f : 1.f / (float)(N - 1);
    for (int j = lane; j < cols; j += 32) {
        float v = 0.f;
        if (j < R && j != gi) {
            const float p = __expf(s[j] * inv_tau - L);
            const bool is_pos = (mode == 0) ? (j == pj) : (j >= 2 * N);
            v = c * (p - (is_pos ? wpos : 0.f));
        }
        g[j] = v;
    }
}

}  // namespace

extern "C" int cb200_lrelu_bwd(const float* dy, const float* act, float* out, long long n, float slope, int round_out,
                               void* stream) {
    CB200_CHECK_ARG(n > 0 && n % 4 == 0, "lrelu_bwd: element count must be a positive multiple of 4");
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(act) |
                      reinterpret_cast<uintptr_t>(out)) & 15) == 0, "lrelu_bwd: pointers must be 16-byte aligned");
    const long long n4 = n / 4;
    lrelu_bwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(act), reinterpret_cast<float4*>(out), n4,
        slope, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("lrelu_bwd");
    return CB200_OK;
}

extern "C" int cb200_rownorm_fwd(const float* x, long long ldx, float* y, float* inv_norm, int rows, int d, float eps,
                                 void* stream) {
    CB200_CHECK_ARG(rows > 0 && d > 0, "rownorm_fwd: empty input");
    rownorm_fwd_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, y, inv_norm, rows, d, eps);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("rownorm_fwd");
    return CB200_OK;
}

extern "C" int cb200_rownorm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, long long lddx,
                                 int rows, int d, int round_out, void* stream) {
    CB200_CHECK_ARG(rows > 0 && d > 0, "rownorm_bwd: empty input");
    rownorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, inv_norm, dx, lddx, rows, d,
                                                                                       round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("rownorm_bwd");
    return CB200_OK;
}

// z [R,128] L2-normalised rows; mode 0: NT-Xent (R = 2N); mode 1: supcon-fake (R = 3N, loss rows 2N..3N-1).
// lse: (mode ? N : 2N) floats (kept for the backward); scratch: 48 * (mode ? N : 2N) floats; loss[0] <- scalar.
extern "C" int cb200_contrastive_fwd(const float* z, int N, int d, int mode, float temperature, float* lse,
                                     float* scratch, float* loss, void* stream) {
    CB200_CHECK_ARG(N > 0 && (mode == 0 || mode == 1), "contrastive_fwd: bad N/mode");
    CB200_CHECK_ARG(d == kD, "contrastive_fwd: embedding width %d != 128", d);
    CB200_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0, "contrastive_fwd: z must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = mode ? 3 * N : 2 * N;
    const int active = mode ? N : 2 * N;
    const int row_blocks = (active + kRowsPerCta - 1) / kRowsPerCta;
    int cps;
    const int splits = pick_splits(row_blocks, R, &cps);
    contrastive_fwd_kernel<<<dim3(row_blocks, splits), kWarps * 32, 0, st>>>(z, R, N, mode, 1.f / temperature, cps, scratch);
    CB200_COUNT_LAUNCH();
    contrastive_finalize_kernel<<<1, 256, 0, st>>>(scratch, active, splits, N, mode, lse, loss);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("contrastive_fwd");
    return CB200_OK;
}

// dz [R,128] = gscale[0] * dLoss/dz  (gscale is a device scalar: the upstream gradient of the loss)
extern "C" int cb200_contrastive_bwd(const float* z, int N, int d, int mode, float temperature, const float* lse,
                                     const float* gscale, float* dz, void* stream) {
    CB200_CHECK_ARG(N > 0 && (mode == 0 || mode == 1), "contrastive_bwd: bad N/mode");
    CB200_CHECK_ARG(d == kD, "contrastive_bwd: embedding width %d != 128", d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = mode ? 3 * N : 2 * N;
    const int row_blocks = (R + kRowsPerCta - 1) / kRowsPerCta;
    int cps;
    const int splits = pick_splits(row_blocks, R, &cps);
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(dz, 0, sizeof(float) * (size_t)R * kD, st);
        if (e != cudaSuccess) { cb200_set_error("contrastive_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    contrastive_bwd_kernel<<<dim3(row_blocks, splits), kWarps * 32, 0, st>>>(z, R, N, mode, 1.f / temperature, cps, lse,
                                                                           gscale, dz);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("contrastive_bwd");
    return CB200_OK;
}

extern "C" int cb200_gan_d_loss(const float* d_real, const float* d_gen, long long stride, int N, int kind, float* out3,
                                float* g_real, float* g_gen, void* stream) {
    CB200_CHECK_ARG(N > 0 && kind >= 0 && kind <= 3, "gan_d_loss: bad N/kind");
    gan_d_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_real, d_gen, stride, N, kind, out3, g_real, g_gen);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("gan_d_loss");
    return CB200_OK;
}

extern "C" int cb200_gan_g_loss(const float* d_gen, long long stride, int N, int kind, float* out1, float* g_gen,
                                void* stream) {
    CB200_CHECK_ARG(N > 0 && kind >= 0 && kind <= 3, "gan_g_loss: bad N/kind");
    gan_g_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_gen, stride, N, kind, out1, g_gen);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("gan_g_loss");
    return CB200_OK;
}

// out[N] = column sums of x[M,N] (row stride ld); out is zeroed here.
extern "C" int cb200_colsum(const float* x, long long ld, int M, int N, float* out, void* stream) {
    CB200_CHECK_ARG(M > 0 && N > 0, "colsum: empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * N, st);
    if (e != cudaSuccess) { cb200_set_error("colsum: memset: %s", cudaGetErrorString(e)); return (int)e; }
    static const bool vec_ok = []() { const char* e = getenv("CB200_COLSUM_VEC"); return !(e && e[0] == '0'); }();
    const bool vec = vec_ok && (N % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int slots = vec ? N / 4 : N;
    int cc = 256;
    while (cc > 8 && cc / 2 >= slots) cc /= 2;              // smallest power of two >= #slots, between 8 and 256
    // rows per CTA: <= 512 for tall matrices; short-and-wide ones (heads, G linear) are cut finer so that the grid
    // still fills the machine (~4 CTAs per SM) instead of a handful of CTAs walking all rows serially.
    const int col_ctas = (slots + cc - 1) / cc;
    int row_chunks = (M + 511) / 512;
    const int want = (592 + col_ctas - 1) / col_ctas;
    if (row_chunks < want) row_chunks = want;
    const int min_rows = 256 / cc;                            // at least one row per row lane
    if (row_chunks > (M + min_rows - 1) / min_rows) row_chunks = (M + min_rows - 1) / min_rows;
    if (row_chunks > 2048) row_chunks = 2048;
    const int rows_per_cta = (M + row_chunks - 1) / row_chunks;
    dim3 grid(col_ctas, (M + rows_per_cta - 1) / rows_per_cta);
    if (vec) colsum_kernel<4><<<grid, 256, 0, st>>>(x, ld, M, N, cc, rows_per_cta, out);
    else colsum_kernel<1><<<grid, 256, 0, st>>>(x, ld, M, N, cc, rows_per_cta, out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("colsum");
    return CB200_OK;
}

// Row reductions of the tensor-core similarity path (see sim_rows_fwd_kernel): S [Ra, lds] dot products of the loss rows
// (global indices row0 ..) against R columns -> lse [Ra], rowloss [Ra] (their sum is the loss).
extern "C" int cb200_sim_rows_fwd(const float* S, long long lds, int Ra, int R, int N, int mode, int row0, float temperature,
                                  float* lse, float* rowloss, void* stream) {
    CB200_CHECK_ARG(Ra > 0 && R > 0 && N > 0 && (mode == 0 || mode == 1) && lds >= R, "sim_rows_fwd: bad shape");
    sim_rows_fwd_kernel<<<(Ra + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, lds, Ra, R, N, mode, row0,
                                                                                     1.f / temperature, lse, rowloss);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sim_rows_fwd");
    return CB200_OK;
}

// G [Ra, ldg] (columns 0 .. cols-1 written, zeros beyond R) = gscale[0] * d loss / d S.
extern "C" int cb200_sim_rows_bwd(const float* S, long long lds, int Ra, int R, int N, int mode, int row0, float temperature,
                                  const float* lse, const float* gscale, float* G, long long ldg, int cols, void* stream) {
    CB200_CHECK_ARG(Ra > 0 && R > 0 && N > 0 && (mode == 0 || mode == 1) && lds >= R && ldg >= cols && cols >= R,
                    "sim_rows_bwd: bad shape");
    sim_rows_bwd_kernel<<<(Ra + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, lds, Ra, R, N, mode, row0,
                                                                                     1.f / temperature, lse, gscale, G, ldg, cols);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("sim_rows_bwd");
    return CB200_OK;
}
