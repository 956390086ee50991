// Generator-side elementwise / reduction kernels of G_SNDCGAN for sm_100a (HBM-bound SIMT):
// train-mode BatchNorm (+ReLU) forward / backward on NHWC activations, the final tanh stage, and a
// TF32 rounding pass.  The dense parts of the generator (Linear, 3x ConvTranspose2d(4,2,1),
// ConvTranspose2d(3,1,1)) run on the tcgen05 tap-GEMM / wgrad kernels (tc_gemm.cu, tc_wgrad.cu).
//
// Reference: models/gan/sndcgan.py:24-48 (nn.BatchNorm2d in train mode: batch statistics, biased variance
// for normalisation, running stats updated with momentum 0.1 and the unbiased variance; nn.ReLU; nn.Tanh;
// `0.5 * y + 0.5`).  Under DDP the reference converts BN to SyncBatchNorm (train_gan.py:268): the two-phase
// split here (partial sums -> finalize) lets the host all-reduce the [2, C] sums between the phases.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kT = 256;

// sums[0][c] += sum_m x[m,c];  sums[1][c] += sum_m x[m,c]^2     (x: [M, C] row-major)
// threads: (256/cc) row lanes x cc columns, cc = min(C, 256)
__global__ void __launch_bounds__(kT)
bn_stats_kernel(const float* __restrict__ x, int M, int C, int cc, int rows_per_cta, float* __restrict__ sums) {
    __shared__ float scratch[2 * 256];
    const int tx = threadIdx.x % cc, ty = threadIdx.x / cc, lanes = kT / cc;
    const int c = blockIdx.x * cc + tx;
    const int m0 = blockIdx.y * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
    float a[2] = {0.f, 0.f};
    if (c < C) {
        float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f, s3 = 0.f, q3 = 0.f;
        int m = m0 + ty;
        for (; m + 3 * lanes < m1; m += 4 * lanes) {          // 4 independent loads in flight
            const float v0 = __ldg(x + (long long)m * C + c);
            const float v1 = __ldg(x + (long long)(m + lanes) * C + c);
            const float v2 = __ldg(x + (long long)(m + 2 * lanes) * C + c);
            const float v3 = __ldg(x + (long long)(m + 3 * lanes) * C + c);
            a[0] += v0; a[1] += v0 * v0; s1 += v1; q1 += v1 * v1; s2 += v2; q2 += v2 * v2; s3 += v3; q3 += v3 * v3;
        }
        for (; m < m1; m += lanes) {
            const float v = __ldg(x + (long long)m * C + c);
            a[0] += v;
            a[1] += v * v;
        }
        a[0] += (s1 + s2) + s3;
        a[1] += (q1 + q2) + q3;
    }
    fold_row_lanes<2>(a, scratch, cc);
    if (ty == 0 && c < C) {
        atomicAdd(sums + c, a[0]);
        atomicAdd(sums + C + c, a[1]);
    }
}

// stats[0][c] = mean, stats[1][c] = rstd; running stats updated in place (momentum, unbiased variance)
__global__ void __launch_bounds__(kT)
bn_finalize_kernel(const float* __restrict__ sums, float count, int C, float eps, float momentum,
                   float* __restrict__ stats, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * kT + threadIdx.x;
    if (c >= C) return;
    const float mean = sums[c] / count;
    const float var = fmaxf(sums[C + c] / count - mean * mean, 0.f);
    stats[c] = mean;
    stats[C + c] = rsqrtf(var + eps);
    if (running_mean) {
        const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// y = relu(gamma * (x - mean) * rstd + beta), optionally TF32-rounded.
// remap_s > 0: x is [M, C*S] with feature index c*S + s (the reference's flattened (c,h,w) order, C = channels,
// S = h*w) and y is NHWC [M, S, C]; statistics are per FEATURE (BatchNorm2d over a 1x1 map, sndcgan.py:42-45).
__global__ void __launch_bounds__(kT)
bn_apply_relu_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ y, long long total, int C, int remap_s,
                     int round_out) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;      // output index
    if (i >= total) return;
    int f;            // statistics / affine index
    long long src;    // input index
    if (remap_s > 0) {
        const int F = C;                          // here C = number of features = channels * S
        const int ch = F / remap_s;
        const long long m = i / F;
        const int r = (int)(i - m * F);           // r = s * ch + c
        const int s = r / ch, c = r - s * ch;
        f = c * remap_s + s;
        src = m * F + f;
    } else {
        f = (int)(i % C);
        src = i;
    }
    float v = (x[src] - stats[f]) * stats[C + f] * gamma[f] + beta[f];
    v = fmaxf(v, 0.f);
    y[i] = round_out ? round_tf32(v) : v;
}

// float4 fast path of bn_apply_relu for the plain NHWC case (C % 4 == 0, no remap): 4 channels per thread.
__global__ void __launch_bounds__(kT)
bn_apply_relu_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float4* __restrict__ y, long long total4, int C, int round_out) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= total4) return;
    const int f = (int)((i * 4) % C);
    const float4 v = x[i];
    const float4 mu = *reinterpret_cast<const float4*>(stats + f);
    const float4 rs = *reinterpret_cast<const float4*>(stats + C + f);
    const float4 g = *reinterpret_cast<const float4*>(gamma + f);
    const float4 b = *reinterpret_cast<const float4*>(beta + f);
    float4 o;
    o.x = fmaxf((v.x - mu.x) * rs.x * g.x + b.x, 0.f); o.y = fmaxf((v.y - mu.y) * rs.y * g.y + b.y, 0.f);
    o.z = fmaxf((v.z - mu.z) * rs.z * g.z + b.z, 0.f); o.w = fmaxf((v.w - mu.w) * rs.w * g.w + b.w, 0.f);
    if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    y[i] = o;
}

// float4 fast path of bn_bwd_apply (C % 4 == 0, no remap)
__global__ void __launch_bounds__(kT)
bn_bwd_apply_vec4_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, const float4* __restrict__ x,
                         const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ sums,
                         float inv_count, float4* __restrict__ dx, long long total4, int C, int round_out) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= total4) return;
    const int f = (int)((i * 4) % C);
    const float4 gy = dy[i], yy = y[i], xx = x[i];
    const float4 mu = *reinterpret_cast<const float4*>(stats + f);
    const float4 rs = *reinterpret_cast<const float4*>(stats + C + f);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + f);
    const float4 s0 = *reinterpret_cast<const float4*>(sums + f);
    const float4 s1 = *reinterpret_cast<const float4*>(sums + C + f);
    float4 o;
    o.x = gm.x * rs.x * ((yy.x > 0.f ? gy.x : 0.f) - s0.x * inv_count - (xx.x - mu.x) * rs.x * s1.x * inv_count);
    o.y = gm.y * rs.y * ((yy.y > 0.f ? gy.y : 0.f) - s0.y * inv_count - (xx.y - mu.y) * rs.y * s1.y * inv_count);
    o.z = gm.z * rs.z * ((yy.z > 0.f ? gy.z : 0.f) - s0.z * inv_count - (xx.z - mu.z) * rs.z * s1.z * inv_count);
    o.w = gm.w * rs.w * ((yy.w > 0.f ? gy.w : 0.f) - s0.w * inv_count - (xx.w - mu.w) * rs.w * s1.w * inv_count);
    if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    dx[i] = o;
}

// sums[0][f] += sum dz, sums[1][f] += sum dz * xhat, with dz = dy * 1[y > 0]   (same indexing as above)
__global__ void __launch_bounds__(kT)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                     const float* __restrict__ stats, int M, int C, int remap_s, int cc, int rows_per_cta,
                     float* __restrict__ sums) {
    __shared__ float scratch[2 * 256];
    const int tx = threadIdx.x % cc, ty = threadIdx.x / cc, lanes = kT / cc;
    const int j = blockIdx.x * cc + tx;          // column of the OUTPUT-side tensors (dy, y)
    int f = j;
    if (remap_s > 0 && j < C) {
        const int ch = C / remap_s;
        const int s = j / ch, c = j - s * ch;
        f = c * remap_s + s;
    }
    float a[2] = {0.f, 0.f};
    if (j < C) {
        const float mean = stats[f], rstd = stats[C + f];
        const int m0 = blockIdx.y * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
        float b0 = 0.f, b1 = 0.f;
        int m = m0 + ty;
        for (; m + lanes < m1; m += 2 * lanes) {               // 2 x 3 independent loads in flight
            const long long r0 = (long long)m * C, r1 = (long long)(m + lanes) * C;
            const float y0 = __ldg(y + r0 + j), y1 = __ldg(y + r1 + j);
            const float d0 = __ldg(dy + r0 + j), d1 = __ldg(dy + r1 + j);
            const float x0 = __ldg(x + r0 + f), x1 = __ldg(x + r1 + f);
            const float g0 = y0 > 0.f ? d0 : 0.f, g1 = y1 > 0.f ? d1 : 0.f;
            a[0] += g0; a[1] += g0 * ((x0 - mean) * rstd);
            b0 += g1; b1 += g1 * ((x1 - mean) * rstd);
        }
        for (; m < m1; m += lanes) {
            const float g = (__ldg(y + (long long)m * C + j) > 0.f) ? __ldg(dy + (long long)m * C + j) : 0.f;
            const float xh = (__ldg(x + (long long)m * C + f) - mean) * rstd;
            a[0] += g;
            a[1] += g * xh;
        }
        a[0] += b0;
        a[1] += b1;
    }
    fold_row_lanes<2>(a, scratch, cc);
    if (ty == 0 && j < C) {
        atomicAdd(sums + f, a[0]);
        atomicAdd(sums + C + f, a[1]);
    }
}

// dx = gamma * rstd * (dz - sum_dz / count - xhat * sum_dz_xhat / count)    (dx in the INPUT-side layout)
__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                    const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ sums,
                    float count, float* __restrict__ dx, long long total, int C, int remap_s, int round_out) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;      // output-side index
    if (i >= total) return;
    int f;
    long long src;
    if (remap_s > 0) {
        const int ch = C / remap_s;
        const long long m = i / C;
        const int r = (int)(i - m * C);
        const int s = r / ch, c = r - s * ch;
        f = c * remap_s + s;
        src = m * C + f;
    } else {
        f = (int)(i % C);
        src = i;
    }
    const float rstd = stats[C + f];
    const float xh = (x[src] - stats[f]) * rstd;
    const float g = (y[i] > 0.f) ? dy[i] : 0.f;
    const float inv = 1.f / count;
    float v = gamma[f] * rstd * (g - sums[f] * inv - xh * sums[C + f] * inv);
    dx[src] = round_out ? round_tf32(v) : v;
}

// Row-tiled variants of the two remapped kernels (the generator's first BatchNorm: features in the reference's (c,h,w)
// order on the input side, NHWC on the output side).  One CTA moves one sample row through shared memory so that BOTH
// layouts are read / written along their contiguous index; the element-indexed kernels above read the input side with a
// stride of S floats (0.8 - 1.1 TB/s in the round-1 launch list).  tile[c * (S + 1) + s]: the odd stride keeps the
// (s-major) and the (c-major) accesses conflict-free.  Dynamic shared memory: (F / S) * (S + 1) floats.
__global__ void __launch_bounds__(kT)
bn_apply_relu_rows_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ y, int F, int S, int round_out) {
    extern __shared__ float tile[];
    const int ch = F / S;
    const float* xr = x + (long long)blockIdx.x * F;
    float* yr = y + (long long)blockIdx.x * F;
    for (int f = threadIdx.x; f < F; f += kT) {                       // input side: f = c * S + s
        float v = (xr[f] - stats[f]) * stats[F + f] * gamma[f] + beta[f];
        v = fmaxf(v, 0.f);
        const int c = f / S, sp = f - c * S;
        tile[c * (S + 1) + sp] = round_out ? round_tf32(v) : v;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < F; r += kT) {                       // output side: r = s * ch + c
        const int sp = r / ch, c = r - sp * ch;
        yr[r] = tile[c * (S + 1) + sp];
    }
}

__global__ void __launch_bounds__(kT)
bn_bwd_apply_rows_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                         const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ sums,
                         float count, float* __restrict__ dx, int F, int S, int round_out) {
    extern __shared__ float tile[];
    const int ch = F / S;
    const long long row = (long long)blockIdx.x * F;
    for (int r = threadIdx.x; r < F; r += kT) {                       // output side: masked gradient
        const int sp = r / ch, c = r - sp * ch;
        tile[c * (S + 1) + sp] = (y[row + r] > 0.f) ? dy[row + r] : 0.f;
    }
    __syncthreads();
    const float inv = 1.f / count;
    for (int f = threadIdx.x; f < F; f += kT) {                       // input side
        const int c = f / S, sp = f - c * S;
        const float rstd = stats[F + f];
        const float xh = (x[row + f] - stats[f]) * rstd;
        const float v = gamma[f] * rstd * (tile[c * (S + 1) + sp] - sums[f] * inv - xh * sums[F + f] * inv);
        dx[row + f] = round_out ? round_tf32(v) : v;
    }
}

// out[n,c,h,w] = 0.5 * tanh(pre[n,h,w,c] + bias[c]) + 0.5     (pre: NHWC with `cpad` channels, c < 3)
__global__ void __launch_bounds__(kT)
g_final_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ bias, float* __restrict__ out, int HW,
                   int cpad, long long total) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;   // NCHW index over B*3*HW
    if (i >= total) return;
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % 3);
    const long long b = i / (3LL * HW);
    const float v = __ldg(pre + (b * HW + p) * cpad + c) + (bias ? __ldg(bias + c) : 0.f);
    out[i] = 0.5f * tanhf(v) + 0.5f;
}

// dpre[n,c,h,w] = dout * 0.5 * (1 - t^2), t = 2*out - 1;  dbias[c] += sum dpre
// Grid-stride: a CTA folds many elements before its three atomics (one element per thread meant 3 same-address
// atomics per 256 elements - 18 k serialised atomics at b512, 35 us for a 6 MB tensor).
__global__ void __launch_bounds__(kT)
g_final_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dpre,
                   float* __restrict__ dbias, int HW, long long total) {
    __shared__ float red[3 * 32];
    float s[3] = {0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
        const float t = 2.f * out[i] - 1.f;
        const float g = dout[i] * 0.5f * (1.f - t * t);
        dpre[i] = g;
        const int c = (int)((i / HW) % 3);
        s[0] += c == 0 ? g : 0.f; s[1] += c == 1 ? g : 0.f; s[2] += c == 2 ? g : 0.f;
    }
    block_sum<3>(s, red);
    if (threadIdx.x == 0 && dbias) { atomicAdd(dbias + 0, s[0]); atomicAdd(dbias + 1, s[1]); atomicAdd(dbias + 2, s[2]); }
}

__global__ void __launch_bounds__(kT)
round_tf32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) y[i] = round_tf32(x[i]);
}

// float4 grid-stride variant (n % 4 == 0, 16-byte aligned): 4 independent 16-byte loads in flight per thread
__global__ void __launch_bounds__(kT)
round_tf32_vec_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n4; i += (long long)gridDim.x * kT) {
        float4 v = x[i];
        v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
        y[i] = v;
    }
}


// Error-compensated TF32 operands ("3xTF32" without touching the GEMM kernels): x = hi + lo with hi = rn_tf32(x),
// lo = rn_tf32(x - hi) (x - hi is exact in fp32).  A product (a_hi + a_lo)(b_hi + b_lo) ~ a_hi b_hi + a_lo b_hi + a_hi b_lo
// is a plain GEMM / convolution over a CONCATENATED reduction axis, so the split operands are written in the layouts
// the existing kernels reduce over:
//   mode 0: out[rows, 3C] = [hi | lo | hi]      (channel concat: pairs with weights [w_hi | w_hi | w_lo] along Cin)
//   mode 1: out[rows, 2C] = [hi | hi]           (pairs with weights [w_hi | w_lo]: only the weight is compensated)
//   mode 2: out[2, rows, C] = hi rows, lo rows  (batch concat for weight-gradient GEMMs, paired with [g ; g])
__global__ void __launch_bounds__(kT)
split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ out, long long rows, int c4, int mode) {
    const long long n4 = rows * c4;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n4; i += (long long)gridDim.x * kT) {
        const float4 v = x[i];
        float4 hi, lo;
        hi.x = round_tf32(v.x); hi.y = round_tf32(v.y); hi.z = round_tf32(v.z); hi.w = round_tf32(v.w);
        lo.x = round_tf32(v.x - hi.x); lo.y = round_tf32(v.y - hi.y); lo.z = round_tf32(v.z - hi.z); lo.w = round_tf32(v.w - hi.w);
        const long long r = i / c4;
        const int c = (int)(i - r * c4);
        if (mode == 0) {
            float4* o = out + r * 3 * c4 + c;
            o[0] = hi; o[c4] = lo; o[2 * c4] = hi;
        } else if (mode == 1) {
            float4* o = out + r * 2 * c4 + c;
            o[0] = hi; o[c4] = hi;
        } else {
            out[i] = hi; out[n4 + i] = lo;
        }
    }
}

// Row chunks of the column-reduction grids: about 128 rows per chunk, but never fewer CTAs than ~8 per SM over the whole
// (column blocks x chunks) grid - [512, 8192] used to run on 128 CTAs, [32768, 256] on 256 (0.8 - 2 TB/s, ncu launch
// list of round 1) - and at least 4 rows per row lane so that the unrolled loop body is used.
int row_chunks(int M, int col_blocks, int lanes, int* rows_per_cta) {
    int chunks = (M + 127) / 128;
    const int want = (148 * 8 + col_blocks - 1) / col_blocks;
    if (chunks < want) chunks = want;
    const int most = M / (4 * lanes);
    if (chunks > most) chunks = most;
    if (chunks > 1024) chunks = 1024;
    if (chunks < 1) chunks = 1;
    *rows_per_cta = (M + chunks - 1) / chunks;
    return (M + *rows_per_cta - 1) / *rows_per_cta;
}

constexpr size_t kRowsSmemMax = 96 * 1024;       // row-tiled remap kernels: one sample row (+ padding) per CTA

// CB200_BN_ROWS=0 selects the element-indexed remap kernels again (A/B measurements).
bool rows_kernels_enabled() {
    static const bool on = []() { const char* e = getenv("CB200_BN_ROWS"); return !(e && e[0] == '0'); }();
    return on;
}

int col_lanes(int C) {          // columns per CTA: smallest power of two >= C, between 32 and 256
    int cc = 256;
    while (cc > 32 && cc / 2 >= C) cc /= 2;
    return cc;
}

}  // namespace

// sums[2,C] (zeroed here) <- per-channel sum and sum of squares of x[M,C].
extern "C" int cb200_bn_stats(const float* x, int M, int C, float* sums, void* stream) {
    CB200_CHECK_ARG(M > 0 && C > 0, "bn_stats: empty input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, st);
    if (e != cudaSuccess) { cb200_set_error("bn_stats: memset: %s", cudaGetErrorString(e)); return (int)e; }
    int rpc;
    const int cc = col_lanes(C);
    const int chunks = row_chunks(M, (C + cc - 1) / cc, kT / cc, &rpc);
    bn_stats_kernel<<<dim3((C + cc - 1) / cc, chunks), kT, 0, st>>>(x, M, C, cc, rpc, sums);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("bn_stats");
    return CB200_OK;
}

// stats[2,C] <- {mean, rstd} from (possibly all-reduced) sums over `count` samples; running stats updated when given.
extern "C" int cb200_bn_finalize(const float* sums, float count, int C, float eps, float momentum, float* stats,
                                 float* running_mean, float* running_var, void* stream) {
    CB200_CHECK_ARG(C > 0 && count > 0.f, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<(C + kT - 1) / kT, kT, 0, static_cast<cudaStream_t>(stream)>>>(sums, count, C, eps, momentum, stats,
                                                                                        running_mean, running_var);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("bn_finalize");
    return CB200_OK;
}

extern "C" int cb200_bn_apply_relu(const float* x, const float* stats, const float* gamma, const float* beta, float* y,
                                   int M, int C, int remap_s, int round_out, void* stream) {
    CB200_CHECK_ARG(M > 0 && C > 0 && (remap_s == 0 || C % remap_s == 0), "bn_apply_relu: bad shape");
    const long long total = (long long)M * C;
    const bool vec = remap_s == 0 && C % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(stats) |
                       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
    const size_t rows_smem = remap_s > 0 ? (size_t)(C / remap_s) * (remap_s + 1) * sizeof(float) : 0;
    if (vec)
        bn_apply_relu_vec4_kernel<<<(unsigned)((total / 4 + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(x), stats, gamma, beta, reinterpret_cast<float4*>(y), total / 4, C, round_out);
    else if (remap_s > 1 && rows_smem <= kRowsSmemMax && rows_kernels_enabled()) {
        cudaFuncSetAttribute(bn_apply_relu_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmemMax);
        bn_apply_relu_rows_kernel<<<M, kT, rows_smem, static_cast<cudaStream_t>(stream)>>>(x, stats, gamma, beta, y, C, remap_s,
                                                                                          round_out);
    } else
        bn_apply_relu_kernel<<<(unsigned)((total + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(
            x, stats, gamma, beta, y, total, C, remap_s, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("bn_apply_relu");
    return CB200_OK;
}

// sums[2,C] (zeroed here) <- {sum dz, sum dz*xhat}
extern "C" int cb200_bn_bwd_reduce(const float* dy, const float* y, const float* x, const float* stats, int M, int C,
                                   int remap_s, float* sums, void* stream) {
    CB200_CHECK_ARG(M > 0 && C > 0 && (remap_s == 0 || C % remap_s == 0), "bn_bwd_reduce: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, st);
    if (e != cudaSuccess) { cb200_set_error("bn_bwd_reduce: memset: %s", cudaGetErrorString(e)); return (int)e; }
    int rpc;
    const int cc = col_lanes(C);
    const int chunks = row_chunks(M, (C + cc - 1) / cc, kT / cc, &rpc);
    bn_bwd_reduce_kernel<<<dim3((C + cc - 1) / cc, chunks), kT, 0, st>>>(dy, y, x, stats, M, C, remap_s, cc, rpc, sums);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("bn_bwd_reduce");
    return CB200_OK;
}

extern "C" int cb200_bn_bwd_apply(const float* dy, const float* y, const float* x, const float* stats, const float* gamma,
                                  const float* sums, float count, float* dx, int M, int C, int remap_s, int round_out,
                                  void* stream) {
    CB200_CHECK_ARG(M > 0 && C > 0 && count > 0.f, "bn_bwd_apply: bad shape");
    const long long total = (long long)M * C;
    const bool vec = remap_s == 0 && C % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) |
                       reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(stats) | reinterpret_cast<uintptr_t>(gamma) |
                       reinterpret_cast<uintptr_t>(sums)) & 15) == 0;
    const size_t rows_smem = remap_s > 0 ? (size_t)(C / remap_s) * (remap_s + 1) * sizeof(float) : 0;
    if (vec)
        bn_bwd_apply_vec4_kernel<<<(unsigned)((total / 4 + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(x),
            stats, gamma, sums, 1.f / count, reinterpret_cast<float4*>(dx), total / 4, C, round_out);
    else if (remap_s > 1 && rows_smem <= kRowsSmemMax && rows_kernels_enabled()) {
        cudaFuncSetAttribute(bn_bwd_apply_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmemMax);
        bn_bwd_apply_rows_kernel<<<M, kT, rows_smem, static_cast<cudaStream_t>(stream)>>>(dy, y, x, stats, gamma, sums, count, dx,
                                                                                         C, remap_s, round_out);
    } else
        bn_bwd_apply_kernel<<<(unsigned)((total + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(
            dy, y, x, stats, gamma, sums, count, dx, total, C, remap_s, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("bn_bwd_apply");
    return CB200_OK;
}

extern "C" int cb200_g_final_fwd(const float* pre, const float* bias, float* out, int B, int H, int W, int cpad,
                                 void* stream) {
    CB200_CHECK_ARG(B > 0 && cpad >= 3, "g_final_fwd: bad shape");
    const long long total = (long long)B * 3 * H * W;
    g_final_fwd_kernel<<<(unsigned)((total + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(pre, bias, out, H * W,
                                                                                                       cpad, total);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("g_final_fwd");
    return CB200_OK;
}

// dpre[B,3,H,W], dbias[3] (zeroed here) from dout and the saved output.
extern "C" int cb200_g_final_bwd(const float* dout, const float* out, float* dpre, float* dbias, int B, int H, int W,
                                 void* stream) {
    CB200_CHECK_ARG(B > 0, "g_final_bwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dbias) {
        cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * 3, st);
        if (e != cudaSuccess) { cb200_set_error("g_final_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const long long total = (long long)B * 3 * H * W;
    long long grid = (total + kT - 1) / kT;
    if (grid > 148 * 4) grid = 148 * 4;
    g_final_bwd_kernel<<<(unsigned)grid, kT, 0, st>>>(dout, out, dpre, dbias, H * W, total);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("g_final_bwd");
    return CB200_OK;
}

extern "C" int cb200_round_tf32(const float* x, float* y, long long n, void* stream) {
    CB200_CHECK_ARG(n > 0, "round_tf32: empty input");
    if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        const long long n4 = n / 4;
        long long grid = (n4 + kT - 1) / kT;
        if (grid > 148 * 16) grid = 148 * 16;
        round_tf32_vec_kernel<<<(unsigned)grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4);
    } else {
        round_tf32_kernel<<<(unsigned)((n + kT - 1) / kT), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("round_tf32");
    return CB200_OK;
}

// x[rows, C] (C % 4 == 0, 16-byte aligned) -> split TF32 operands, see split_tf32_kernel (mode 0: [rows, 3C] hi|lo|hi,
// mode 1: [rows, 2C] hi|hi, mode 2: [2, rows, C] hi ; lo).
extern "C" int cb200_split_tf32(const float* x, float* out, long long rows, int C, int mode, void* stream) {
    CB200_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && mode >= 0 && mode <= 2, "split_tf32: rows=%lld C=%d mode=%d", rows, C, mode);
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "split_tf32: alignment");
    const long long n4 = rows * (C / 4);
    long long grid = (n4 + kT - 1) / kT;
    if (grid > 148 * 16) grid = 148 * 16;
    split_tf32_kernel<<<(unsigned)grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), rows, C / 4, mode);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("split_tf32");
    return CB200_OK;
}
