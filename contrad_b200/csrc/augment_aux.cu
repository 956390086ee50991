// Row f4: the light augmentations of the CR / bCR baselines that share the discriminators of the hot path.
//
//   shift_flip   HorizontalFlipRandomCrop ('hfrt') and RandomCrop (augment/spatial.py:14-67): per-sample horizontal
//                mirror (+-1) and translation, resampled by grid_sample(mode='nearest', padding_mode=..., align_corners=
//                False) on an affine grid.  The reference spends an affine_grid + grid_sample pair on what is an index
//                permutation: here each CTA builds the per-image source-row / source-column tables once in shared
//                memory and the body is gather -> store (forward) or load -> red.add (backward).
//   noise_clamp  Gaussian (augment/__init__.py:40-49): clamp(x + noise * sigma, 0, 1) and its gradient mask; the noise
//                itself stays torch.randn_like so that the Philox stream is the reference's.
//
//   diffaug      DiffAugment(policy = color / translation / cutout subsets, third_party/diffaug.py) of the `diffaug`
//                baselines (EXPERIMENTS.md: --mode=aug_both --aug=diffaug): brightness, saturation, contrast, integer
//                translation with zero fill and a square cutout, all per sample.  The chain is affine in x for fixed
//                draws, with one per-image reduction (the contrast mean; the mean of the gradient in the backward pass):
//                a reduction launch writes [B] scalars, a pointwise launch does the rest.
//
// All are HBM-bound: 8 B/element (shift_flip fwd), 12 B/element (noise fwd: x, noise in, y out), 12 B/element (diffaug:
// x read twice, y written).
#include "common.cuh"

namespace {

constexpr int kT = 256;

enum PadMode { kZeros = 0, kBorder = 1, kReflection = 2 };

// [host-testable: nearest_source]  (tests/test_host_logic.py compiles this function with g++ against the oracle)
// grid_sample's coordinate pipeline for one axis (ATen GridSampler.h: grid_sampler_unnormalize, clip_coordinates,
// reflect_coordinates for align_corners=False, then nearbyint).  Returns the source index, or -1 when the tap falls
// outside the image under padding_mode='zeros'.
__device__ __forceinline__ int nearest_source(float g, int size, int pad) {
    const float fs = (float)size;
    float c = ((g + 1.f) * fs - 1.f) * 0.5f;
    if (pad == kBorder) {
        c = fminf(fmaxf(c, 0.f), fs - 1.f);
    } else if (pad == kReflection) {
        // reflect about -0.5 and size-0.5 (twice_low = -1, twice_high = 2*size-1), then clip
        const float span = fs;
        float t = fabsf(c + 0.5f);
        const float extra = fmodf(t, span);
        const int flips = (int)floorf(t / span);
        c = (flips & 1) ? (span - extra - 0.5f) : (extra - 0.5f);
        c = fminf(fmaxf(c, 0.f), fs - 1.f);
    }
    const float r = nearbyintf(c);
    if (!(r >= 0.f && r <= fs - 1.f)) return -1;          // also catches NaN
    return (int)r;
}
// [host-testable: end]

// params [3, B]: sign (+-1, x axis), bias_x, bias_y (already divided by width/2 like the reference's r_bias).
__device__ __forceinline__ void build_tables(int* col, int* row, const float* __restrict__ params, int B, int b, int H,
                                             int W, int pad) {
    const float sign = __ldg(params + b), bx = __ldg(params + B + b), by = __ldg(params + 2 * B + b);
    for (int e = threadIdx.x; e < W + H; e += blockDim.x) {
        if (e < W) {
            const float base = (2.f * (float)e + 1.f) / (float)W - 1.f;      // affine_grid, align_corners=False
            col[e] = nearest_source(sign * base + bx, W, pad);
        } else {
            const int i = e - W;
            const float base = (2.f * (float)i + 1.f) / (float)H - 1.f;
            row[i] = nearest_source(base + by, H, pad);
        }
    }
}

// grid = (chunks, B); dynamic shared memory (W + H) ints.  P planes per image.
__global__ void __launch_bounds__(kT) shift_flip_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            const float* __restrict__ params, int B, int P, int H, int W,
                                                            int pad) {
    extern __shared__ int tables[];
    int* col = tables;
    int* row = tables + W;
    const int b = blockIdx.y;
    build_tables(col, row, params, B, b, H, W, pad);
    __syncthreads();
    const int per_img = P * H * W;
    const float* xb = x + (size_t)b * per_img;
    float* yb = y + (size_t)b * per_img;
    if ((W & 3) == 0) {
        const int W4 = W >> 2;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < per_img / 4; q += gridDim.x * blockDim.x) {
            const int j = (q % W4) * 4, i = (q / W4) % H, p = q / (W4 * H);
            const int r = row[i];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r >= 0) {
                const float* src = xb + ((size_t)p * H + r) * W;
                const int c0 = col[j], c1 = col[j + 1], c2 = col[j + 2], c3 = col[j + 3];
                if (c0 >= 0) v.x = __ldg(src + c0);
                if (c1 >= 0) v.y = __ldg(src + c1);
                if (c2 >= 0) v.z = __ldg(src + c2);
                if (c3 >= 0) v.w = __ldg(src + c3);
            }
            __stcs(reinterpret_cast<float4*>(yb) + q, v);
        }
    } else {
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per_img; e += gridDim.x * blockDim.x) {
            const int j = e % W, i = (e / W) % H, p = e / (W * H);
            const int r = row[i], c = col[j];
            yb[e] = (r >= 0 && c >= 0) ? __ldg(xb + ((size_t)p * H + r) * W + c) : 0.f;
        }
    }
}

// dx (zeroed by the caller of the kernel) += transpose of the gather: several outputs may read one source pixel under
// border / reflection padding, hence fp32 red.add.
__global__ void __launch_bounds__(kT) shift_flip_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                            const float* __restrict__ params, int B, int P, int H, int W,
                                                            int pad) {
    extern __shared__ int tables[];
    int* col = tables;
    int* row = tables + W;
    const int b = blockIdx.y;
    build_tables(col, row, params, B, b, H, W, pad);
    __syncthreads();
    const int per_img = P * H * W;
    const float* dyb = dy + (size_t)b * per_img;
    float* dxb = dx + (size_t)b * per_img;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per_img; e += gridDim.x * blockDim.x) {
        const int j = e % W, i = (e / W) % H, p = e / (W * H);
        const int r = row[i], c = col[j];
        if (r >= 0 && c >= 0) atomicAdd(dxb + ((size_t)p * H + r) * W + c, __ldcs(dyb + e));
    }
}

inline dim3 image_grid(int B, long long work_items_per_image) {
    long long blocks = (work_items_per_image + kT - 1) / kT;
    const long long cap = (148LL * 16 + B - 1) / B;               // about 16 resident CTAs per SM over the whole batch
    if (blocks > cap) blocks = cap < 1 ? 1 : cap;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, (unsigned)B);
}

// ---- DiffAugment ---------------------------------------------------------------------------------------------------
// [host-testable: diffaug]  (tests/test_host_logic.py runs these four kernels single-threaded under a g++ shim)
// params [7, B]: r_brightness, r_saturation, r_contrast (the raw U[0,1) draws), translation along H, along W (integers),
// cutout offset along H, along W (integers).  flags: 1 color, 2 translation, 4 cutout (applied in this order).
struct DiffAugParams {
    float bright, sat, con;     // additive brightness (r - 0.5), saturation factor 2 r, contrast factor r + 0.5
    int th, tw;                 // out[i][j] = in[i + th][j + tw] (zero outside)
    int ch_lo, ch_hi, cw_lo, cw_hi;   // cutout rows / columns (inclusive); empty when lo > hi
};

__device__ __forceinline__ DiffAugParams load_diffaug(const float* __restrict__ p, int B, int b, int H, int W, int flags) {
    DiffAugParams d;
    const bool color = flags & 1;
    d.bright = color ? __ldg(p + b) - 0.5f : 0.f;
    d.sat = color ? __ldg(p + B + b) * 2.f : 1.f;
    d.con = color ? __ldg(p + 2 * B + b) + 0.5f : 1.f;
    d.th = (flags & 2) ? (int)__ldg(p + 3 * B + b) : 0;
    d.tw = (flags & 2) ? (int)__ldg(p + 4 * B + b) : 0;
    d.ch_lo = 0; d.ch_hi = -1; d.cw_lo = 0; d.cw_hi = -1;
    if (flags & 4) {            // rand_cutout, third_party/diffaug.py:61-76: size = int(dim * 0.5 + 0.5), clamped index range
        const int sh = (int)((float)H * 0.5f + 0.5f), sw = (int)((float)W * 0.5f + 0.5f);
        const int lo_h = (int)__ldg(p + 5 * B + b) - sh / 2, lo_w = (int)__ldg(p + 6 * B + b) - sw / 2;
        d.ch_lo = max(lo_h, 0); d.ch_hi = min(lo_h + sh - 1, H - 1);
        d.cw_lo = max(lo_w, 0); d.cw_hi = min(lo_w + sw - 1, W - 1);
    }
    return d;
}

// brightness + saturation of one pixel in [-1, 1] space (rand_brightness / rand_saturation, diffaug.py:24-33)
__device__ __forceinline__ void diffaug_color_pre(float (&v)[3], const DiffAugParams& d) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (2.f * v[c] - 1.f) + d.bright;
    const float m = (v[0] + v[1] + v[2]) / 3.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (v[c] - m) * d.sat + m;
}

// sums[b] += sum over the image of the post-saturation values (the contrast mean is sums / (3 H W))
__global__ void __launch_bounds__(kT) diffaug_mean_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                                          float* __restrict__ sums, int B, int H, int W, int flags) {
    __shared__ float red[32];
    const int b = blockIdx.y, HW = H * W;
    const DiffAugParams d = load_diffaug(params, B, b, H, W, flags);
    const float* xb = x + (size_t)b * 3 * HW;
    float acc[1] = {0.f};
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float v[3] = {__ldg(xb + pix), __ldg(xb + HW + pix), __ldg(xb + 2 * HW + pix)};
        diffaug_color_pre(v, d);
        acc[0] += (v[0] + v[1]) + v[2];
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(sums + b, acc[0]);
}

__global__ void __launch_bounds__(kT) diffaug_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ params, const float* __restrict__ sums,
                                                           int B, int H, int W, int flags) {
    const int b = blockIdx.y, HW = H * W;
    const DiffAugParams d = load_diffaug(params, B, b, H, W, flags);
    const float mean = (flags & 1) ? __ldg(sums + b) / (float)(3 * HW) : 0.f;
    const float* xb = x + (size_t)b * 3 * HW;
    float* yb = y + (size_t)b * 3 * HW;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        const int i = pix / W, j = pix % W;
        const int si = i + d.th, sj = j + d.tw;
        float v[3] = {0.f, 0.f, 0.f};                             // zero fill in [-1, 1] space (diffaug.py:56)
        const bool cut = i >= d.ch_lo && i <= d.ch_hi && j >= d.cw_lo && j <= d.cw_hi;
        if (!cut && si >= 0 && si < H && sj >= 0 && sj < W) {
            const int sp = si * W + sj;
            v[0] = __ldg(xb + sp); v[1] = __ldg(xb + HW + sp); v[2] = __ldg(xb + 2 * HW + sp);
            if (flags & 1) {
                diffaug_color_pre(v, d);
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = (v[c] - mean) * d.con + mean;
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = 2.f * v[c] - 1.f;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) __stcs(yb + c * HW + pix, 0.5f * v[c] + 0.5f);
    }
}

// Gradient of the chain.  At source pixel (s, t): g = 0.5 * dy[s - th][t - tw] when that output exists and is not cut.
__device__ __forceinline__ void diffaug_grad_in(const float* __restrict__ dyb, int HW, int H, int W, int s, int t,
                                                const DiffAugParams& d, float (&g)[3]) {
    const int i = s - d.th, j = t - d.tw;
    g[0] = g[1] = g[2] = 0.f;
    if (i < 0 || i >= H || j < 0 || j >= W) return;
    if (i >= d.ch_lo && i <= d.ch_hi && j >= d.cw_lo && j <= d.cw_hi) return;
    const int op = i * W + j;
    g[0] = 0.5f * __ldg(dyb + op); g[1] = 0.5f * __ldg(dyb + HW + op); g[2] = 0.5f * __ldg(dyb + 2 * HW + op);
}

__global__ void __launch_bounds__(kT) diffaug_bwd_sum_kernel(const float* __restrict__ dy, const float* __restrict__ params,
                                                             float* __restrict__ gsums, int B, int H, int W, int flags) {
    __shared__ float red[32];
    const int b = blockIdx.y, HW = H * W;
    const DiffAugParams d = load_diffaug(params, B, b, H, W, flags);
    const float* dyb = dy + (size_t)b * 3 * HW;
    float acc[1] = {0.f};
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float g[3];
        diffaug_grad_in(dyb, HW, H, W, pix / W, pix % W, d, g);
        acc[0] += (g[0] + g[1]) + g[2];
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(gsums + b, acc[0]);
}

__global__ void __launch_bounds__(kT) diffaug_bwd_apply_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                               const float* __restrict__ params, const float* __restrict__ gsums,
                                                               int B, int H, int W, int flags) {
    const int b = blockIdx.y, HW = H * W;
    const DiffAugParams d = load_diffaug(params, B, b, H, W, flags);
    const float gmean = (flags & 1) ? __ldg(gsums + b) / (float)(3 * HW) : 0.f;
    const float* dyb = dy + (size_t)b * 3 * HW;
    float* dxb = dx + (size_t)b * 3 * HW;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float g[3];
        diffaug_grad_in(dyb, HW, H, W, pix / W, pix % W, d, g);
        if (flags & 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = d.con * g[c] + (1.f - d.con) * gmean;         // contrast adjoint
            const float m = (g[0] + g[1] + g[2]) / 3.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = d.sat * g[c] + (1.f - d.sat) * m;             // saturation adjoint
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) __stcs(dxb + c * HW + pix, 2.f * g[c]);                 // x -> 2 x - 1
    }
}

// [host-testable: end diffaug]

__global__ void __launch_bounds__(kT) noise_clamp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                             float* __restrict__ y, float sigma, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
        const float4 a = ldg_stream4(x + 4 * i), z = ldg_stream4(noise + 4 * i);
        float4 r;
        // x + noise * sigma evaluated as the reference does (a rounded product, then a rounded sum - no fma), then clamp
        r.x = fminf(fmaxf(__fadd_rn(a.x, __fmul_rn(z.x, sigma)), 0.f), 1.f);
        r.y = fminf(fmaxf(__fadd_rn(a.y, __fmul_rn(z.y, sigma)), 0.f), 1.f);
        r.z = fminf(fmaxf(__fadd_rn(a.z, __fmul_rn(z.z, sigma)), 0.f), 1.f);
        r.w = fminf(fmaxf(__fadd_rn(a.w, __fmul_rn(z.w, sigma)), 0.f), 1.f);
        stg_stream4(y + 4 * i, r);
    }
    for (long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
        y[t] = fminf(fmaxf(__fadd_rn(x[t], __fmul_rn(noise[t], sigma)), 0.f), 1.f);
}

// clamp backward: the gradient passes where 0 <= x + noise * sigma <= 1 (torch.clamp: inclusive at both ends)
__global__ void __launch_bounds__(kT) noise_clamp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                             const float* __restrict__ dy, float* __restrict__ dx,
                                                             float sigma, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float u = __fadd_rn(x[i], __fmul_rn(noise[i], sigma));
        dx[i] = (u >= 0.f && u <= 1.f) ? dy[i] : 0.f;
    }
}

inline int flat_grid(long long n) {
    long long blocks = (n + kT - 1) / kT;
    const long long cap = 148LL * 16;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}


// ------------------------------------------------------------------------------------------------
// Parameter block of the fused SimCLR chain from raw draws (one launch instead of ~20 ATen ops per call):
//   boxes [4,B]  sx, sy, bx, by  - numpy draws of RandomResizeCropLayer.sample (augment/spatial.py:119-143), host-staged
//   u     [7,B]  i.i.d. U[0,1)   - device draws: flip, apply-jitter, contrast, hue, saturation, value, apply-gray
//   order_src    device scalar (0 / 1) or NULL -> params row 11 is written only when rows == 12
// Mapping = the reference's own: bernoulli(p) = (u < p) (HorizontalFlipLayer: sign = 2*b - 1, spatial.py:86-88;
// RandomApply mask, augment/__init__.py:100-103), uniform_(lo, hi) = lo + (hi - lo) * u (color_jitter.py:44-63).
// ------------------------------------------------------------------------------------------------
struct SimclrDrawCfg {
    float p_flip, p_jitter, p_gray;
    float c_lo, c_hi, h_lo, h_hi, s_lo, s_hi, v_lo, v_hi;
};

__global__ void __launch_bounds__(256)
simclr_params_kernel(const float* __restrict__ boxes, const float* __restrict__ u, const float* __restrict__ order_src,
                     float* __restrict__ params, int B, int rows, SimclrDrawCfg c) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B) return;
#pragma unroll
    for (int r = 0; r < 4; ++r) params[(size_t)r * B + i] = boxes[(size_t)r * B + i];
    params[(size_t)4 * B + i] = (u[i] < c.p_flip) ? 1.f : -1.f;
    params[(size_t)5 * B + i] = (u[(size_t)B + i] < c.p_jitter) ? 1.f : 0.f;
    params[(size_t)6 * B + i] = c.c_lo + (c.c_hi - c.c_lo) * u[(size_t)2 * B + i];
    params[(size_t)7 * B + i] = c.h_lo + (c.h_hi - c.h_lo) * u[(size_t)3 * B + i];
    params[(size_t)8 * B + i] = c.s_lo + (c.s_hi - c.s_lo) * u[(size_t)4 * B + i];
    params[(size_t)9 * B + i] = c.v_lo + (c.v_hi - c.v_lo) * u[(size_t)5 * B + i];
    params[(size_t)10 * B + i] = (u[(size_t)6 * B + i] < c.p_gray) ? 1.f : 0.f;
    if (rows > 11) params[(size_t)11 * B + i] = order_src ? order_src[0] : 0.f;
}

}  // namespace

extern "C" int cb200_shift_flip_fwd(const float* x, float* y, const float* params, int B, int P, int H, int W,
                                    int padding_mode, void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && P > 0 && H > 0 && W > 0, "shift_flip_fwd: bad shape");
    CB200_CHECK_ARG(padding_mode >= kZeros && padding_mode <= kReflection,
                    "shift_flip_fwd: padding_mode %d (0 zeros, 1 border, 2 reflection)", padding_mode);
    CB200_CHECK_ARG((size_t)(H + W) * sizeof(int) <= 48 * 1024, "shift_flip_fwd: H + W = %d too large", H + W);
    if (B == 0) return CB200_OK;
    // the kernel picks the float4-store path from W alone, so an unaligned y is rejected rather than mis-served
    CB200_CHECK_ARG(W % 4 != 0 || (reinterpret_cast<uintptr_t>(y) & 15) == 0, "shift_flip_fwd: y must be 16-byte aligned");
    const long long items = (long long)P * H * W / (W % 4 == 0 ? 4 : 1);
    shift_flip_fwd_kernel<<<image_grid(B, items), kT, (size_t)(H + W) * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
        x, y, params, B, P, H, W, padding_mode);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("shift_flip_fwd");
    return CB200_OK;
}

extern "C" int cb200_shift_flip_bwd(const float* dy, float* dx, const float* params, int B, int P, int H, int W,
                                    int padding_mode, void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && P > 0 && H > 0 && W > 0, "shift_flip_bwd: bad shape");
    CB200_CHECK_ARG(padding_mode >= kZeros && padding_mode <= kReflection,
                    "shift_flip_bwd: padding_mode %d (0 zeros, 1 border, 2 reflection)", padding_mode);
    CB200_CHECK_ARG((size_t)(H + W) * sizeof(int) <= 48 * 1024, "shift_flip_bwd: H + W = %d too large", H + W);
    if (B == 0) return CB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * P * H * W, st);
    if (e != cudaSuccess) { cb200_set_error("shift_flip_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    shift_flip_bwd_kernel<<<image_grid(B, (long long)P * H * W), kT, (size_t)(H + W) * sizeof(int), st>>>(
        dy, dx, params, B, P, H, W, padding_mode);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("shift_flip_bwd");
    return CB200_OK;
}

extern "C" int cb200_noise_clamp_fwd(const float* x, const float* noise, float* y, float sigma, long long n, void* stream) {
    CB200_CHECK_ARG(n >= 0, "noise_clamp_fwd: negative size");
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
                    "noise_clamp_fwd: pointers must be 16-byte aligned");
    if (n == 0) return CB200_OK;
    noise_clamp_fwd_kernel<<<flat_grid((n + 3) / 4), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, noise, y, sigma, n);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("noise_clamp_fwd");
    return CB200_OK;
}

extern "C" int cb200_noise_clamp_bwd(const float* x, const float* noise, const float* dy, float* dx, float sigma,
                                     long long n, void* stream) {
    CB200_CHECK_ARG(n >= 0, "noise_clamp_bwd: negative size");
    if (n == 0) return CB200_OK;
    noise_clamp_bwd_kernel<<<flat_grid(n), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, noise, dy, dx, sigma, n);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("noise_clamp_bwd");
    return CB200_OK;
}

// DiffAugment forward / backward (third_party/diffaug.py:8-21 with AUGMENT_FNS 'color', 'translation', 'cutout' in that
// order; flags = 1 | 2 | 4).  x, y, dy, dx [B,3,H,W]; params [7,B] (see load_diffaug); sums / gsums [B] scratch.
extern "C" int cb200_diffaug_fwd(const float* x, float* y, const float* params, float* sums, int B, int H, int W, int flags,
                                 void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && H > 0 && W > 0 && flags >= 0 && flags <= 7, "diffaug_fwd: bad shape / flags");
    if (B == 0) return CB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid = image_grid(B, (long long)H * W);
    if (flags & 1) {
        cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)B, st);
        if (e != cudaSuccess) { cb200_set_error("diffaug_fwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
        diffaug_mean_kernel<<<grid, kT, 0, st>>>(x, params, sums, B, H, W, flags);
        CB200_COUNT_LAUNCH();
    }
    diffaug_apply_kernel<<<grid, kT, 0, st>>>(x, y, params, sums, B, H, W, flags);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("diffaug_fwd");
    return CB200_OK;
}

extern "C" int cb200_diffaug_bwd(const float* dy, float* dx, const float* params, float* gsums, int B, int H, int W,
                                 int flags, void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && H > 0 && W > 0 && flags >= 0 && flags <= 7, "diffaug_bwd: bad shape / flags");
    if (B == 0) return CB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid = image_grid(B, (long long)H * W);
    if (flags & 1) {
        cudaError_t e = cudaMemsetAsync(gsums, 0, sizeof(float) * (size_t)B, st);
        if (e != cudaSuccess) { cb200_set_error("diffaug_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
        diffaug_bwd_sum_kernel<<<grid, kT, 0, st>>>(dy, params, gsums, B, H, W, flags);
        CB200_COUNT_LAUNCH();
    }
    diffaug_bwd_apply_kernel<<<grid, kT, 0, st>>>(dy, dx, params, gsums, B, H, W, flags);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("diffaug_bwd");
    return CB200_OK;
}

// params[rows,B] (rows = 11, or 12 with the per-image jitter order in row 11) from host-staged crop boxes [4,B] and
// device uniforms u[7,B]; cfg = {p_flip, p_jitter, p_gray, contrast lo/hi, hue lo/hi, saturation lo/hi, value lo/hi}.
extern "C" int cb200_augment_simclr_params(const float* boxes, const float* u, const float* order_src, float* params,
                                           int B, int rows, const float* cfg11, void* stream) {
    CB200_CHECK_ARG(B > 0 && (rows == 11 || rows == 12), "augment_simclr_params: B=%d rows=%d", B, rows);
    SimclrDrawCfg c;
    c.p_flip = cfg11[0]; c.p_jitter = cfg11[1]; c.p_gray = cfg11[2];
    c.c_lo = cfg11[3]; c.c_hi = cfg11[4]; c.h_lo = cfg11[5]; c.h_hi = cfg11[6];
    c.s_lo = cfg11[7]; c.s_hi = cfg11[8]; c.v_lo = cfg11[9]; c.v_hi = cfg11[10];
    simclr_params_kernel<<<(B + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, u, order_src, params, B,
                                                                                          rows, c);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("augment_simclr_params");
    return CB200_OK;
}
