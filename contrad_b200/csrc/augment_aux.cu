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
// Both are HBM-bound: 8 B/element (shift_flip fwd), 12 B/element (noise fwd: x, noise in, y out).
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
