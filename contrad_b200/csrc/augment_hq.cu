// High-resolution tail of the SimCLR chain for sm_100a: `simclr_hq` and
// `simclr_hq_cutout` append RandomApply(GaussianBlur, 0.5) and RandomApply(CutOut, 0.5) to the fused chain
// (reference: augment/__init__.py:52-78,115-133; augment/spatial.py:151-181).
//
//   GaussianBlur   the reference builds the DENSE k x k outer-product Gaussian (k = 2*floor((H/10)/2)+1: 3 at 32x32,
//                  51 at 512x512; one sigma per batch, drawn on the host) and runs kornia.filter2D with 'reflect' padding:
//                  2*k^2 FLOPs per element (4.1 GFLOP per 512x512 image).  The kernel is an outer product of two
//                  normalised 1-D Gaussians, so two 1-D passes give the same result with 2*2k FLOPs per element.
//                  Per-sample Bernoulli mask: unselected images are copied through (RandomApply blends with 0/1 masks).
//                  Backward = the adjoint of (reflect-pad -> correlate), per axis: the reflected halo folds back onto
//                  the border pixels.
//   CutOut         zeroes a (2p+1)^2 square around a random centre, clipped at the border (the reference's
//                  conv1d-of-one-hot / einsum mask); its backward is the same masking of the gradient.
// Both are HBM-bound elementwise passes (blur: 2 x 8 B/elem, cutout: 8 B/elem).
#include "common.cuh"
#include "contrad_b200.h"

namespace {

constexpr int kT = 256;
constexpr int kMaxTaps = 129;

inline int grid_for(long long work, int per_block = kT, int cap = 148 * 16) {
    long long g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

__device__ __forceinline__ int reflect_index(int p, int n) {       // F.pad(mode='reflect'): -1 -> 1, n -> n-2
    if (p < 0) p = -p;
    if (p > n - 1) p = 2 * (n - 1) - p;
    return p;
}

// One 1-D pass along `axis` (0: rows / H, 1: columns / W) over x[B, P, H, W] (P planes per image).
//   adjoint == 0: y[i] = sum_t w[t] * x[reflect(i + t - r)]
//   adjoint == 1: y[i] = sum_t w[t] * (x[i-d] + [i>0] x[-i-d] + [i<n-1] x[2(n-1)-i-d]),  d = t - r, terms inside [0, n)
// Images with on[b] == 0 are skipped (pass_through == 0: left untouched, the next pass does not read them) or copied
// from `orig` (pass_through == 1: the final pass of the pair).
__global__ void __launch_bounds__(kT) blur_axis_kernel(const float* __restrict__ x, const float* __restrict__ orig,
                                                       float* __restrict__ y, const float* __restrict__ taps,
                                                       const float* __restrict__ on, int B, int P, int H, int W, int k,
                                                       int axis, int adjoint, int pass_through) {
    __shared__ float w[kMaxTaps];
    for (int i = threadIdx.x; i < k; i += blockDim.x) w[i] = __ldg(taps + i);
    __syncthreads();
    const int r = k / 2;
    const long long per_img = (long long)P * H * W, total = per_img * B;
    const int n = axis == 0 ? H : W;
    const int stride = axis == 0 ? W : 1;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / per_img);
        if (on != nullptr && __ldg(on + b) == 0.f) {
            if (pass_through) y[idx] = orig[idx];
            continue;
        }
        const int col = (int)(idx % W), row = (int)((idx / W) % H);
        const int i = axis == 0 ? row : col;
        const float* line = x + (idx - (long long)i * stride);        // element 0 of this row / column
        float acc = 0.f;
        if (adjoint ? (i - r > 0 && i + r < n - 1) : (i - r >= 0 && i + r < n)) {
            // interior: no reflection involved (for the adjoint the reflected halo folds back onto 1..r and
            // n-2-r..n-2, hence the stricter test); the adjoint reads the taps mirrored: x[i - d] = x[i + r - t]
            const float* pt = line + (long long)(adjoint ? i + r : i - r) * stride;
            const long long step = adjoint ? -(long long)stride : (long long)stride;
#pragma unroll 4
            for (int t = 0; t < k; ++t) acc = fmaf(w[t], pt[t * step], acc);
        } else if (!adjoint) {
            for (int t = 0; t < k; ++t) acc = fmaf(w[t], line[(long long)reflect_index(i + t - r, n) * stride], acc);
        } else {
            for (int t = 0; t < k; ++t) {
                const int d = t - r;
                float s = 0.f;
                int o = i - d;
                if (o >= 0 && o < n) s += line[(long long)o * stride];
                o = -i - d;
                if (i > 0 && o >= 0 && o < n) s += line[(long long)o * stride];
                o = 2 * (n - 1) - i - d;
                if (i < n - 1 && o >= 0 && o < n) s += line[(long long)o * stride];
                acc = fmaf(w[t], s, acc);
            }
        }
        y[idx] = acc;
    }
}

// float4 along W (W % 4 == 0): params as below
__global__ void __launch_bounds__(kT) cutout_vec_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                        const float* __restrict__ params, int B, int P, int H, int W4, int half) {
    const int b = blockIdx.y;
    const long long per4 = (long long)P * H * W4;
    const float4* xb = x + (long long)b * per4;
    float4* yb = y + (long long)b * per4;
    const bool on = __ldg(params + b) != 0.f;
    const int hc = (int)__ldg(params + B + b), wc = (int)__ldg(params + 2 * B + b);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = xb[i];
        if (on) {
            const int col = (int)(i % W4) * 4, row = (int)((i / W4) % H);
            if (abs(row - hc) <= half) {
                if (abs(col - wc) <= half) v.x = 0.f;
                if (abs(col + 1 - wc) <= half) v.y = 0.f;
                if (abs(col + 2 - wc) <= half) v.z = 0.f;
                if (abs(col + 3 - wc) <= half) v.w = 0.f;
            }
        }
        yb[i] = v;
    }
}

// params [3, B]: on (0/1), h centre, w centre.  y = x outside the square, 0 inside (only where on[b] != 0).
__global__ void __launch_bounds__(kT) cutout_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                    const float* __restrict__ params, int B, int P, int H, int W, int half) {
    const long long per_img = (long long)P * H * W, total = per_img * B;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / per_img);
        float v = x[idx];
        if (__ldg(params + b) != 0.f) {
            const int col = (int)(idx % W), row = (int)((idx / W) % H);
            const int hc = (int)__ldg(params + B + b), wc = (int)__ldg(params + 2 * B + b);
            if (abs(row - hc) <= half && abs(col - wc) <= half) v = 0.f;
        }
        y[idx] = v;
    }
}

}  // namespace

// y = on[b] ? gaussian_blur(x[b]) : x[b] with the separable 1-D kernel `taps` (k odd, k/2 < min(H, W)); `tmp` is scratch
// of x's size.  adjoint = 1 applies the transpose (backward pass: x = dy, y = dx).
extern "C" int cb200_gaussian_blur(const float* x, float* tmp, float* y, const float* taps, const float* on, int B, int P,
                                   int H, int W, int k, int adjoint, void* stream) {
    CB200_CHECK_ARG(B > 0 && P > 0 && H > 0 && W > 0, "gaussian_blur: empty tensor");
    CB200_CHECK_ARG(k >= 1 && (k & 1) && k <= kMaxTaps && k / 2 < H && k / 2 < W,
                    "gaussian_blur: kernel size %d must be odd, <= %d and its radius smaller than the image", k, kMaxTaps);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for((long long)B * P * H * W);
    // forward: rows then columns; adjoint: the same two (commuting, each self-contained) passes transposed
    blur_axis_kernel<<<grid, kT, 0, st>>>(x, nullptr, tmp, taps, on, B, P, H, W, k, 1, adjoint, 0);
    CB200_COUNT_LAUNCH();
    blur_axis_kernel<<<grid, kT, 0, st>>>(tmp, x, y, taps, on, B, P, H, W, k, 0, adjoint, 1);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("gaussian_blur");
    return CB200_OK;
}

extern "C" int cb200_cutout(const float* x, float* y, const float* params, int B, int P, int H, int W, int length,
                            void* stream) {
    CB200_CHECK_ARG(B > 0 && P > 0 && H > 0 && W > 0, "cutout: empty tensor");
    CB200_CHECK_ARG(length >= 1 && (length & 1), "cutout: length %d must be odd (augment/spatial.py:155-156)", length);
    if (W % 4 == 0 && B <= 65535 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        dim3 grid(grid_for((long long)P * H * (W / 4), kT, (148 * 16 + B - 1) / B), B);
        cutout_vec_kernel<<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), params, B, P, H, W / 4, (length - 1) / 2);
    } else {
        cutout_kernel<<<grid_for((long long)B * P * H * W), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, params, B, P, H, W,
                                                                                                        (length - 1) / 2);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("cutout");
    return CB200_OK;
}
