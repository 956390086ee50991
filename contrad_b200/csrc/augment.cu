// Fused SimCLR augmentation chain (random-resized-crop -> hflip -> color-jitter -> grayscale)
// for sm_100a: ONE forward kernel and ONE backward kernel replace the ~120 ATen ops / ~25
// full-tensor round trips of the reference chain (reference: augment/__init__.py:106-112,
// augment/spatial.py:84-148, augment/color_jitter.py:44-104, augment/utils.py:27-63; exact
// arithmetic restated in SURVEY.md A.1 and oracle/contrad_oracle.py).
//
// Roofline: HBM-bound, 8 algorithmic bytes per tensor element (4 read + 4 written, fp32).
//
// Small-image path (H*W <= 4096, i.e. every 32x32 / 64x64 config): one CTA owns one image.
//   * the [3,H,W] image is staged into shared memory with coalesced, L1-bypassing float4 loads;
//   * every thread keeps its output pixels (quads of 4 along W, 3 channels) in registers across
//     the per-channel mean that `adjust_contrast` needs (a block reduction - the image never
//     leaves the SM between the stages of the chain);
//   * bilinear taps (reflection padding, align_corners=False) are gathered from shared memory;
//     the horizontal flip is an index mirror folded into the gather (exact at power-of-two
//     sizes);
//   * output is written once with float4 streaming stores.
// Per-sample parameters arrive as an SoA block [11, B] (see PARAM_FIELDS in the oracle /
// include/contrad_b200.h); sampling them stays on the host so that the reference's numpy +
// torch RNG interleaving is reproduced exactly.
//
// Backward (needed in every G step): HSV is straight-through (color_jitter.py:97-104); contrast is
// dx = f*dy' + (1-f)*mean(dy') with dy' = dy*1[0<=u<=1]; gray/blend are linear; the crop is the
// transposed bilinear gather, accumulated with shared-memory atomics and written out once.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>

namespace {

constexpr int kMaxThreads = 256;

struct SampleParams {
    float sx, sy, bx, by, flip, cj_on, fc, fh, fs, fv, gray_on;
};

__device__ __forceinline__ SampleParams load_params(const float* __restrict__ p, int B, int b) {
    SampleParams s;
    s.sx = __ldg(p + 0 * B + b);
    s.sy = __ldg(p + 1 * B + b);
    s.bx = __ldg(p + 2 * B + b);
    s.by = __ldg(p + 3 * B + b);
    s.flip = __ldg(p + 4 * B + b);
    s.cj_on = __ldg(p + 5 * B + b);
    s.fc = __ldg(p + 6 * B + b);
    s.fh = __ldg(p + 7 * B + b);
    s.fs = __ldg(p + 8 * B + b);
    s.fv = __ldg(p + 9 * B + b);
    s.gray_on = __ldg(p + 10 * B + b);
    return s;
}

// Colour-jitter order (0: contrast then hsv, 1: hsv then contrast).  `order` >= 0 is the launch-wide value;
// order < 0 reads it per image from row 11 of the parameter block (device-resident, so that a CUDA graph of the
// train step can be replayed with a freshly drawn order).
__device__ __forceinline__ int resolve_order(const float* __restrict__ p, int B, int b, int order) {
    return order >= 0 ? order : (__ldg(p + 11 * B + b) != 0.f ? 1 : 0);
}

// grid_sample(padding_mode='reflection', align_corners=False): reflect about -0.5 and size-0.5,
// clip to [0, size-1].
// (fmod via floor: exact for power-of-two sizes; elsewhere the two branches agree at the reflection points.)
__device__ __forceinline__ float fold_coord(float coord, float size, float inv_size) {
    float t = fabsf(coord + 0.5f);
    float fl = floorf(t * inv_size);
    float extra = fmaf(-fl, size, t);
    int flips = (int)fl;
    float r = (flips & 1) ? (size - extra - 0.5f) : (extra - 0.5f);
    return fminf(fmaxf(r, 0.f), size - 1.f);
}

struct Tap {
    int i0, i1;
    float w0, w1;
};

// Source taps of output index `o` along an axis of length n for scale s and bias b.
__device__ __forceinline__ Tap axis_tap(int o, int n, float s, float b) {
    float fn = (float)n;
    float inv = 1.f / fn;
    float base = (2.f * (float)o + 1.f) * inv - 1.f;
    float g = s * base + b;
    float p = fold_coord(((g + 1.f) * fn - 1.f) * 0.5f, fn, inv);
    float p0 = floorf(p);
    Tap t;
    t.i0 = (int)p0;
    t.w1 = p - p0;
    t.w0 = 1.f - t.w1;
    t.i1 = t.i0 + 1;
    if (t.i1 > n - 1) {   // out-of-range tap contributes zero
        t.i1 = n - 1;
        t.w1 = 0.f;
    }
    return t;
}

// atan2(y, x) / (2 pi) folded into [0, 1): 8-term minimax odd polynomial on [0,1] (max abs error 1.2e-7 rad), its
// coefficients pre-multiplied by 1 / (2 pi) so that the result is in turns without a final multiply, + octant fix-ups;
// the division is the fast reciprocal (2 ulp) - the result feeds a piecewise-linear colour wheel with slope <= 6, so
// 1e-7 in the turn fraction is far below the fp32 noise of the chain.  mx == 0 (a gray pixel) gives t = 0 -> hue 0 like
// atan2(0, 0).
template <int HV>
__device__ __forceinline__ float hue_turns(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = HV >= 2 ? fast_div(mn, fmaxf(mx, 1e-30f)) : __fdividef(mn, fmaxf(mx, 1e-30f));
    const float u = t * t;
    float p = -0.00064527396948442972f;
    p = fmaf(p, u, 0.0034794815687994203f);
    p = fmaf(p, u, -0.0088985447628120544f);
    p = fmaf(p, u, 0.015345893529268476f);
    p = fmaf(p, u, -0.022136211470001277f);
    p = fmaf(p, u, 0.031745931893189944f);
    p = fmaf(p, u, -0.053046118722407817f);
    p = fmaf(p, u, 0.15915483874178302f);
    float a = p * t;                                   // atan(mn/mx) / (2 pi) in [0, 1/8]
    if (ay > ax) a = 0.25f - a;                        // first quadrant
    if (x < 0.f) a = 0.5f - a;                         // upper half plane
    if (y < 0.f) a = 1.f - a;                          // == (atan2 < 0 ? atan2 + 2 pi : atan2)
    return a;
}

// RandomHSVFunction.forward on one pixel (augment/color_jitter.py:83-95, augment/utils.py:27-38,55-63).
// `hshift` = (f_h * 255) / 360 is hoisted per image.
// hsv -> rgb (utils.py:55-63: k = (n + 6 h) mod 6, t = clamp(min(k, 4 - k), 0, 1), out = v - c t for n = 5, 3, 1) is
// evaluated as t = sat(2 - d) with d the CIRCULAR distance (period 6) of 6 h to the channel's centre 3, 5, 1:
// min(k, 4 - k) = 2 - |k - 2| and |k - 2| is that distance; for the centre 3 it never wraps.  Same function, 10 instead of
// 18 instructions for the three channels (the chain is issue-bound, profiles/prof_r1_augment.md).
// HV selects the build: 1 = first (all other kernels), 2 = guard-free reciprocals (bit-identical for inputs in range:
// cmax + 1e-8 >= 1e-30; 8 instructions per pixel fewer incl. the predicate parking the guards caused).
template <int HV = 1>
__device__ __forceinline__ void hsv_jitter(float& r, float& g, float& b, float hshift, float fs, float fv) {
    float cmax = fmaxf(r, fmaxf(g, b));
    float cmin = fminf(r, fminf(g, b));
    float hue = hue_turns<HV>(1.7320508075688772f * (g - b), 2.f * r - g - b);     // finite for finite inputs
    // the reference zeroes non-finite hsv entries (utils.py:37): a NaN saturation is mapped to 0 by the __saturatef below
    // (it cannot be +-inf: cmin / (cmax + 1e-8) is finite or NaN), the value needs the explicit test (+inf -> 0)
    const float sat = 1.f - (HV >= 2 ? fast_div(cmin, cmax + 1e-8f) : __fdividef(cmin, cmax + 1e-8f));
    const float val = isfinite(cmax) ? cmax : 0.f;
    float h = hue + hshift;
    h = h - floorf(h);
    const float s = __saturatef(sat * fs);
    const float v = __saturatef(val * fv);
    h = __saturatef(h);
    const float c = v * s;
    const float h6 = h * 6.f;
    const float dr = fabsf(h6 - 3.f);
    const float ag = fabsf(h6 - 5.f), ab = fabsf(h6 - 1.f);
    const float dg = fminf(ag, 6.f - ag), db = fminf(ab, 6.f - ab);
    r = fmaf(-c, __saturatef(2.f - dr), v);
    g = fmaf(-c, __saturatef(2.f - dg), v);
    b = fmaf(-c, __saturatef(2.f - db), v);
}

__device__ __forceinline__ float clamp01(float x) { return __saturatef(x); }   // one FADD.SAT; NaN -> 0 like fmin(fmax())

// Per-image tap tables in shared memory: column taps already include the horizontal flip.
struct TapTables {
    int* xi0; int* xi1; float* xw0; float* xw1;     // [W]
    int* yi0; int* yi1; float* yw0; float* yw1;     // [H]
};

__device__ __forceinline__ TapTables carve_taps(float* base, int H, int W) {
    TapTables t;
    t.xi0 = reinterpret_cast<int*>(base);          t.xi1 = t.xi0 + W;
    t.xw0 = base + 2 * W;                          t.xw1 = base + 3 * W;
    float* yb = base + 4 * W;
    t.yi0 = reinterpret_cast<int*>(yb);            t.yi1 = t.yi0 + H;
    t.yw0 = yb + 2 * H;                            t.yw1 = yb + 3 * H;
    return t;
}

__device__ __forceinline__ void fill_taps(const TapTables& t, int H, int W, const SampleParams& sp) {
    for (int e = threadIdx.x; e < W + H; e += blockDim.x) {
        if (e < W) {
            const int jj = (sp.flip < 0.f) ? (W - 1 - e) : e;
            const Tap a = axis_tap(jj, W, sp.sx, sp.bx);
            t.xi0[e] = a.i0; t.xi1[e] = a.i1; t.xw0[e] = a.w0; t.xw1[e] = a.w1;
        } else {
            const int i = e - W;
            const Tap a = axis_tap(i, H, sp.sy, sp.by);
            t.yi0[i] = a.i0; t.yi1[i] = a.i1; t.yw0[i] = a.w0; t.yw1[i] = a.w1;
        }
    }
}

// crop+flip for the quad (row i, cols j0..j0+3): out[c][k]
__device__ __forceinline__ void gather_quad(const float* xs, int H, int W, int i, int j0, const TapTables& t,
                                            float (&out)[3][4]) {
    const int HW = H * W;
    const float wy0 = t.yw0[i], wy1 = t.yw1[i];
    const float* r0 = xs + t.yi0[i] * W;
    const float* r1 = xs + t.yi1[i] * W;
    const int4 i0 = *reinterpret_cast<const int4*>(t.xi0 + j0);
    const int4 i1 = *reinterpret_cast<const int4*>(t.xi1 + j0);
    const float4 w0 = *reinterpret_cast<const float4*>(t.xw0 + j0);
    const float4 w1 = *reinterpret_cast<const float4*>(t.xw1 + j0);
    const int a0[4] = {i0.x, i0.y, i0.z, i0.w}, a1[4] = {i1.x, i1.y, i1.z, i1.w};
    const float b0[4] = {w0.x, w0.y, w0.z, w0.w}, b1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float w00 = b0[k] * wy0, w01 = b1[k] * wy0, w10 = b0[k] * wy1, w11 = b1[k] * wy1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            out[c][k] = r0[c * HW + a0[k]] * w00 + r0[c * HW + a1[k]] * w01 + r1[c * HW + a0[k]] * w10 +
                        r1[c * HW + a1[k]] * w11;
        }
    }
}

// (async staging helpers - bar_init / bar_expect_tx / bulk_load / bar_wait - live in common.cuh)

// Forward: PERSISTENT CTAs (grid = resident CTAs) walk the batch with a two-deep prefetch ring: while image i is
// being processed, image i + gridDim.x is already in flight into the other shared-memory buffer (cp.async.bulk +
// mbarrier), so global-load latency never sits on the critical path (it was 25 % of the stall samples of the
// non-pipelined version, profiles/prof_r1_augment.md).  S != 0 bakes the image size into the code.
template <int QPT, int S>
__global__ void __launch_bounds__(kMaxThreads)
augment_simclr_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params,
                          int B, int H_, int W_, int order) {
    extern __shared__ __align__(128) float smem[];
    const int H = S ? S : H_, W = S ? S : W_;
    const int HW = H * W, Wq = W >> 2, nquads = HW >> 2;
    float* red = smem + 6 * HW;            // [3*32]
    const TapTables taps = carve_taps(red + 96, H, W);
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 96 + 4 * W + 4 * H);
    const uint32_t img_bytes = (uint32_t)(3 * HW * sizeof(float));

    if (threadIdx.x == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    int b = blockIdx.x;
    if (threadIdx.x == 0 && b < B) {
        bar_expect_tx(&bars[0], img_bytes);
        bulk_load(smem, x + (size_t)b * 3 * HW, img_bytes, &bars[0]);
    }
    for (int it = 0; b < B; b += gridDim.x, ++it) {
        const int buf = it & 1;
        const int nb = b + gridDim.x;
        if (threadIdx.x == 0 && nb < B) {          // prefetch the next image of this CTA into the other buffer
            bar_expect_tx(&bars[buf ^ 1], img_bytes);
            bulk_load(smem + (buf ^ 1) * 3 * HW, x + (size_t)nb * 3 * HW, img_bytes, &bars[buf ^ 1]);
        }
        const SampleParams sp = load_params(params, B, b);
        const int ord = resolve_order(params, B, b, order);
        const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
        fill_taps(taps, H, W, sp);
        bar_wait(&bars[buf], (uint32_t)(it >> 1) & 1u);
        __syncthreads();
        const float* xs = smem + buf * 3 * HW;

        float v[QPT][3][4];
        bool live[QPT];
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            int quad = threadIdx.x + q * blockDim.x;
            live[q] = quad < nquads;
            if (live[q]) {
                gather_quad(xs, H, W, quad / Wq, (quad % Wq) * 4, taps, v[q]);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[q][c][k] = 0.f;
            }
        }

        if (sp.cj_on != 0.f) {          // uniform across the CTA (one image per CTA iteration)
            if (ord == 1) {
#pragma unroll
                for (int q = 0; q < QPT; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (live[q]) hsv_jitter(v[q][0][k], v[q][1][k], v[q][2][k], hshift, sp.fs, sp.fv);
            }
            float sums[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < QPT; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) sums[c] += (v[q][c][0] + v[q][c][1]) + (v[q][c][2] + v[q][c][3]);
            block_sum<3>(sums, red);
            const float inv = 1.f / (float)HW;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float m = sums[c] * inv;
#pragma unroll
                for (int q = 0; q < QPT; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[q][c][k] = clamp01((v[q][c][k] - m) * sp.fc + m);
            }
            if (ord == 0) {
#pragma unroll
                for (int q = 0; q < QPT; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (live[q]) hsv_jitter(v[q][0][k], v[q][1][k], v[q][2][k], hshift, sp.fs, sp.fv);
            }
        }
        float* yb = y + (size_t)b * 3 * HW;
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            if (!live[q]) continue;
            int quad = threadIdx.x + q * blockDim.x;
            if (sp.gray_on != 0.f) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float l = 0.299f * v[q][0][k] + 0.587f * v[q][1][k] + 0.114f * v[q][2][k];
                    v[q][0][k] = l; v[q][1][k] = l; v[q][2][k] = l;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                stg_stream4(yb + c * HW + quad * 4, make_float4(v[q][c][0], v[q][c][1], v[q][c][2], v[q][c][3]));
        }
        __syncthreads();      // tap tables / reduction scratch / this image buffer are reused by the next iteration
    }
}

// Sum of three per-thread values over a 256-thread CTA (8 warps), result in every thread.  The generic block_sum<3>
// spends 30 shuffles + 30 adds per thread (three independent 5-step butterflies, twice); here the three values share ONE
// butterfly: after the xor-16 and xor-8 exchanges every lane carries a single value (lanes 0-7: v0, 8-15: v1, 16-23: v2,
// 24-31: zero), three more steps finish the warp sums, and the 3 x 8 warp partials are folded by one 3-step butterfly and
// three broadcasts: 12 shuffles.  `scratch` holds 24 floats; the CALLER must synchronise the CTA before reusing it (the
// image loop does, at the end of every iteration).
__device__ __forceinline__ void block_sum3_256(float (&v)[3], float* scratch) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hi = (lane & 16) != 0;
    // lower half-warp keeps (v0, v1) and hands v2 to its partner; upper half keeps v2 and hands (v0, v1) over
    const float r0 = __shfl_xor_sync(full, hi ? v[0] : v[2], 16);
    const float r1 = __shfl_xor_sync(full, hi ? v[1] : 0.f, 16);
    float a = (hi ? v[2] : v[0]) + r0;
    const float b = (hi ? 0.f : v[1]) + r1;
    const bool q = (lane & 8) != 0;
    a = (q ? b : a) + __shfl_xor_sync(full, q ? a : b, 8);
    a += __shfl_xor_sync(full, a, 4);
    a += __shfl_xor_sync(full, a, 2);
    a += __shfl_xor_sync(full, a, 1);
    if ((lane & 7) == 0 && lane < 24) scratch[(lane >> 3) * 8 + warp] = a;      // [value][warp]
    __syncthreads();
    float t = (lane < 24) ? scratch[lane] : 0.f;
    t += __shfl_xor_sync(full, t, 4);
    t += __shfl_xor_sync(full, t, 2);
    t += __shfl_xor_sync(full, t, 1);
    v[0] = __shfl_sync(full, t, 0);
    v[1] = __shfl_sync(full, t, 8);
    v[2] = __shfl_sync(full, t, 16);
}

// Per-image arithmetic of the column-mapped forward kernels: thread = output column j, rows i_first .. i_first+NPX-1.
// `xs` is the fp32 [3,S,S] source image in shared memory, `yb` points at this thread's first output element.
template <int S>
__device__ __forceinline__ void cols_process(const float* xs, const float4* xtap, const float4* ytap, float* red,
                                             const SampleParams& sp, int ord, float hshift, float* yb, int j,
                                             int i_first) {
    constexpr int HW = S * S, NPX = HW / kMaxThreads;
    float v[NPX][3];
    {
        const float4 tx = xtap[j];
        const int x0 = __float_as_int(tx.x), x1 = __float_as_int(tx.y);
#pragma unroll
        for (int m = 0; m < NPX; ++m) {
            const float4 ty = ytap[i_first + m];
            const float* r0 = xs + __float_as_int(ty.x);
            const float* r1 = xs + __float_as_int(ty.y);
            const float w00 = tx.z * ty.z, w01 = tx.w * ty.z, w10 = tx.z * ty.w, w11 = tx.w * ty.w;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[m][c] = r0[c * HW + x0] * w00 + r0[c * HW + x1] * w01 + r1[c * HW + x0] * w10 +
                          r1[c * HW + x1] * w11;
        }
    }
    if (sp.cj_on != 0.f) {          // uniform across the CTA
        if (ord == 1) {
#pragma unroll
            for (int m = 0; m < NPX; ++m) hsv_jitter(v[m][0], v[m][1], v[m][2], hshift, sp.fs, sp.fv);
        }
        float sums[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < NPX; ++m)
#pragma unroll
            for (int c = 0; c < 3; ++c) sums[c] += v[m][c];
        block_sum3_256(sums, red);          // blockDim.x == 256; `red` is not touched again before the CTA-wide sync
        constexpr float inv = 1.f / (float)HW;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mean = sums[c] * inv;
#pragma unroll
            for (int m = 0; m < NPX; ++m) v[m][c] = clamp01((v[m][c] - mean) * sp.fc + mean);
        }
        if (ord == 0) {
#pragma unroll
            for (int m = 0; m < NPX; ++m) hsv_jitter(v[m][0], v[m][1], v[m][2], hshift, sp.fs, sp.fv);
        }
    }
#pragma unroll
    for (int m = 0; m < NPX; ++m) {
        if (sp.gray_on != 0.f) {
            const float l = 0.299f * v[m][0] + 0.587f * v[m][1] + 0.114f * v[m][2];
            v[m][0] = l; v[m][1] = l; v[m][2] = l;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) __stcs(yb + c * HW + m * S, v[m][c]);
    }
}

// The two halves of block_sum3_256 around ONE CTA barrier that the caller places (and shares with its other hand-offs):
// `partial` leaves 3 x 8 warp sums in `scratch` (24 floats), `total` folds them after the barrier.
__device__ __forceinline__ void block_sum3_partial(const float (&v)[3], float* scratch) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hi = (lane & 16) != 0;
    const float r0 = __shfl_xor_sync(full, hi ? v[0] : v[2], 16);
    const float r1 = __shfl_xor_sync(full, hi ? v[1] : 0.f, 16);
    float a = (hi ? v[2] : v[0]) + r0;
    const float b = (hi ? 0.f : v[1]) + r1;
    const bool q = (lane & 8) != 0;
    a = (q ? b : a) + __shfl_xor_sync(full, q ? a : b, 8);
    a += __shfl_xor_sync(full, a, 4);
    a += __shfl_xor_sync(full, a, 2);
    a += __shfl_xor_sync(full, a, 1);
    if ((lane & 7) == 0 && lane < 24) scratch[(lane >> 3) * 8 + warp] = a;      // [value][warp]
}

__device__ __forceinline__ void block_sum3_total(float (&v)[3], const float* scratch) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float t = (lane < 24) ? scratch[lane] : 0.f;
    t += __shfl_xor_sync(full, t, 4);
    t += __shfl_xor_sync(full, t, 2);
    t += __shfl_xor_sync(full, t, 1);
    v[0] = __shfl_sync(full, t, 0);
    v[1] = __shfl_sync(full, t, 8);
    v[2] = __shfl_sync(full, t, 16);
}

// Forward, square S x S images with S a multiple of 32 (the CIFAR / 64x64 cases): COLUMN mapping.  Lane = output
// column, each thread walks NPX = S*S/256 consecutive output rows.  The bilinear gather then reads, per warp
// instruction, 32 (almost always distinct) columns of ONE source row: no shared-memory bank conflicts, where the
// quad mapping above put four source rows - all hitting the same banks, row stride S floats = 0 mod 32 - into
// each warp instruction (4-way conflicts on every gather; profiles/prof_r1_augment.md: l1tex 81 %, mio_throttle).
// Column taps are loaded once per thread, row taps are one broadcast LDS.128 per row; stores are 128 B per warp.
//
// Software pipeline with ONE CTA barrier per image (round 1 had four: tap tables, image arrival, the contrast mean's
// block reduction, end of iteration - ncu: 1.3 barrier stalls per issued instruction at 50 % warp occupancy):
//   * tap tables, reduction scratch and image buffers are double-buffered by iteration parity;
//   * while image `it` is gathered, the tap tables of image `it + 1` are computed (the five parameters they need were
//     requested at the top of the iteration) and the warp partials of the contrast sums are written;
//   * the single __syncthreads then publishes the next tap tables and the partial sums AND certifies that nobody reads
//     the current image buffer any more, so thread 0 refills it with image `it + 2` right behind the barrier (two bulk
//     copies stay in flight);
//   * the per-image parameters of the current image are requested at the top and first needed after the gather.
template <int S, int OCC>          // OCC = resident CTAs per SM the register allocation is held to (32x32: 6 -> 40 registers)
__global__ void __launch_bounds__(kMaxThreads, OCC)
augment_simclr_fwd_cols_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params,
                               int B, int order) {
    extern __shared__ __align__(128) float smem[];
    constexpr int HW = S * S, NPX = HW / kMaxThreads;
    float* red = smem + 6 * HW;                                   // [2][32]
    float4* xtap = reinterpret_cast<float4*>(red + 96);           // [2][S] {i0, i1 (int bits), w0, w1}, flip folded in
    float4* ytap = xtap + 2 * S;                                  // [2][S] {i0*S, i1*S (int bits), w0, w1}
    uint64_t* bars = reinterpret_cast<uint64_t*>(ytap + 2 * S);
    constexpr uint32_t img_bytes = (uint32_t)(3 * HW * sizeof(float));
    const int j = threadIdx.x % S, i_first = (threadIdx.x / S) * NPX;
    const int G = gridDim.x;

    auto write_taps = [&](int par, float sx, float sy, float bx, float by, float flip) {
        const int e = threadIdx.x;
        if (e < S) {
            const Tap a = axis_tap((flip < 0.f) ? (S - 1 - e) : e, S, sx, bx);
            xtap[par * S + e] = make_float4(__int_as_float(a.i0), __int_as_float(a.i1), a.w0, a.w1);
        } else if (e < 2 * S) {
            const Tap a = axis_tap(e - S, S, sy, by);
            ytap[par * S + e - S] = make_float4(__int_as_float(a.i0 * S), __int_as_float(a.i1 * S), a.w0, a.w1);
        }
    };

    if (threadIdx.x == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    int b = blockIdx.x;
    if (b >= B) return;
    if (threadIdx.x == 0) {
        bar_expect_tx(&bars[0], img_bytes);
        bulk_load(smem, x + (size_t)b * 3 * HW, img_bytes, &bars[0]);
        if (b + G < B) {
            bar_expect_tx(&bars[1], img_bytes);
            bulk_load(smem + 3 * HW, x + (size_t)(b + G) * 3 * HW, img_bytes, &bars[1]);
        }
    }
    if (threadIdx.x < 2 * S)
        write_taps(0, __ldg(params + 0 * B + b), __ldg(params + 1 * B + b), __ldg(params + 2 * B + b),
                   __ldg(params + 3 * B + b), __ldg(params + 4 * B + b));
    __syncthreads();

    for (int it = 0; b < B; b += G, ++it) {
        const int par = it & 1;
        const int nb = b + G;
        // requests first: this image's colour parameters, and (tap writers) the crop parameters of the next image
        const float cj_on = __ldg(params + 5 * B + b), fc = __ldg(params + 6 * B + b), fh = __ldg(params + 7 * B + b);
        const float fs = __ldg(params + 8 * B + b), fv = __ldg(params + 9 * B + b), gray_on = __ldg(params + 10 * B + b);
        const int ord = resolve_order(params, B, b, order);
        float nsx = 1.f, nsy = 1.f, nbx = 0.f, nby = 0.f, nflip = 1.f;
        const bool next_taps = nb < B && threadIdx.x < 2 * S;
        if (next_taps) {
            nsx = __ldg(params + 0 * B + nb); nsy = __ldg(params + 1 * B + nb); nbx = __ldg(params + 2 * B + nb);
            nby = __ldg(params + 3 * B + nb); nflip = __ldg(params + 4 * B + nb);
        }
        bar_wait(&bars[par], (uint32_t)(it >> 1) & 1u);
        const float* xs = smem + par * 3 * HW;

        // ---- gather (crop + flip) into registers
        float v[NPX][3];
        {
            const float4 tx = xtap[par * S + j];
            const int x0 = __float_as_int(tx.x), x1 = __float_as_int(tx.y);
#pragma unroll
            for (int m = 0; m < NPX; ++m) {
                const float4 ty = ytap[par * S + i_first + m];
                const float* r0 = xs + __float_as_int(ty.x);
                const float* r1 = xs + __float_as_int(ty.y);
                const float w00 = tx.z * ty.z, w01 = tx.w * ty.z, w10 = tx.z * ty.w, w11 = tx.w * ty.w;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[m][c] = r0[c * HW + x0] * w00 + r0[c * HW + x1] * w01 + r1[c * HW + x0] * w10 +
                              r1[c * HW + x1] * w11;
            }
        }
        const float hshift = fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
        float sums[3] = {0.f, 0.f, 0.f};
        if (cj_on != 0.f) {          // uniform across the CTA
            if (ord == 1) {
#pragma unroll
                for (int m = 0; m < NPX; ++m) hsv_jitter(v[m][0], v[m][1], v[m][2], hshift, fs, fv);
            }
#pragma unroll
            for (int m = 0; m < NPX; ++m)
#pragma unroll
                for (int c = 0; c < 3; ++c) sums[c] += v[m][c];
            block_sum3_partial(sums, red + par * 32);
        }
        if (next_taps) write_taps(par ^ 1, nsx, nsy, nbx, nby, nflip);
        __syncthreads();          // the only CTA barrier of the iteration (see the header comment)
        if (threadIdx.x == 0 && nb + G < B) {
            bar_expect_tx(&bars[par], img_bytes);
            bulk_load(smem + par * 3 * HW, x + (size_t)(nb + G) * 3 * HW, img_bytes, &bars[par]);
        }
        if (cj_on != 0.f) {
            block_sum3_total(sums, red + par * 32);
            constexpr float inv = 1.f / (float)HW;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float mean = sums[c] * inv;
#pragma unroll
                for (int m = 0; m < NPX; ++m) v[m][c] = clamp01((v[m][c] - mean) * fc + mean);
            }
            if (ord == 0) {
#pragma unroll
                for (int m = 0; m < NPX; ++m) hsv_jitter(v[m][0], v[m][1], v[m][2], hshift, fs, fv);
            }
        }
        float* yb = y + (size_t)b * 3 * HW + i_first * S + j;
#pragma unroll
        for (int m = 0; m < NPX; ++m) {
            if (gray_on != 0.f) {
                const float l = 0.299f * v[m][0] + 0.587f * v[m][1] + 0.114f * v[m][2];
                v[m][0] = l; v[m][1] = l; v[m][2] = l;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) __stcs(yb + c * HW + m * S, v[m][c]);
        }
    }
}

// Second build of the same pipeline - the default since it was timed on the B200 (0.81 against 0.65 of the copy
// bandwidth at 32 x 32, 0.71 against 0.60 at 64 x 64: profiles/augment_ab_r2.json); CB200_AUGMENT_V=1 selects the first.
// SASS of the kernel above (profiles/prof_r2_augment.md: issue-bound, 4 893 warp instructions per image) spends ~100 of its ~625 instructions per thread and image on the per-image
// parameters: every thread forms eleven 64-bit addresses `params + k*B + b` and loads the same eleven values, and ~24 on
// turning tap indices into shared-memory addresses (buffer parity * 3HW + row + column, then scale + base).  Here
//   * twelve lanes of the last warp fetch the parameter column of image `it + 2` (one LDG each) and park it in a
//     4-slot shared-memory ring before the iteration's barrier; everybody else reads its colour parameters with two
//     broadcast LDS.128, the tap writers their crop parameters with two more;
//   * the tap tables hold BYTE offsets, and the row entries already contain the offset of the ring buffer the image
//     will arrive in, so a bilinear tap address is one three-input add (row + column + shared window base).
// The arithmetic on the pixels is instruction for instruction the one above: outputs are bit-identical (test).  HV picks
// the build of the HSV stage (see hsv_jitter; 2 is still bit-identical); the grayscale branch (taken by 20 % of the
// images) is a real branch around its own store sequence instead of 20 predicated instructions for everybody.
template <int S, int OCC, int HV>
__global__ void __launch_bounds__(kMaxThreads, OCC)
augment_simclr_fwd_cols2_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params,
                                int B, int order) {
    extern __shared__ __align__(128) float smem[];
    constexpr int HW = S * S, NPX = HW / kMaxThreads;
    constexpr int kLoader0 = kMaxThreads - 32;                   // first lane of the parameter-loading warp
    float* red = smem + 6 * HW;                                   // [2][32]
    float4* xtap = reinterpret_cast<float4*>(red + 96);           // [2][S] {4*i0, 4*i1 (int bits), w0, w1}, flip folded in
    float4* ytap = xtap + 2 * S;                                  // [2][S] {buffer + 4*S*i0, buffer + 4*S*i1 (bytes), w0, w1}
    float* pslot = reinterpret_cast<float*>(ytap + 2 * S);        // [4][16] parameter columns: sx sy bx by | flip cj fc fh | fs fv gray order
    uint64_t* bars = reinterpret_cast<uint64_t*>(pslot + 64);
    constexpr uint32_t img_bytes = (uint32_t)(3 * HW * sizeof(float));
    const int j = threadIdx.x % S, i_first = (threadIdx.x / S) * NPX;
    const int G = gridDim.x;
    const int pl = (int)threadIdx.x - kLoader0;                   // parameter row this thread fetches (0..11), else none
    const bool loader = pl >= 0 && pl < 12;

    auto fetch_param = [&](int img) -> float {                    // loader lanes only
        if (pl == 11) return order >= 0 ? (float)order : (__ldg(params + (size_t)11 * B + img) != 0.f ? 1.f : 0.f);
        return __ldg(params + (size_t)pl * B + img);
    };
    auto write_taps = [&](int par, const float4 crop, float flip) {      // crop = {sx, sy, bx, by}
        const int e = threadIdx.x;
        if (e < S) {
            const Tap a = axis_tap((flip < 0.f) ? (S - 1 - e) : e, S, crop.x, crop.z);
            xtap[par * S + e] = make_float4(__int_as_float(a.i0 * 4), __int_as_float(a.i1 * 4), a.w0, a.w1);
        } else if (e < 2 * S) {
            const Tap a = axis_tap(e - S, S, crop.y, crop.w);
            const int buf = par * (int)img_bytes;
            ytap[par * S + e - S] = make_float4(__int_as_float(buf + a.i0 * (S * 4)), __int_as_float(buf + a.i1 * (S * 4)),
                                                a.w0, a.w1);
        }
    };

    if (threadIdx.x == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        fence_barrier_init();
    }
    int b = blockIdx.x;
    if (loader && b < B) {
        pslot[pl] = fetch_param(b);
        if (b + G < B) pslot[16 + pl] = fetch_param(b + G);
    }
    __syncthreads();
    if (b >= B) return;
    if (threadIdx.x == 0) {
        bar_expect_tx(&bars[0], img_bytes);
        bulk_load(smem, x + (size_t)b * 3 * HW, img_bytes, &bars[0]);
        if (b + G < B) {
            bar_expect_tx(&bars[1], img_bytes);
            bulk_load(smem + 3 * HW, x + (size_t)(b + G) * 3 * HW, img_bytes, &bars[1]);
        }
    }
    if (threadIdx.x < 2 * S) write_taps(0, *reinterpret_cast<const float4*>(pslot), pslot[4]);
    __syncthreads();

    const char* sbytes = reinterpret_cast<const char*>(smem);
    float* yb = y + (size_t)b * 3 * HW + i_first * S + j;
    for (int it = 0; b < B; b += G, ++it, yb += (size_t)G * 3 * HW) {
        const int par = it & 1;
        const int nb = b + G;
        // requests first: (loader lanes) the parameter column of image it + 2
        float pv = 0.f;
        const bool fetch = loader && nb + G < B;
        if (fetch) pv = fetch_param(nb + G);
        // this image's colour parameters, (tap writers) the crop parameters of the next image: shared-memory broadcasts
        const float* ps = pslot + (it & 3) * 16;
        const float4 c0 = *reinterpret_cast<const float4*>(ps + 4);      // flip cj_on fc fh
        const float4 c1 = *reinterpret_cast<const float4*>(ps + 8);      // fs fv gray_on order
        const float cj_on = c0.y, fc = c0.z, fs = c1.x, fv = c1.y, gray_on = c1.z;
        const int ord = c1.w != 0.f ? 1 : 0;
        const float hshift = c0.w * (255.f / 360.f);   // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
        const bool next_taps = nb < B && threadIdx.x < 2 * S;
        bar_wait(&bars[par], (uint32_t)(it >> 1) & 1u);

        // ---- gather (crop + flip) into registers
        float v[NPX][3];
        {
            const float4 tx = xtap[par * S + j];
            const int x0 = __float_as_int(tx.x), x1 = __float_as_int(tx.y);
            const float4* yt = ytap + par * S + i_first;
#pragma unroll
            for (int m = 0; m < NPX; ++m) {
                const float4 ty = yt[m];
                const int y0 = __float_as_int(ty.x), y1 = __float_as_int(ty.y);
                const float* p00 = reinterpret_cast<const float*>(sbytes + (y0 + x0));
                const float* p01 = reinterpret_cast<const float*>(sbytes + (y0 + x1));
                const float* p10 = reinterpret_cast<const float*>(sbytes + (y1 + x0));
                const float* p11 = reinterpret_cast<const float*>(sbytes + (y1 + x1));
                const float w00 = tx.z * ty.z, w01 = tx.w * ty.z, w10 = tx.z * ty.w, w11 = tx.w * ty.w;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[m][c] = p00[c * HW] * w00 + p01[c * HW] * w01 + p10[c * HW] * w10 + p11[c * HW] * w11;
            }
        }
        float sums[3] = {0.f, 0.f, 0.f};
        if (cj_on != 0.f) {          // uniform across the CTA
            if (ord == 1) {
#pragma unroll
                for (int m = 0; m < NPX; ++m) hsv_jitter<HV>(v[m][0], v[m][1], v[m][2], hshift, fs, fv);
            }
#pragma unroll
            for (int m = 0; m < NPX; ++m)
#pragma unroll
                for (int c = 0; c < 3; ++c) sums[c] += v[m][c];
            block_sum3_partial(sums, red + par * 32);
        }
        if (next_taps) {
            const float* pn = pslot + ((it + 1) & 3) * 16;
            write_taps(par ^ 1, *reinterpret_cast<const float4*>(pn), pn[4]);
        }
        if (fetch) pslot[((it + 2) & 3) * 16 + pl] = pv;
        __syncthreads();          // the only CTA barrier of the iteration (see the header comment of the kernel above)
        if (threadIdx.x == 0 && nb + G < B) {
            bar_expect_tx(&bars[par], img_bytes);
            bulk_load(smem + par * 3 * HW, x + (size_t)(nb + G) * 3 * HW, img_bytes, &bars[par]);
        }
        if (cj_on != 0.f) {
            block_sum3_total(sums, red + par * 32);
            constexpr float inv = 1.f / (float)HW;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float mean = sums[c] * inv;
#pragma unroll
                for (int m = 0; m < NPX; ++m) v[m][c] = clamp01((v[m][c] - mean) * fc + mean);
            }
            if (ord == 0) {
#pragma unroll
                for (int m = 0; m < NPX; ++m) hsv_jitter<HV>(v[m][0], v[m][1], v[m][2], hshift, fs, fv);
            }
        }
        if (gray_on != 0.f) {        // uniform across the CTA
#pragma unroll
            for (int m = 0; m < NPX; ++m) {
                const float l = 0.299f * v[m][0] + 0.587f * v[m][1] + 0.114f * v[m][2];
#pragma unroll
                for (int c = 0; c < 3; ++c) __stcs(yb + c * HW + m * S, l);
            }
        } else {
#pragma unroll
            for (int m = 0; m < NPX; ++m)
#pragma unroll
                for (int c = 0; c < 3; ++c) __stcs(yb + c * HW + m * S, v[m][c]);
        }
    }
}

// Row f3: `ToTensor` folded into the augmentation.  The reference converts the dataset's uint8 images to
// fp32 on the host (datasets.py:10-21 -> torchvision to_tensor: byte / 255), copies 4 B/element to the device and
// concatenates [x, x, G(z)] (training/gan/contrad.py:38-40) before augmenting.  Here view b of the launch reads
//     b <  n_u8_views : uint8 image (b mod n_u8)   (1 B/element of HBM traffic, several views may share a source)
//     b >= n_u8_views : fp32 image (b - n_u8_views)
// so the D-step batch cat[x, x, G(z)] is never materialised.  Bytes are expanded through a 256-entry table of
// correctly rounded k / 255 (bit-identical to ToTensor) into a separate fp32 buffer: the prefetch ring is only ever
// written by the async proxy, as in the fp32 kernel above.
template <int S>
__global__ void __launch_bounds__(kMaxThreads)
augment_simclr_fwd_mixed_cols_kernel(const uint8_t* __restrict__ xu, int n_u8, int n_u8_views,
                                     const float* __restrict__ xf, float* __restrict__ y,
                                     const float* __restrict__ params, int B, int order) {
    extern __shared__ __align__(128) float smem[];
    constexpr int HW = S * S, NPX = HW / kMaxThreads;
    float* conv = smem + 6 * HW;                                  // [3*HW] fp32 copy of a uint8 source
    float* lut = conv + 3 * HW;                                   // [256]
    float* red = lut + 256;                                       // [3*32]
    float4* xtap = reinterpret_cast<float4*>(red + 96);
    float4* ytap = xtap + S;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ytap + S);
    const int j = threadIdx.x % S, i_first = (threadIdx.x / S) * NPX;

    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);      // blockDim.x == 256
    if (threadIdx.x == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto prefetch = [&](int view, int slot) {                     // thread 0 only
        const bool bytes_src = view < n_u8_views;
        const uint32_t nbytes = bytes_src ? (uint32_t)(3 * HW) : (uint32_t)(3 * HW * sizeof(float));
        const void* src = bytes_src ? static_cast<const void*>(xu + (size_t)(view % n_u8) * 3 * HW)
                                    : static_cast<const void*>(xf + (size_t)(view - n_u8_views) * 3 * HW);
        bar_expect_tx(&bars[slot], nbytes);
        bulk_load(smem + slot * 3 * HW, src, nbytes, &bars[slot]);
    };
    int b = blockIdx.x;
    if (threadIdx.x == 0 && b < B) prefetch(b, 0);
    for (int it = 0; b < B; b += gridDim.x, ++it) {
        const int buf = it & 1;
        const int nb = b + gridDim.x;
        if (threadIdx.x == 0 && nb < B) prefetch(nb, buf ^ 1);
        const SampleParams sp = load_params(params, B, b);
        const int ord = resolve_order(params, B, b, order);
        const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
        if (threadIdx.x < 2 * S) {
            const int e = threadIdx.x;
            if (e < S) {
                const Tap a = axis_tap((sp.flip < 0.f) ? (S - 1 - e) : e, S, sp.sx, sp.bx);
                xtap[e] = make_float4(__int_as_float(a.i0), __int_as_float(a.i1), a.w0, a.w1);
            } else {
                const Tap a = axis_tap(e - S, S, sp.sy, sp.by);
                ytap[e - S] = make_float4(__int_as_float(a.i0 * S), __int_as_float(a.i1 * S), a.w0, a.w1);
            }
        }
        bar_wait(&bars[buf], (uint32_t)(it >> 1) & 1u);
        const float* xs = smem + buf * 3 * HW;
        if (b < n_u8_views) {                                     // uniform across the CTA
            const uint32_t* packed = reinterpret_cast<const uint32_t*>(xs);
            for (int w = threadIdx.x; w < 3 * HW / 4; w += kMaxThreads) {      // conflict-free LDS.32 / STS.128
                const uint32_t pk = packed[w];
                reinterpret_cast<float4*>(conv)[w] =
                    make_float4(lut[pk & 255u], lut[(pk >> 8) & 255u], lut[(pk >> 16) & 255u], lut[pk >> 24]);
            }
            xs = conv;
        }
        __syncthreads();
        cols_process<S>(xs, xtap, ytap, red, sp, ord, hshift, y + (size_t)b * 3 * HW + i_first * S + j, j, i_first);
        __syncthreads();      // tap tables / scratch / conversion buffer / this ring slot are reused by the next iteration
    }
}

template <int QPT>
__global__ void __launch_bounds__(kMaxThreads)
augment_simclr_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                          const float* __restrict__ params, int B, int H, int W, int order) {
    extern __shared__ __align__(128) float smem[];
    const int HW = H * W, Wq = W >> 2, nquads = HW >> 2;
    float* xs = smem;                  // [3*HW] forward image
    float* gs = smem + 3 * HW;         // [3*HW] dx accumulator
    float* red = smem + 6 * HW;        // [3*32]
    const TapTables taps = carve_taps(red + 96, H, W);
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 96 + 4 * W + 4 * H);
    const int b = blockIdx.x;
    const SampleParams sp = load_params(params, B, b);
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    if (sp.cj_on != 0.f && threadIdx.x == 0) {     // x only feeds the clamp mask; one async bulk copy stages it
        bar_init(bar, 1);
        fence_barrier_init();
        bar_expect_tx(bar, (uint32_t)(3 * HW * sizeof(float)));
        bulk_load(xs, x + (size_t)b * 3 * HW, (uint32_t)(3 * HW * sizeof(float)), bar);
    }
    for (int e = threadIdx.x * 4; e < 3 * HW; e += blockDim.x * 4)
        *reinterpret_cast<float4*>(gs + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    fill_taps(taps, H, W, sp);
    __syncthreads();

    float f[QPT][3][4];   // forward value at the contrast input (crop+flip, optionally hsv)
    float g[QPT][3][4];   // gradient
    bool live[QPT];
    const float* dyb = dy + (size_t)b * 3 * HW;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        int quad = threadIdx.x + q * blockDim.x;
        live[q] = quad < nquads;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) { f[q][c][k] = 0.f; g[q][c][k] = 0.f; }
        if (!live[q]) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 t = ldg_stream4(dyb + c * HW + quad * 4);
            g[q][c][0] = t.x; g[q][c][1] = t.y; g[q][c][2] = t.z; g[q][c][3] = t.w;
        }
        if (sp.gray_on != 0.f) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float s = g[q][0][k] + g[q][1][k] + g[q][2][k];
                g[q][0][k] = 0.299f * s; g[q][1][k] = 0.587f * s; g[q][2][k] = 0.114f * s;
            }
        }
    }
    if (sp.cj_on != 0.f) {
        bar_wait(bar, 0);     // barrier init by thread 0 is ordered before this by the __syncthreads above
#pragma unroll
        for (int q = 0; q < QPT; ++q) {
            if (!live[q]) continue;
            int quad = threadIdx.x + q * blockDim.x;
            gather_quad(xs, H, W, quad / Wq, (quad % Wq) * 4, taps, f[q]);
            if (ord == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) hsv_jitter(f[q][0][k], f[q][1][k], f[q][2][k], hshift, sp.fs, sp.fv);
            }
        }
        float sums[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < QPT; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) sums[c] += (f[q][c][0] + f[q][c][1]) + (f[q][c][2] + f[q][c][3]);
        block_sum<3>(sums, red);
        const float inv = 1.f / (float)HW;
        float gsum[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float m = sums[c] * inv;
#pragma unroll
            for (int q = 0; q < QPT; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float u = (f[q][c][k] - m) * sp.fc + m;
                    float gg = (u >= 0.f && u <= 1.f) ? g[q][c][k] : 0.f;   // clamp backward (inclusive)
                    g[q][c][k] = gg;
                    gsum[c] += gg;
                }
        }
        block_sum<3>(gsum, red);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float gm = gsum[c] * inv * (1.f - sp.fc);
#pragma unroll
            for (int q = 0; q < QPT; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) g[q][c][k] = sp.fc * g[q][c][k] + gm;
        }
    }
    // transposed crop: scatter into the shared accumulator
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        if (!live[q]) continue;
        int quad = threadIdx.x + q * blockDim.x;
        int i = quad / Wq, j0 = (quad % Wq) * 4;
        const float wy0 = taps.yw0[i], wy1 = taps.yw1[i];
        const int y0 = taps.yi0[i] * W, y1 = taps.yi1[i] * W;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = j0 + k;
            const int x0 = taps.xi0[j], x1 = taps.xi1[j];
            const float wx0 = taps.xw0[j], wx1 = taps.xw1[j];
            float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float gv = g[q][c][k];
                float* base = gs + c * HW;
                atomicAdd(base + y0 + x0, gv * w00);
                if (w01 != 0.f) atomicAdd(base + y0 + x1, gv * w01);
                if (w10 != 0.f) atomicAdd(base + y1 + x0, gv * w10);
                if (w11 != 0.f) atomicAdd(base + y1 + x1, gv * w11);
            }
        }
    }
    __syncthreads();
    float* dxb = dx + (size_t)b * 3 * HW;
    for (int e = threadIdx.x * 4; e < 3 * HW; e += blockDim.x * 4)
        stg_stream4(dxb + e, *reinterpret_cast<const float4*>(gs + e));
}

// ------------------------------------------------------------------------------------------------
// Large-image path (H*W > 4096: the 512x512 StyleGAN2 configs).  The image no longer fits one CTA's shared memory, so
// the chain runs from global memory, one thread per output pixel, and the two per-image reductions the chain contains
// (the per-channel mean of `adjust_contrast` and, in the backward pass, the mean of the masked gradient) become
// separate reduction launches that write [B,3] scalars:
//     fwd: mean kernel (skipped per image when colour jitter is off) -> apply kernel
//     bwd: masked-gradient-sum kernel -> scatter kernel (transposed bilinear crop with global fp32 atomics)
// The crop / hsv arithmetic is recomputed in each pass instead of being stored: x is read through L2 (the 4 bilinear
// taps of neighbouring pixels overlap), y / dy / dx move once.
struct PixelTaps {
    int o00, o01, o10, o11;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ PixelTaps pixel_taps(int i, int j, int H, int W, const SampleParams& sp) {
    const Tap ty = axis_tap(i, H, sp.sy, sp.by);
    const Tap tx = axis_tap((sp.flip < 0.f) ? (W - 1 - j) : j, W, sp.sx, sp.bx);
    PixelTaps t;
    t.o00 = ty.i0 * W + tx.i0; t.o01 = ty.i0 * W + tx.i1; t.o10 = ty.i1 * W + tx.i0; t.o11 = ty.i1 * W + tx.i1;
    t.w00 = tx.w0 * ty.w0; t.w01 = tx.w1 * ty.w0; t.w10 = tx.w0 * ty.w1; t.w11 = tx.w1 * ty.w1;
    return t;
}

__device__ __forceinline__ void gather_pixel(const float* __restrict__ xb, int HW, const PixelTaps& t, float (&v)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* xc = xb + (size_t)c * HW;
        v[c] = __ldg(xc + t.o00) * t.w00 + __ldg(xc + t.o01) * t.w01 + __ldg(xc + t.o10) * t.w10 + __ldg(xc + t.o11) * t.w11;
    }
}

// The same gather from a uint8 image; `lut` (shared memory) holds the correctly rounded k / 255 of ToTensor.
__device__ __forceinline__ void gather_pixel(const uint8_t* __restrict__ xb, const float* lut, int HW, const PixelTaps& t,
                                             float (&v)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint8_t* xc = xb + (size_t)c * HW;
        v[c] = lut[__ldg(xc + t.o00)] * t.w00 + lut[__ldg(xc + t.o01)] * t.w01 + lut[__ldg(xc + t.o10)] * t.w10 +
               lut[__ldg(xc + t.o11)] * t.w11;
    }
}

// means[b,3] += per-channel mean of the contrast input (crop+flip, then hsv when the order is [hsv, contrast])
__global__ void __launch_bounds__(kMaxThreads)
augment_large_mean_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ means, int B,
                          int H, int W, int order) {
    __shared__ float red[96];
    const int b = blockIdx.y;
    const SampleParams sp = load_params(params, B, b);
    if (sp.cj_on == 0.f) return;                                   // uniform per CTA
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    const int HW = H * W;
    const float* xb = x + (size_t)b * 3 * HW;
    float sums[3] = {0.f, 0.f, 0.f};
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float v[3];
        gather_pixel(xb, HW, pixel_taps(pix / W, pix % W, H, W, sp), v);
        if (ord == 1) hsv_jitter(v[0], v[1], v[2], hshift, sp.fs, sp.fv);
        sums[0] += v[0]; sums[1] += v[1]; sums[2] += v[2];
    }
    block_sum<3>(sums, red);
    if (threadIdx.x == 0) {
        const float inv = 1.f / (float)HW;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(means + b * 3 + c, sums[c] * inv);
    }
}

__global__ void __launch_bounds__(kMaxThreads)
augment_large_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params,
                           const float* __restrict__ means, int B, int H, int W, int order) {
    const int b = blockIdx.y;
    const SampleParams sp = load_params(params, B, b);
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    const int HW = H * W;
    const float* xb = x + (size_t)b * 3 * HW;
    float* yb = y + (size_t)b * 3 * HW;
    float m[3] = {0.f, 0.f, 0.f};
    if (sp.cj_on != 0.f) { m[0] = __ldg(means + b * 3); m[1] = __ldg(means + b * 3 + 1); m[2] = __ldg(means + b * 3 + 2); }
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float v[3];
        gather_pixel(xb, HW, pixel_taps(pix / W, pix % W, H, W, sp), v);
        if (sp.cj_on != 0.f) {
            if (ord == 1) hsv_jitter(v[0], v[1], v[2], hshift, sp.fs, sp.fv);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = clamp01((v[c] - m[c]) * sp.fc + m[c]);
            if (ord == 0) hsv_jitter(v[0], v[1], v[2], hshift, sp.fs, sp.fv);
        }
        if (sp.gray_on != 0.f) {
            const float l = 0.299f * v[0] + 0.587f * v[1] + 0.114f * v[2];
            v[0] = l; v[1] = l; v[2] = l;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) __stcs(yb + (size_t)c * HW + pix, v[c]);
    }
}

// Mixed-source variants of the two forward kernels (row f3, see augment_simclr_fwd_mixed_cols_kernel): view b reads the
// uint8 image (b mod n_u8) when b < n_u8_views, else the fp32 image (b - n_u8_views).  The source type is uniform per CTA.
__global__ void __launch_bounds__(kMaxThreads)
augment_large_mean_mixed_kernel(const uint8_t* __restrict__ xu, int n_u8, int n_u8_views, const float* __restrict__ xf,
                                const float* __restrict__ params, float* __restrict__ means, int B, int H, int W, int order) {
    __shared__ float red[96];
    __shared__ float lut[256];
    const int b = blockIdx.y;
    const SampleParams sp = load_params(params, B, b);
    if (sp.cj_on == 0.f) return;                                   // uniform per CTA
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);       // blockDim.x == 256
    __syncthreads();
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    const int HW = H * W;
    const bool bytes_src = b < n_u8_views;
    const uint8_t* xub = xu + (size_t)(bytes_src ? b % n_u8 : 0) * 3 * HW;
    const float* xfb = xf + (size_t)(bytes_src ? 0 : b - n_u8_views) * 3 * HW;
    float sums[3] = {0.f, 0.f, 0.f};
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float v[3];
        const PixelTaps t = pixel_taps(pix / W, pix % W, H, W, sp);
        if (bytes_src) gather_pixel(xub, lut, HW, t, v); else gather_pixel(xfb, HW, t, v);
        if (ord == 1) hsv_jitter(v[0], v[1], v[2], hshift, sp.fs, sp.fv);
        sums[0] += v[0]; sums[1] += v[1]; sums[2] += v[2];
    }
    block_sum<3>(sums, red);
    if (threadIdx.x == 0) {
        const float inv = 1.f / (float)HW;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(means + b * 3 + c, sums[c] * inv);
    }
}

__global__ void __launch_bounds__(kMaxThreads)
augment_large_apply_mixed_kernel(const uint8_t* __restrict__ xu, int n_u8, int n_u8_views, const float* __restrict__ xf,
                                 float* __restrict__ y, const float* __restrict__ params, const float* __restrict__ means,
                                 int B, int H, int W, int order) {
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);       // blockDim.x == 256
    __syncthreads();
    const int b = blockIdx.y;
    const SampleParams sp = load_params(params, B, b);
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    const int HW = H * W;
    const bool bytes_src = b < n_u8_views;
    const uint8_t* xub = xu + (size_t)(bytes_src ? b % n_u8 : 0) * 3 * HW;
    const float* xfb = xf + (size_t)(bytes_src ? 0 : b - n_u8_views) * 3 * HW;
    float* yb = y + (size_t)b * 3 * HW;
    float m[3] = {0.f, 0.f, 0.f};
    if (sp.cj_on != 0.f) { m[0] = __ldg(means + b * 3); m[1] = __ldg(means + b * 3 + 1); m[2] = __ldg(means + b * 3 + 2); }
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float v[3];
        const PixelTaps t = pixel_taps(pix / W, pix % W, H, W, sp);
        if (bytes_src) gather_pixel(xub, lut, HW, t, v); else gather_pixel(xfb, HW, t, v);
        if (sp.cj_on != 0.f) {
            if (ord == 1) hsv_jitter(v[0], v[1], v[2], hshift, sp.fs, sp.fv);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = clamp01((v[c] - m[c]) * sp.fc + m[c]);
            if (ord == 0) hsv_jitter(v[0], v[1], v[2], hshift, sp.fs, sp.fv);
        }
        if (sp.gray_on != 0.f) {
            const float l = 0.299f * v[0] + 0.587f * v[1] + 0.114f * v[2];
            v[0] = l; v[1] = l; v[2] = l;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) __stcs(yb + (size_t)c * HW + pix, v[c]);
    }
}

// Gradient at the contrast OUTPUT for one pixel: gray backward, hsv straight-through, clamp mask (needs the forward
// value at the contrast input, recomputed here).  Returns false when the image has no colour jitter.
__device__ __forceinline__ void masked_grad(const float* __restrict__ xb, const float* __restrict__ dyb, int HW, int pix,
                                            const PixelTaps& t, const SampleParams& sp, int ord, float hshift,
                                            const float* m, float (&g)[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = __ldcs(dyb + (size_t)c * HW + pix);
    if (sp.gray_on != 0.f) {
        const float s = g[0] + g[1] + g[2];
        g[0] = 0.299f * s; g[1] = 0.587f * s; g[2] = 0.114f * s;
    }
    if (sp.cj_on != 0.f) {
        float f[3];
        gather_pixel(xb, HW, t, f);
        if (ord == 1) hsv_jitter(f[0], f[1], f[2], hshift, sp.fs, sp.fv);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float u = (f[c] - m[c]) * sp.fc + m[c];
            if (!(u >= 0.f && u <= 1.f)) g[c] = 0.f;               // clamp backward (inclusive)
        }
    }
}

// gsums[b,3] += sum over pixels of the masked gradient (only images with colour jitter)
__global__ void __launch_bounds__(kMaxThreads)
augment_large_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ params,
                                const float* __restrict__ means, float* __restrict__ gsums, int B, int H, int W, int order) {
    __shared__ float red[96];
    const int b = blockIdx.y;
    const SampleParams sp = load_params(params, B, b);
    if (sp.cj_on == 0.f) return;
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    const int HW = H * W;
    const float* xb = x + (size_t)b * 3 * HW;
    const float* dyb = dy + (size_t)b * 3 * HW;
    const float m[3] = {__ldg(means + b * 3), __ldg(means + b * 3 + 1), __ldg(means + b * 3 + 2)};
    float sums[3] = {0.f, 0.f, 0.f};
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float g[3];
        masked_grad(xb, dyb, HW, pix, pixel_taps(pix / W, pix % W, H, W, sp), sp, ord, hshift, m, g);
        sums[0] += g[0]; sums[1] += g[1]; sums[2] += g[2];
    }
    block_sum<3>(sums, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(gsums + b * 3 + c, sums[c]);
    }
}

// dx (zeroed by the host) += transposed crop of the gradient at the crop output
__global__ void __launch_bounds__(kMaxThreads)
augment_large_bwd_scatter_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                 const float* __restrict__ params, const float* __restrict__ means,
                                 const float* __restrict__ gsums, int B, int H, int W, int order) {
    const int b = blockIdx.y;
    const SampleParams sp = load_params(params, B, b);
    const int ord = resolve_order(params, B, b, order);
    const float hshift = sp.fh * (255.f / 360.f);     // color_jitter.py:88, to 1 ulp (a true division costs ~8 issue slots)
    const int HW = H * W;
    const float* xb = x + (size_t)b * 3 * HW;
    const float* dyb = dy + (size_t)b * 3 * HW;
    float* dxb = dx + (size_t)b * 3 * HW;
    float m[3] = {0.f, 0.f, 0.f}, gm[3] = {0.f, 0.f, 0.f};
    if (sp.cj_on != 0.f) {
        const float inv = (1.f - sp.fc) / (float)HW;
#pragma unroll
        for (int c = 0; c < 3; ++c) { m[c] = __ldg(means + b * 3 + c); gm[c] = __ldg(gsums + b * 3 + c) * inv; }
    }
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        const PixelTaps t = pixel_taps(pix / W, pix % W, H, W, sp);
        float g[3];
        masked_grad(xb, dyb, HW, pix, t, sp, ord, hshift, m, g);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float gv = (sp.cj_on != 0.f) ? fmaf(sp.fc, g[c], gm[c]) : g[c];
            float* base = dxb + (size_t)c * HW;
            atomicAdd(base + t.o00, gv * t.w00);
            if (t.w01 != 0.f) atomicAdd(base + t.o01, gv * t.w01);
            if (t.w10 != 0.f) atomicAdd(base + t.o10, gv * t.w10);
            if (t.w11 != 0.f) atomicAdd(base + t.o11, gv * t.w11);
        }
    }
}

inline dim3 large_grid(int B, int H, int W) {
    long long blocks = ((long long)H * W + kMaxThreads - 1) / kMaxThreads;
    const long long cap = (148LL * 16 + B - 1) / B;                // about 16 resident CTAs per SM over the whole batch
    if (blocks > cap) blocks = cap < 1 ? 1 : cap;
    return dim3((unsigned)blocks, (unsigned)B);
}

struct LaunchShape {
    int threads, qpt;
};

bool pick_shape(int H, int W, LaunchShape* s) {
    int nquads = (H * W) / 4;
    int threads = nquads < kMaxThreads ? ((nquads + 31) / 32) * 32 : kMaxThreads;
    int qpt = (nquads + threads - 1) / threads;
    if (qpt > 4) return false;
    s->threads = threads;
    s->qpt = qpt == 3 ? 4 : qpt;
    return true;
}

}  // namespace

extern "C" int cb200_augment_simclr_fwd(const float* x, float* y, const float* params, int B, int H, int W,
                                        int order, void* stream) {
    CB200_CHECK_ARG(B >= 0 && H > 0 && W > 0 && (order >= -1 && order <= 1), "augment_fwd: bad shape/order");
    CB200_CHECK_ARG(W % 4 == 0 && (H * W) % 4 == 0, "augment_fwd: W must be a multiple of 4 (got %d)", W);
    CB200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                    "augment_fwd: x/y must be 16-byte aligned");
    if (B == 0) return CB200_OK;
    LaunchShape ls;
    CB200_CHECK_ARG(pick_shape(H, W, &ls),
                    "augment_fwd: images larger than 64x64 (%dx%d) need the tiled path (not built yet)", H, W);
    size_t smem = (size_t)(6 * H * W + 96 + 8 * W + 8 * H) * sizeof(float) + 2 * sizeof(uint64_t);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    int per_sm = (int)((200 * 1024) / smem);           // resident CTAs per SM by shared memory (<= 8 by threads)
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    const int grid = B < sm_count * per_sm ? B : sm_count * per_sm;
#define LAUNCH_FWD(Q, SZ)                                                                                      \
    do {                                                                                                       \
        cudaFuncSetAttribute(augment_simclr_fwd_kernel<Q, SZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        augment_simclr_fwd_kernel<Q, SZ><<<grid, ls.threads, smem, st>>>(x, y, params, B, H, W, order);        \
    } while (0)
    static const bool cols_ok = []() { const char* e = getenv("CB200_AUGMENT_COLS"); return !(e && e[0] == '0'); }();
    if (cols_ok && H == W && (H == 32 || H == 64)) {
        // 32x32: ncu showed the kernel occupancy-limited at 5 CTAs / SM by registers; the default build holds it to 40
        // registers (6 CTAs / SM, 4 bytes of spill); CB200_AUGMENT_OCC=5 selects the unconstrained build (A/B runs)
        static const int occ = []() { const char* e = getenv("CB200_AUGMENT_OCC"); return e ? atoi(e) : 6; }();
        // default build 2: parameters staged through shared memory + byte-offset tap tables (bit-identical outputs to build 1)
        // (read per call - one getenv - so that one process can compare the two builds)
        const char* ev = getenv("CB200_AUGMENT_V");
        const int variant = ev ? atoi(ev) : 2;
        if (variant == 2) {
            const size_t smem2 = smem + 64 * sizeof(float);
#define LAUNCH_COLS2(SZ, OC, HV, GRID)                                                                                   \
    do {                                                                                                                 \
        cudaFuncSetAttribute(augment_simclr_fwd_cols2_kernel<SZ, OC, HV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem2);                                                                                \
        augment_simclr_fwd_cols2_kernel<SZ, OC, HV><<<GRID, kMaxThreads, smem2, st>>>(x, y, params, B, order);           \
    } while (0)
            const int g6 = B < sm_count * 6 ? B : sm_count * 6;
            if (H == 32) LAUNCH_COLS2(32, 6, 2, g6);
            else LAUNCH_COLS2(64, 1, 2, grid);
#undef LAUNCH_COLS2
        } else if (H == 32 && occ >= 6) {
            const int g6 = B < sm_count * 6 ? B : sm_count * 6;
            cudaFuncSetAttribute(augment_simclr_fwd_cols_kernel<32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            augment_simclr_fwd_cols_kernel<32, 6><<<g6, kMaxThreads, smem, st>>>(x, y, params, B, order);
        } else if (H == 32) {
            const int g5 = B < sm_count * 5 ? B : sm_count * 5;
            cudaFuncSetAttribute(augment_simclr_fwd_cols_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            augment_simclr_fwd_cols_kernel<32, 1><<<g5, kMaxThreads, smem, st>>>(x, y, params, B, order);
        } else {
            cudaFuncSetAttribute(augment_simclr_fwd_cols_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            augment_simclr_fwd_cols_kernel<64, 1><<<grid, kMaxThreads, smem, st>>>(x, y, params, B, order);
        }
    } else if (H == 32 && W == 32) LAUNCH_FWD(1, 32);
    else if (ls.qpt == 1) LAUNCH_FWD(1, 0); else if (ls.qpt == 2) LAUNCH_FWD(2, 0); else LAUNCH_FWD(4, 0);
#undef LAUNCH_FWD
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("augment_simclr_fwd");
    return CB200_OK;
}

extern "C" int cb200_augment_simclr_bwd(const float* x, const float* dy, float* dx, const float* params, int B,
                                        int H, int W, int order, void* stream) {
    CB200_CHECK_ARG(B >= 0 && H > 0 && W > 0 && (order >= -1 && order <= 1), "augment_bwd: bad shape/order");
    CB200_CHECK_ARG(W % 4 == 0, "augment_bwd: W must be a multiple of 4 (got %d)", W);
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) |
                      reinterpret_cast<uintptr_t>(dx)) & 15) == 0, "augment_bwd: pointers must be 16-byte aligned");
    if (B == 0) return CB200_OK;
    LaunchShape ls;
    CB200_CHECK_ARG(pick_shape(H, W, &ls),
                    "augment_bwd: images larger than 64x64 (%dx%d) need the tiled path (not built yet)", H, W);
    size_t smem = (size_t)(6 * H * W + 96 + 4 * W + 4 * H) * sizeof(float) + 2 * sizeof(uint64_t);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LAUNCH_BWD(Q)                                                                                          \
    do {                                                                                                       \
        cudaFuncSetAttribute(augment_simclr_bwd_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        augment_simclr_bwd_kernel<Q><<<B, ls.threads, smem, st>>>(x, dy, dx, params, B, H, W, order);          \
    } while (0)
    if (ls.qpt == 1) LAUNCH_BWD(1); else if (ls.qpt == 2) LAUNCH_BWD(2); else LAUNCH_BWD(4);
#undef LAUNCH_BWD
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("augment_simclr_bwd");
    return CB200_OK;
}

// Any image size (used for H*W > 4096).  `means` [B,3] is written by the forward pass and must be handed to the
// backward pass unchanged; `gsums` [B,3] is backward scratch.  Both are caller-allocated.
extern "C" int cb200_augment_simclr_large_fwd(const float* x, float* y, const float* params, float* means, int B, int H,
                                              int W, int order, void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && H > 0 && W > 0 && (order >= -1 && order <= 1), "augment_large_fwd: bad shape/order");
    if (B == 0) return CB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(means, 0, sizeof(float) * 3 * (size_t)B, st);
    if (e != cudaSuccess) { cb200_set_error("augment_large_fwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const dim3 grid = large_grid(B, H, W);
    augment_large_mean_kernel<<<grid, kMaxThreads, 0, st>>>(x, params, means, B, H, W, order);
    CB200_COUNT_LAUNCH();
    augment_large_apply_kernel<<<grid, kMaxThreads, 0, st>>>(x, y, params, means, B, H, W, order);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("augment_simclr_large_fwd");
    return CB200_OK;
}

extern "C" int cb200_augment_simclr_large_bwd(const float* x, const float* dy, float* dx, const float* params,
                                              const float* means, float* gsums, int B, int H, int W, int order,
                                              void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && H > 0 && W > 0 && (order >= -1 && order <= 1), "augment_large_bwd: bad shape/order");
    if (B == 0) return CB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(gsums, 0, sizeof(float) * 3 * (size_t)B, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dx, 0, sizeof(float) * 3 * (size_t)B * H * W, st);
    if (e != cudaSuccess) { cb200_set_error("augment_large_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const dim3 grid = large_grid(B, H, W);
    augment_large_bwd_reduce_kernel<<<grid, kMaxThreads, 0, st>>>(x, dy, params, means, gsums, B, H, W, order);
    CB200_COUNT_LAUNCH();
    augment_large_bwd_scatter_kernel<<<grid, kMaxThreads, 0, st>>>(x, dy, dx, params, means, gsums, B, H, W, order);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("augment_simclr_large_bwd");
    return CB200_OK;
}


// Row f3: forward of the chain over B = n_u8_views + n_f32 views, where view b < n_u8_views reads the uint8 image
// (b mod n_u8) of `x_u8` [n_u8,3,H,W] (value / 255, ToTensor) and the others the fp32 images of `x_f32`
// [B - n_u8_views,3,H,W] in order.  With n_u8_views = 2 * n_u8 this is augment(cat[x, x, G(z)]) of
// training/gan/contrad.py:38-41 without the host-side conversion, the 4 B/element upload and the concatenation.
// `means` [B,3] is caller-allocated scratch (written only on the any-size path; hand it to
// cb200_augment_simclr_large_bwd, sliced to the fp32 views, for the gradient of those).  Either source may be empty
// (n_u8_views == 0 or == B); its pointer is then ignored.
extern "C" int cb200_augment_simclr_mixed_fwd(const unsigned char* x_u8, int n_u8, int n_u8_views, const float* x_f32,
                                              float* y, const float* params, float* means, int B, int H, int W,
                                              int order, void* stream) {
    CB200_CHECK_ARG(B >= 0 && B <= 65535 && H > 0 && W > 0 && (order >= -1 && order <= 1), "augment_mixed_fwd: bad shape/order");
    CB200_CHECK_ARG(n_u8_views >= 0 && n_u8_views <= B && (n_u8_views == 0 || n_u8 > 0),
                    "augment_mixed_fwd: %d uint8 views of %d images in a batch of %d", n_u8_views, n_u8, B);
    if (B == 0) return CB200_OK;
    if (n_u8 < 1) n_u8 = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(x_u8) | reinterpret_cast<uintptr_t>(x_f32)) & 15) == 0;
    if (H == W && (H == 32 || H == 64) && aligned) {
        const size_t smem = (size_t)(9 * H * W + 256 + 96 + 8 * H) * sizeof(float) + 2 * sizeof(uint64_t);
        int dev = 0, sm_count = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
        int per_sm = (int)((200 * 1024) / smem);
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        const int grid = B < sm_count * per_sm ? B : sm_count * per_sm;
        if (H == 32) {
            cudaFuncSetAttribute(augment_simclr_fwd_mixed_cols_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            augment_simclr_fwd_mixed_cols_kernel<32><<<grid, kMaxThreads, smem, st>>>(x_u8, n_u8, n_u8_views, x_f32, y, params, B, order);
        } else {
            cudaFuncSetAttribute(augment_simclr_fwd_mixed_cols_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            augment_simclr_fwd_mixed_cols_kernel<64><<<grid, kMaxThreads, smem, st>>>(x_u8, n_u8, n_u8_views, x_f32, y, params, B, order);
        }
        CB200_COUNT_LAUNCH();
        CB200_CHECK_LAUNCH("augment_simclr_mixed_fwd");
        return CB200_OK;
    }
    cudaError_t e = cudaMemsetAsync(means, 0, sizeof(float) * 3 * (size_t)B, st);
    if (e != cudaSuccess) { cb200_set_error("augment_mixed_fwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const dim3 grid = large_grid(B, H, W);
    augment_large_mean_mixed_kernel<<<grid, kMaxThreads, 0, st>>>(x_u8, n_u8, n_u8_views, x_f32, params, means, B, H, W, order);
    CB200_COUNT_LAUNCH();
    augment_large_apply_mixed_kernel<<<grid, kMaxThreads, 0, st>>>(x_u8, n_u8, n_u8_views, x_f32, y, params, means, B, H, W, order);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("augment_simclr_mixed_fwd");
    return CB200_OK;
}
