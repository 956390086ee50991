// First discriminator layer of SNDCGAN: Conv2d(3 -> 64, 3x3, stride 1, pad 1) + bias + LeakyReLU,
// with the `x*2-1` input affine folded in (reference: models/gan/sndcgan.py:91-93,122-124).
//
// K = 27 makes this layer HBM/LSU-bound (AI ~ 13 FLOP/B), so it is a register-tiled
// SIMT direct convolution rather than a tensor-core GEMM:
//   forward : NCHW [B,3,H,W] image in -> NHWC [B,H,W,64] activation out (the layout the tcgen05
//             kernels of the next layers consume), TF32-rounded because it feeds tcgen05.mma.
//             CTA = 4 output rows x 32 columns of one image; thread = 8 pixels x 4 channels.
//             Algorithmic bytes: 12 B/pixel in + 256 B/pixel out.
//   wgrad   : dW[64,3,3,3] and db[64] from dY (NHWC, already multiplied by lrelu') and the image.
//   dgrad   : (G step only) runs on the tensor cores through cb200_conv2d_nhwc_dgrad with the input
//             channels padded 3 -> 32; `conv_first_dgrad_finish` extracts the 3 real channels,
//             applies the factor 2 of the input affine and transposes NHWC -> NCHW.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kCo = 64;
constexpr int kThreads = 256;
constexpr int kRows = 4;        // output rows per CTA
constexpr int kCols = 32;       // output columns per CTA

// wmat: [64][27] with column ci*9 + kh*3 + kw (the OIHW weight / sigma), bias [64]
//
// PERSISTENT CTAs (round 2): the grid is a few CTAs per SM, each walking a strided list of (image, 4-row x 32-column)
// tiles.  The 27 x 64 weight block is transposed into shared memory ONCE per CTA (round 1 re-read it with a 108-byte
// stride in every one of the 12 288 CTAs of a B = 1536 launch), the halo tile of the NEXT iteration is fetched into
// registers before the FMA loop and parked in the other half of a shared-memory double buffer after it, so the global
// loads of tile i+1 overlap the arithmetic and the stores of tile i, with one CTA barrier per tile.
constexpr int kFwdXsElems = 3 * (kRows + 2) * (kCols + 2);                  // 612 image values (+halo) per tile
constexpr int kFwdXsPerThread = (kFwdXsElems + kThreads - 1) / kThreads;     // 3

__global__ void __launch_bounds__(kThreads)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ sigma,
                      const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W, float slope,
                      int round_out, float in_scale, float in_shift) {
    // row stride 36 floats (144 B): the 10 input values of a thread start at a 16-byte boundary (seg = 0, 8, 16, 24), so
    // they are two LDS.128 + one LDS.64 instead of ten LDS.32 - the shared-memory pipe, not the FMA pipe, was the
    // busier one (22 wavefronts per 96 FMAs per warp and (ci, kh) step; now 15)
    __shared__ __align__(16) float xs[2][3][kRows + 2][kCols + 4];
    __shared__ __align__(16) float ws[27][kCo];
    const int tiles_w = (W + kCols - 1) / kCols, tiles_h = (H + kRows - 1) / kRows;
    const int ntiles = tiles_w * tiles_h * B;
    const float inv_sigma = sigma ? sigma[1] : 1.f;
    for (int i = threadIdx.x; i < 27 * kCo; i += kThreads) {        // lanes along co: conflict-free shared stores
        int t = i / kCo, co = i % kCo;
        ws[t][co] = __ldg(w + co * 27 + t) * inv_sigma;
    }
    // this thread's slots of the halo tile (tile independent)
    int sc[kFwdXsPerThread], sr[kFwdXsPerThread], scc[kFwdXsPerThread];
#pragma unroll
    for (int k = 0; k < kFwdXsPerThread; ++k) {
        const int i = threadIdx.x + k * kThreads;
        sc[k] = i / ((kRows + 2) * (kCols + 2));
        sr[k] = (i / (kCols + 2)) % (kRows + 2);
        scc[k] = i % (kCols + 2);
    }
    auto origin = [&](int tile, int& b, int& h0, int& w0) {
        w0 = (tile % tiles_w) * kCols;
        const int rest = tile / tiles_w;
        h0 = (rest % tiles_h) * kRows;
        b = rest / tiles_h;
    };
    auto fetch = [&](int tile, float (&v)[kFwdXsPerThread]) {
        int b, h0, w0;
        origin(tile, b, h0, w0);
#pragma unroll
        for (int k = 0; k < kFwdXsPerThread; ++k) {
            v[k] = 0.f;
            if (threadIdx.x + k * kThreads < kFwdXsElems) {
                const int hh = h0 + sr[k] - 1, ww = w0 + scc[k] - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    v[k] = __ldg(x + ((size_t)(b * 3 + sc[k]) * H + hh) * W + ww) * in_scale + in_shift;
            }
        }
    };
    auto park = [&](int buf, const float (&v)[kFwdXsPerThread]) {
#pragma unroll
        for (int k = 0; k < kFwdXsPerThread; ++k)
            if (threadIdx.x + k * kThreads < kFwdXsElems) xs[buf][sc[k]][sr[k]][scc[k]] = v[k];
    };
    const int cg = threadIdx.x & 15;          // channels 4cg .. 4cg+3
    const int pg = threadIdx.x >> 4;          // 16 pixel groups: row = pg/4, 8-column segment = pg%4
    const int row = pg >> 2, seg = (pg & 3) * 8;
    const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cg) : make_float4(0.f, 0.f, 0.f, 0.f);

    int tile = blockIdx.x;
    if (tile >= ntiles) return;
    {
        float v[kFwdXsPerThread];
        fetch(tile, v);
        park(0, v);
    }
    __syncthreads();
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int next = tile + gridDim.x;
        float nv[kFwdXsPerThread];
        if (next < ntiles) fetch(next, nv);                        // in flight during the FMA loop below
        float acc[8][4];
#pragma unroll
        for (int p = 0; p < 8; ++p) { acc[p][0] = bv.x; acc[p][1] = bv.y; acc[p][2] = bv.z; acc[p][3] = bv.w; }
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* src = &xs[buf][ci][row + kh][seg];
                const float4 i0 = *reinterpret_cast<const float4*>(src), i1 = *reinterpret_cast<const float4*>(src + 4);
                const float2 i2 = *reinterpret_cast<const float2*>(src + 8);
                const float in[10] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y};
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 wv = *reinterpret_cast<const float4*>(&ws[ci * 9 + kh * 3 + kw][cg * 4]);
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        acc[p][0] += in[p + kw] * wv.x; acc[p][1] += in[p + kw] * wv.y;
                        acc[p][2] += in[p + kw] * wv.z; acc[p][3] += in[p + kw] * wv.w;
                    }
                }
            }
        }
        int b, h0, w0;
        origin(tile, b, h0, w0);
        const int hh = h0 + row;
        if (hh < H) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int ww = w0 + seg + p;
                if (ww >= W) continue;
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float v = acc[p][e];
                    v = v > 0.f ? v : v * slope;
                    o[e] = round_out ? round_tf32(v) : v;
                }
                *reinterpret_cast<float4*>(y + (((size_t)b * H + hh) * W + ww) * kCo + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        if (next < ntiles) park(buf ^ 1, nv);
        __syncthreads();      // next tile visible; everybody is done reading xs[buf] (refilled two iterations from now)
    }
}

// dW_hat[co][ci*9+kh*3+kw] += sum_pixels dY[pix][co] * (2x-1)[ci][h+kh-1][w+kw-1];   db[co] += sum dY
// (async staging helpers: common.cuh)
constexpr int kXsElems = 3 * (kRows + 2) * (kCols + 2);            // 612 image values (+halo) per tile
constexpr int kXsPerThread = (kXsElems + kThreads - 1) / kThreads;  // 3
constexpr int kDyTileFloats = kRows * kCols * kCo;                  // 8192 floats = 32 KB
constexpr size_t kWgradSmemBytes = (size_t)(2 * kDyTileFloats + 2 * kXsElems) * sizeof(float) + 2 * sizeof(uint64_t);

// Persistent CTAs: each CTA walks a strided list of (image, 4-row x 32-col) tiles and keeps its partial dW in
// REGISTERS across tiles, so the global atomics are issued once per CTA.  Threads: 16 channel groups (4 channels)
// x 4 tap groups (7 taps; 27 = 7+7+7+6) x 4 pixel slices (one output row each).
// The dY tile of the NEXT iteration (4 rows x 32 px x 64 ch = 32 KB, contiguous per row) is fetched by
// cp.async.bulk into the other half of a shared-memory ring while the current one is consumed, and the image halo
// tile of the next iteration is loaded into registers before the FMA loop and parked in shared memory after it:
// the first version read dY straight from global memory and sat in long-scoreboard stalls 7.8 of every 8 issue
// slots (profiles/prof_r1_conv_first_wgrad.md).
__global__ void __launch_bounds__(kThreads, 3)
conv_first_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                        float* __restrict__ db, int B, int H, int W, float in_scale, float in_shift) {
    extern __shared__ __align__(128) float wsm[];
    float* dys = wsm;                                   // [2][kRows][kCols][kCo]
    float* xsb = wsm + 2 * kDyTileFloats;               // [2][3][kRows+2][kCols+2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsb + 2 * kXsElems);
    const int tiles_w = (W + kCols - 1) / kCols, tiles_h = (H + kRows - 1) / kRows;
    const int ntiles = tiles_w * tiles_h * B;
    const int cg = threadIdx.x & 15;
    const int tg = (threadIdx.x >> 4) & 3;
    const int ps = threadIdx.x >> 6;          // pixel slice = output row within the tile

    // this thread's slots of the image halo tile (tile independent)
    int xs_c[kXsPerThread], xs_r[kXsPerThread], xs_cc[kXsPerThread];
#pragma unroll
    for (int k = 0; k < kXsPerThread; ++k) {
        const int i = threadIdx.x + k * kThreads;
        xs_c[k] = i / ((kRows + 2) * (kCols + 2));
        xs_r[k] = (i / (kCols + 2)) % (kRows + 2);
        xs_cc[k] = i % (kCols + 2);
    }
    // this thread's taps: offset of tap t inside the halo tile
    int tap_off[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
        const int tap = min(tg * 7 + t, 26);
        tap_off[t] = ((tap / 9) * (kRows + 2) + (tap % 9) / 3) * (kCols + 2) + tap % 3;
    }
    auto tile_origin = [&](int tile, int& b, int& h0, int& w0) {
        b = tile / (tiles_w * tiles_h);
        const int rem = tile - b * tiles_w * tiles_h;
        h0 = (rem / tiles_w) * kRows;
        w0 = (rem % tiles_w) * kCols;
    };
    auto load_x = [&](int tile, float (&v)[kXsPerThread]) {
        int b, h0, w0;
        tile_origin(tile, b, h0, w0);
#pragma unroll
        for (int k = 0; k < kXsPerThread; ++k) {
            const int hh = h0 + xs_r[k] - 1, ww = w0 + xs_cc[k] - 1;
            v[k] = 0.f;
            if (threadIdx.x + k * kThreads < kXsElems && hh >= 0 && hh < H && ww >= 0 && ww < W)
                v[k] = __ldg(x + ((size_t)(b * 3 + xs_c[k]) * H + hh) * W + ww) * in_scale + in_shift;
        }
    };
    auto park_x = [&](int buf, const float (&v)[kXsPerThread]) {
#pragma unroll
        for (int k = 0; k < kXsPerThread; ++k)
            if (threadIdx.x + k * kThreads < kXsElems) xsb[buf * kXsElems + threadIdx.x + k * kThreads] = v[k];
    };
    auto fetch_dy = [&](int tile, int buf) {              // one elected thread
        int b, h0, w0;
        tile_origin(tile, b, h0, w0);
        const int rows = min(kRows, H - h0), ncols = min(kCols, W - w0);
        const uint32_t row_bytes = (uint32_t)ncols * kCo * sizeof(float);
        bar_expect_tx(&bars[buf], row_bytes * rows);
        for (int r = 0; r < rows; ++r)
            bulk_load(dys + buf * kDyTileFloats + r * kCols * kCo, dy + (((size_t)b * H + h0 + r) * W + w0) * kCo, row_bytes,
                      &bars[buf]);
    };

    float acc[7][4];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 7; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }

    if (threadIdx.x == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    int tile = blockIdx.x;
    if (tile < ntiles) {
        if (threadIdx.x == 0) fetch_dy(tile, 0);
        float v[kXsPerThread];
        load_x(tile, v);
        park_x(0, v);
    }
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int next = tile + gridDim.x;
        float vnext[kXsPerThread];
        if (next < ntiles) {
            if (threadIdx.x == 0) fetch_dy(next, buf ^ 1);
            load_x(next, vnext);                      // in flight during the FMA loop below
        }
        bar_wait(&bars[buf], (uint32_t)(it >> 1) & 1u);
        __syncthreads();                               // halo tile `buf` parked by everybody
        int b, h0, w0;
        tile_origin(tile, b, h0, w0);
        if (h0 + ps < H) {
            const float* grow = dys + buf * kDyTileFloats + (ps * kCols) * kCo + cg * 4;
            const float* xrow = xsb + buf * kXsElems + ps * (kCols + 2);
            const int ncols = min(kCols, W - w0);
#pragma unroll 4
            for (int c = 0; c < ncols; ++c) {
                const float4 g = *reinterpret_cast<const float4*>(grow + c * kCo);
                if (tg == 0) { bacc[0] += g.x; bacc[1] += g.y; bacc[2] += g.z; bacc[3] += g.w; }
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    const float xv = xrow[tap_off[t] + c];
                    acc[t][0] += g.x * xv; acc[t][1] += g.y * xv; acc[t][2] += g.z * xv; acc[t][3] += g.w * xv;
                }
            }
        }
        if (next < ntiles) park_x(buf ^ 1, vnext);
        __syncthreads();                               // buffer `buf` is free for the prefetch of iteration it+1
    }
    // reduce the four pixel slices through shared memory (the dY ring is dead now), then one atomic per element
    float (*part)[28][kCo] = reinterpret_cast<float (*)[28][kCo]>(dys);
#pragma unroll
    for (int t = 0; t < 7; ++t) {
        const int tap = tg * 7 + t;
        if (tap < 27) *reinterpret_cast<float4*>(&part[ps][tap][cg * 4]) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    }
    if (tg == 0) *reinterpret_cast<float4*>(&part[ps][27][cg * 4]) = make_float4(bacc[0], bacc[1], bacc[2], bacc[3]);
    __syncthreads();
    for (int i = threadIdx.x; i < 28 * kCo; i += kThreads) {
        const int tap = i / kCo, co = i % kCo;
        const float s = (part[0][tap][co] + part[1][tap][co]) + (part[2][tap][co] + part[3][tap][co]);
        if (tap < 27) atomicAdd(dw + co * 27 + tap, s);
        else if (db) atomicAdd(db + co, s);
    }
}

// ---- weight gradient, second mapping (CB200_CONV_FIRST_WGRAD=2; written without GPU access, to be timed) ----
// The kernel above spends, per pixel and thread, one LDS.128 (dY) + seven LDS.32 (image) on 28 FMAs: ~11 shared-memory
// wavefronts per warp against 28 FMA issue slots, and with 8 warps per CTA sharing one shared-memory pipe that pipe
// (88 cycles per 4-pixel step) - not the FMA pipes (56) - bounds it.  Here a thread owns ALL nine taps of one input
// channel (3 tap groups = the 3 input channels, 192 threads): the three image rows it needs slide along the row in
// registers, so a pixel costs one LDS.128 + three LDS.32 for 36 FMAs (7 wavefronts; 42 vs 54 cycles per step).
constexpr int kThreadsV2 = 192;
constexpr int kXsPerThreadV2 = (kXsElems + kThreadsV2 - 1) / kThreadsV2;      // 4

__global__ void __launch_bounds__(kThreadsV2, 3)
conv_first_wgrad_v2_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                           float* __restrict__ db, int B, int H, int W, float in_scale, float in_shift) {
    extern __shared__ __align__(128) float wsm[];
    float* dys = wsm;                                   // [2][kRows][kCols][kCo]
    float* xsb = wsm + 2 * kDyTileFloats;               // [2][3][kRows+2][kCols+2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsb + 2 * kXsElems);
    const int tiles_w = (W + kCols - 1) / kCols, tiles_h = (H + kRows - 1) / kRows;
    const int ntiles = tiles_w * tiles_h * B;
    const int cg = threadIdx.x & 15;          // channels 4 cg .. 4 cg + 3
    const int ci = (threadIdx.x >> 4) % 3;    // input channel = tap group (taps ci*9 .. ci*9 + 8)
    const int ps = threadIdx.x / 48;          // output row within the tile

    auto tile_origin = [&](int tile, int& b, int& h0, int& w0) {
        b = tile / (tiles_w * tiles_h);
        const int rem = tile - b * tiles_w * tiles_h;
        h0 = (rem / tiles_w) * kRows;
        w0 = (rem % tiles_w) * kCols;
    };
    auto load_x = [&](int tile, float (&v)[kXsPerThreadV2]) {
        int b, h0, w0;
        tile_origin(tile, b, h0, w0);
#pragma unroll
        for (int k = 0; k < kXsPerThreadV2; ++k) {
            const int i = threadIdx.x + k * kThreadsV2;
            const int c = i / ((kRows + 2) * (kCols + 2)), r = (i / (kCols + 2)) % (kRows + 2), cc = i % (kCols + 2);
            const int hh = h0 + r - 1, ww = w0 + cc - 1;
            v[k] = 0.f;
            if (i < kXsElems && hh >= 0 && hh < H && ww >= 0 && ww < W)
                v[k] = __ldg(x + ((size_t)(b * 3 + c) * H + hh) * W + ww) * in_scale + in_shift;
        }
    };
    auto park_x = [&](int buf, const float (&v)[kXsPerThreadV2]) {
#pragma unroll
        for (int k = 0; k < kXsPerThreadV2; ++k)
            if (threadIdx.x + k * kThreadsV2 < kXsElems) xsb[buf * kXsElems + threadIdx.x + k * kThreadsV2] = v[k];
    };
    auto fetch_dy = [&](int tile, int buf) {              // one elected thread
        int b, h0, w0;
        tile_origin(tile, b, h0, w0);
        const int rows = min(kRows, H - h0), ncols = min(kCols, W - w0);
        const uint32_t row_bytes = (uint32_t)ncols * kCo * sizeof(float);
        bar_expect_tx(&bars[buf], row_bytes * rows);
        for (int r = 0; r < rows; ++r)
            bulk_load(dys + buf * kDyTileFloats + r * kCols * kCo, dy + (((size_t)b * H + h0 + r) * W + w0) * kCo, row_bytes,
                      &bars[buf]);
    };

    float acc[9][4];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }

    if (threadIdx.x == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    int tile = blockIdx.x;
    if (tile < ntiles) {
        if (threadIdx.x == 0) fetch_dy(tile, 0);
        float v[kXsPerThreadV2];
        load_x(tile, v);
        park_x(0, v);
    }
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int next = tile + gridDim.x;
        float vnext[kXsPerThreadV2];
        if (next < ntiles) {
            if (threadIdx.x == 0) fetch_dy(next, buf ^ 1);
            load_x(next, vnext);                      // in flight during the FMA loop below
        }
        bar_wait(&bars[buf], (uint32_t)(it >> 1) & 1u);
        __syncthreads();                               // halo tile `buf` parked by everybody
        int b, h0, w0;
        tile_origin(tile, b, h0, w0);
        if (h0 + ps < H) {
            const float* grow = dys + buf * kDyTileFloats + (ps * kCols) * kCo + cg * 4;
            const float* xr0 = xsb + buf * kXsElems + (ci * (kRows + 2) + ps) * (kCols + 2);      // image row ps-1+kh, kh = 0
            const float* xr1 = xr0 + (kCols + 2), *xr2 = xr1 + (kCols + 2);
            const int ncols = min(kCols, W - w0);
            float a0 = xr0[0], b0 = xr0[1], a1 = xr1[0], b1 = xr1[1], a2 = xr2[0], b2 = xr2[1];   // columns c, c+1 of the window
#pragma unroll 4
            for (int c = 0; c < ncols; ++c) {
                const float4 g = *reinterpret_cast<const float4*>(grow + c * kCo);
                const float c0 = xr0[c + 2], c1 = xr1[c + 2], c2 = xr2[c + 2];                     // column c+2 enters the window
                if (ci == 0) { bacc[0] += g.x; bacc[1] += g.y; bacc[2] += g.z; bacc[3] += g.w; }
                const float xv[9] = {a0, b0, c0, a1, b1, c1, a2, b2, c2};                          // tap = kh * 3 + kw
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    acc[t][0] += g.x * xv[t]; acc[t][1] += g.y * xv[t]; acc[t][2] += g.z * xv[t]; acc[t][3] += g.w * xv[t];
                }
                a0 = b0; b0 = c0; a1 = b1; b1 = c1; a2 = b2; b2 = c2;
            }
        }
        if (next < ntiles) park_x(buf ^ 1, vnext);
        __syncthreads();                               // buffer `buf` is free for the prefetch of iteration it+1
    }
    // reduce the four pixel slices through shared memory (the dY ring is dead now), then one atomic per element
    float (*part)[28][kCo] = reinterpret_cast<float (*)[28][kCo]>(dys);
#pragma unroll
    for (int t = 0; t < 9; ++t)
        *reinterpret_cast<float4*>(&part[ps][ci * 9 + t][cg * 4]) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    if (ci == 0) *reinterpret_cast<float4*>(&part[ps][27][cg * 4]) = make_float4(bacc[0], bacc[1], bacc[2], bacc[3]);
    __syncthreads();
    for (int i = threadIdx.x; i < 28 * kCo; i += kThreadsV2) {
        const int tap = i / kCo, co = i % kCo;
        const float sum = (part[0][tap][co] + part[1][tap][co]) + (part[2][tap][co] + part[3][tap][co]);
        if (tap < 27) atomicAdd(dw + co * 27 + tap, sum);
        else if (db) atomicAdd(db + co, sum);
    }
}

int wgrad_variant() {
    // default = the second mapping (nine taps of one input channel per thread, sliding register window): measured 3-8 %
    // faster than the first at B = 192 / 512 / 1536 (tools/bench_conv_first.py, round 2); CB200_CONV_FIRST_WGRAD=1 selects
    // the first one for A/B runs
    static const int v = []() { const char* e = getenv("CB200_CONV_FIRST_WGRAD"); return (e && e[0] == '1') ? 1 : 2; }();
    return v;
}

// dx[b,c,h,w] = 2 * dpad[b,h,w,c]   (c < 3 of the 32 padded channels)
__global__ void __launch_bounds__(kThreads)
conv_first_dgrad_finish_kernel(const float* __restrict__ dpad, float* __restrict__ dx, int HW, int cpad,
                               long long total) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;   // over B*3*HW (NCHW order)
    if (i >= total) return;
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % 3);
    const long long b = i / (3LL * HW);
    dx[i] = 2.f * __ldg(dpad + (b * HW + p) * cpad + c);
}

}  // namespace

// y[B,H,W,64] = lrelu_slope(conv3x3(in_scale*x+in_shift, w/sigma) + bias); x NCHW [B,3,H,W]; w = OIHW [64,3,3,3];
// sigma [2] or NULL.  D's first layer uses in_scale=2, in_shift=-1 (the folded `x*2-1`).
extern "C" int cb200_conv_first_fwd(const float* x, const float* w, const float* sigma, const float* bias, float* y,
                                    int B, int H, int W, float slope, int round_out, float in_scale, float in_shift,
                                    void* stream) {
    CB200_CHECK_ARG(B > 0 && H > 0 && W > 0, "conv_first_fwd: empty input");
    CB200_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv_first_fwd: y must be 16-byte aligned");
    const long long ntiles = (long long)((W + kCols - 1) / kCols) * ((H + kRows - 1) / kRows) * B;
    CB200_CHECK_ARG(ntiles < (1LL << 31), "conv_first_fwd: too many tiles");
    static const int per_sm = []() {          // resident CTAs per SM (80 registers x 256 threads -> 3); env override for A/B runs
        const char* e = getenv("CB200_CONV_FIRST_FWD_CTAS");
        if (e) return atoi(e);
        int n = 3;
#if defined(__CUDACC__)
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv_first_fwd_kernel, kThreads, 0) != cudaSuccess || n < 1) n = 3;
#endif
        return n;
    }();
    const int grid = (int)(ntiles < (long long)per_sm * 148 ? ntiles : (long long)per_sm * 148);
    conv_first_fwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, w, sigma, bias, y, B, H, W, slope,
                                                                                    round_out, in_scale, in_shift);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("conv_first_fwd");
    return CB200_OK;
}

// dw_hat[64,27] (OIHW order, gradient w.r.t. w/sigma) and db[64] are ACCUMULATED into (caller zeroes them).
extern "C" int cb200_conv_first_wgrad(const float* x, const float* dy, float* dw_hat, float* db, int B, int H, int W,
                                      float in_scale, float in_shift, void* stream) {
    CB200_CHECK_ARG(B > 0 && H > 0 && W > 0, "conv_first_wgrad: empty input");
    const long long ntiles = (long long)((W + kCols - 1) / kCols) * ((H + kRows - 1) / kRows) * B;
    CB200_CHECK_ARG((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "conv_first_wgrad: dy must be 16-byte aligned");
    const int grid = (int)(ntiles < 3 * 148 ? ntiles : 3 * 148);           // 3 resident CTAs per SM (69 KB smem each)
    if (wgrad_variant() == 2) {
        cudaFuncSetAttribute(conv_first_wgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmemBytes);
        conv_first_wgrad_v2_kernel<<<grid, kThreadsV2, kWgradSmemBytes, static_cast<cudaStream_t>(stream)>>>(
            x, dy, dw_hat, db, B, H, W, in_scale, in_shift);
    } else {
        cudaFuncSetAttribute(conv_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmemBytes);
        conv_first_wgrad_kernel<<<grid, kThreads, kWgradSmemBytes, static_cast<cudaStream_t>(stream)>>>(
            x, dy, dw_hat, db, B, H, W, in_scale, in_shift);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("conv_first_wgrad");
    return CB200_OK;
}

// dx[B,3,H,W] = 2 * dpad[B,H,W,cpad][..., :3]
extern "C" int cb200_conv_first_dgrad_finish(const float* dpad, float* dx, int B, int H, int W, int cpad, void* stream) {
    CB200_CHECK_ARG(B > 0 && cpad >= 3, "conv_first_dgrad_finish: bad shape");
    const long long total = (long long)B * 3 * H * W;
    conv_first_dgrad_finish_kernel<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0,
                                     static_cast<cudaStream_t>(stream)>>>(dpad, dx, H * W, cpad, total);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("conv_first_dgrad_finish");
    return CB200_OK;
}
