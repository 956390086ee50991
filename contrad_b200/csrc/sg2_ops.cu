// StyleGAN2-side kernels of the ContraD hot path for sm_100a: everything around the
// tensor-core GEMMs of ResidualDiscriminatorP / Generator is HBM-bound SIMT work on NHWC activations.
//
//   upfirdn2d           models/gan/stylegan2/op/upfirdn2d.py:159-200 (and upfirdn2d_kernel.cu): zero-stuffing by `up`,
//                       padding / cropping, FIR, decimation by `down`, on tensors with explicit strides (NCHW or NHWC)
//   patch_s2_*          the 3x3 stride-2 patch gather (im2col) behind `EqualConv2d(..., stride=2, padding=0)` after
//                       Blur (layers.py:174-198) and its transpose = the scatter of `conv_transpose2d(stride=2)`
//                       (generator.py:66-75); both convolutions then are plain tensor-core GEMMs
//   bias_act(_grad)     FusedLeakyReLU / fused_leaky_relu (op/fused_act.py:74-94) with an optional fused residual
//   modulate, mul_reduce, mod_epilogue, noise_grad
//                       ModulatedConv2d in its "scale activations, shared weights, scale outputs" form
//                       (generator.py:52-82 is the equivalent per-sample-weight grouped conv), NoiseInjection
//                       (generator.py:85-94) + FusedLeakyReLU fused in one epilogue
//   stddev_*            _minibatch_stddev_layer (discriminator.py:22-33) forward, backward and backward-of-backward
//                       (the R1 penalty differentiates the input gradient once more, train_stylegan2.py:106-113)
//   rgb_to_nhwc / nhwc_to_rgb, pixelnorm, row_sqsum / row_scale (R1's per-sample |g|^2), axpby, ema_lerp
//                       (utils.py:130-143 `accumulate`)
#include "common.cuh"
#include "contrad_b200.h"

namespace {

constexpr int kT = 256;

inline int grid_for(long long work, int per_block = kT, int cap = 148 * 16) {
    long long g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------ upfirdn2d
struct UpfirdnParams {
    const float* x; float* y; const float* k;
    long long xs_n, xs_h, xs_w, xs_c, ys_n, ys_h, ys_w, ys_c;
    int N, C, Hi, Wi, Ho, Wo, up, down, px0, py0, kh, kw, flip, c_fast, round_out;
    float gain;
};

__global__ void __launch_bounds__(kT) upfirdn2d_kernel(const __grid_constant__ UpfirdnParams p) {
    __shared__ float ks[64];
    for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
        // the reference correlates the padded signal with the FLIPPED kernel (upfirdn2d.py:185); `flip` undoes it
        // for the backward pass, which uses the flipped kernel again (upfirdn2d.py:113).
        const int ky = i / p.kw, kx = i % p.kw;
        const int src = p.flip ? i : (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx);
        ks[i] = __ldg(p.k + src) * p.gain;
    }
    __syncthreads();
    const long long total = (long long)p.N * p.C * p.Ho * p.Wo;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int n, c, oy, ox;
        long long t = idx;
        if (p.c_fast) { c = (int)(t % p.C); t /= p.C; ox = (int)(t % p.Wo); t /= p.Wo; oy = (int)(t % p.Ho); n = (int)(t / p.Ho); }
        else { ox = (int)(t % p.Wo); t /= p.Wo; oy = (int)(t % p.Ho); t /= p.Ho; c = (int)(t % p.C); n = (int)(t / p.C); }
        const float* xb = p.x + n * p.xs_n + c * p.xs_c;
        float acc = 0.f;
        for (int ky = 0; ky < p.kh; ++ky) {
            const int py = oy * p.down + ky - p.py0;           // position on the zero-stuffed grid
            if (py < 0 || py % p.up != 0) continue;
            const int iy = py / p.up;
            if (iy >= p.Hi) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                const int px = ox * p.down + kx - p.px0;
                if (px < 0 || px % p.up != 0) continue;
                const int ix = px / p.up;
                if (ix >= p.Wi) continue;
                acc = fmaf(__ldg(xb + iy * p.xs_h + ix * p.xs_w), ks[ky * p.kw + kx], acc);
            }
        }
        if (p.round_out) acc = round_tf32(acc);
        p.y[n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w] = acc;
    }
}

// Contiguous NHWC, C % 4 == 0: one thread = one output pixel x 4 channels (float4, channel fastest -> coalesced); the
// taps that hit the zero-stuffed grid are enumerated with stride `up` (no modulo in the loops); neighbouring pixels
// re-read the same inputs from L1/L2, HBM sees each input and output once.
// A CTA owns one kTile x kTile tile of output pixels of one image (all channels), so that the (kTile*down + k - 1)^2
// input pixels it touches stay in this SM's L1 while neighbouring output pixels re-read them: HBM / L2 see each input
// about once instead of once per tap row.
// UP / KS != 0 bake the up-sampling factor and the (square) FIR size into the code: the divisions by `up` disappear and
// the 4x4 tap loops of the [1,3,3,1] kernels unroll (the runtime-parameter version spent most of its issue slots on
// integer division).
constexpr int kTile = 8;
template <int UP, int KS>
__global__ void __launch_bounds__(kT) upfirdn2d_nhwc4_kernel(const __grid_constant__ UpfirdnParams p_) {
    UpfirdnParams p = p_;
    if (UP) p.up = UP;
    if (KS) { p.kh = KS; p.kw = KS; }
    __shared__ float ks[64];
    for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
        const int ky = i / p.kw, kx = i % p.kw;
        const int src = p.flip ? i : (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx);
        ks[i] = __ldg(p.k + src) * p.gain;
    }
    __syncthreads();
    const int C4 = p.C >> 2;
    const int tiles_x = (p.Wo + kTile - 1) / kTile, tiles_y = (p.Ho + kTile - 1) / kTile;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(p.x);
    float4* __restrict__ y4 = reinterpret_cast<float4*>(p.y);
    const int per_tile = kTile * kTile * C4;
    const long long n_tiles = (long long)p.N * tiles_y * tiles_x;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y), n = (int)(tile / ((long long)tiles_x * tiles_y));
      for (int e = threadIdx.x; e < per_tile; e += blockDim.x) {
        const int c = e % C4, pix = e / C4;
        const int ox = tx * kTile + (pix % kTile), oy = ty * kTile + (pix / kTile);
        if (ox >= p.Wo || oy >= p.Ho) continue;
        const long long idx = (((long long)n * p.Ho + oy) * p.Wo + ox) * C4 + c;
        const int py_base = oy * p.down - p.py0, px_base = ox * p.down - p.px0;
        const int ky0 = UP == 1 ? 0 : ((-py_base) % p.up + p.up) % p.up, kx0 = UP == 1 ? 0 : ((-px_base) % p.up + p.up) % p.up;
        const int iy0 = (py_base + ky0) / p.up, ix0 = (px_base + kx0) / p.up;     // input index advances by 1 per tap step
        const float4* img = x4 + (long long)n * p.Hi * p.Wi * C4 + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int kSteps = KS ? (KS + (UP ? UP : 1) - 1) / (UP ? UP : 1) : 64;   // taps per axis that can hit the grid
        constexpr int kUnroll = KS ? kSteps : 1;
#pragma unroll kUnroll
        for (int sy = 0; sy < kSteps; ++sy) {
            const int ky = ky0 + sy * p.up, iy = iy0 + sy;
            if (ky >= p.kh) break;
            if (iy < 0 || iy >= p.Hi) continue;
            const float4* row = img + (long long)iy * p.Wi * C4;
#pragma unroll kUnroll
            for (int sx = 0; sx < kSteps; ++sx) {
                const int kx = kx0 + sx * p.up, ix = ix0 + sx;
                if (kx >= p.kw) break;
                if (ix < 0 || ix >= p.Wi) continue;
                const float4 v = __ldg(row + ix * C4);
                const float w = ks[ky * p.kw + kx];
                acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
            }
        }
        if (p.round_out) { acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w); }
        y4[idx] = acc;
      }
    }
}

// ------------------------------------------------------------------------------------------------ 3x3 stride-2 patches
// gather : u[b, oh, ow, kh*3+kw, c] = x[b, 2*oh+kh, 2*ow+kw, c]          (x is [B, 2*Ho+1, 2*Wo+1, C])
// scatter: x[b, r, q, c] = sum over (oh,kh),(ow,kw) with 2*oh+kh = r, 2*ow+kw = q of u[b, oh, ow, kh*3+kw, c]
__global__ void __launch_bounds__(kT) patch_s2_gather_kernel(const float* __restrict__ x, float* __restrict__ u, int B, int Ho,
                                                             int Wo, int C4, int round_out) {
    const int Hi = 2 * Ho + 1, Wi = 2 * Wo + 1;
    const long long total = (long long)B * Ho * Wo * 9 * C4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* u4 = reinterpret_cast<float4*>(u);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long t = idx;
        const int c = (int)(t % C4); t /= C4;
        const int k = (int)(t % 9); t /= 9;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int r = 2 * oh + k / 3, q = 2 * ow + k % 3;
        float4 v = __ldg(x4 + (((long long)b * Hi + r) * Wi + q) * C4 + c);
        if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        u4[idx] = v;
    }
}

__global__ void __launch_bounds__(kT) patch_s2_scatter_kernel(const float* __restrict__ u, float* __restrict__ x, int B, int Ho,
                                                              int Wo, int C4, int round_out) {
    const int Hi = 2 * Ho + 1, Wi = 2 * Wo + 1;
    const long long total = (long long)B * Hi * Wi * C4;
    const float4* u4 = reinterpret_cast<const float4*>(u);
    float4* x4 = reinterpret_cast<float4*>(x);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long t = idx;
        const int c = (int)(t % C4); t /= C4;
        const int q = (int)(t % Wi); t /= Wi;
        const int r = (int)(t % Hi);
        const int b = (int)(t / Hi);
        // the (o, k) pairs that touch coordinate r: even r -> (r/2, 0) and (r/2 - 1, 2); odd r -> ((r-1)/2, 1)
        int oh[2], kh[2], nh = 0, ow[2], kw[2], nw = 0;
        if (r & 1) { oh[0] = r >> 1; kh[0] = 1; nh = 1; }
        else {
            if ((r >> 1) < Ho) { oh[nh] = r >> 1; kh[nh] = 0; ++nh; }
            if ((r >> 1) >= 1) { oh[nh] = (r >> 1) - 1; kh[nh] = 2; ++nh; }
        }
        if (q & 1) { ow[0] = q >> 1; kw[0] = 1; nw = 1; }
        else {
            if ((q >> 1) < Wo) { ow[nw] = q >> 1; kw[nw] = 0; ++nw; }
            if ((q >> 1) >= 1) { ow[nw] = (q >> 1) - 1; kw[nw] = 2; ++nw; }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < nh; ++i)
            for (int j = 0; j < nw; ++j) {
                const float4 v = __ldg(u4 + ((((long long)b * Ho + oh[i]) * Wo + ow[j]) * 9 + kh[i] * 3 + kw[j]) * C4 + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        if (round_out) { acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w); }
        x4[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ bias + leaky relu
// mode 0: y = lrelu(x + b) * gain (+ res)        mode 1: y = x * ((ref + b) > 0 ? gain : gain * slope)
__global__ void __launch_bounds__(kT) bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                      const float* __restrict__ ref, const float* __restrict__ res,
                                                      float* __restrict__ y, long long n, int C, int mode, float slope,
                                                      float gain, int round_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float b = bias ? __ldg(bias + (int)(i % C)) : 0.f;
        float v;
        if (mode == 0) {
            const float t = x[i] + b;
            v = (t > 0.f ? t : t * slope) * gain;
            if (res) v += res[i];
        } else {
            v = x[i] * ((ref[i] + b) > 0.f ? gain : gain * slope);
        }
        y[i] = round_out ? round_tf32(v) : v;
    }
}

// float4 variant (C % 4 == 0, 16-byte aligned operands)
__global__ void __launch_bounds__(kT) bias_act_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ bias,
                                                          const float4* __restrict__ ref, const float4* __restrict__ res,
                                                          float4* __restrict__ y, long long n4, int C4, int mode, float slope,
                                                          float gain, int round_out) {
    const float gneg = gain * slope;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 b = bias ? __ldg(bias + (int)(i % C4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v = x[i];
        float4 o;
        if (mode == 0) {
            float t;
            t = v.x + b.x; o.x = (t > 0.f ? t : t * slope) * gain;
            t = v.y + b.y; o.y = (t > 0.f ? t : t * slope) * gain;
            t = v.z + b.z; o.z = (t > 0.f ? t : t * slope) * gain;
            t = v.w + b.w; o.w = (t > 0.f ? t : t * slope) * gain;
            if (res) { const float4 r = res[i]; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
        } else {
            const float4 r = ref[i];
            o.x = v.x * ((r.x + b.x) > 0.f ? gain : gneg);
            o.y = v.y * ((r.y + b.y) > 0.f ? gain : gneg);
            o.z = v.z * ((r.z + b.z) > 0.f ? gain : gneg);
            o.w = v.w * ((r.w + b.w) > 0.f ? gain : gneg);
        }
        if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        y[i] = o;
    }
}

// float4 variants of modulate / mod_epilogue: grid (chunks of one image, B); no per-element division by the image size
__global__ void __launch_bounds__(kT) modulate_vec_kernel(const float4* __restrict__ x, long long xbs4, const float4* __restrict__ s,
                                                          float4* __restrict__ y, long long per4, int C4, float alpha,
                                                          int round_out) {
    const int b = blockIdx.y;
    const float4* xb = x + (long long)b * xbs4;
    float4* yb = y + (long long)b * per4;
    const float4* sb = s + (long long)b * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = xb[i];
        const float4 m = __ldg(sb + (int)(i % C4));
        float4 o = make_float4(v.x * m.x * alpha, v.y * m.y * alpha, v.z * m.z * alpha, v.w * m.w * alpha);
        if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        yb[i] = o;
    }
}

__global__ void __launch_bounds__(kT) mod_epilogue_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ d,
                                                              const float* __restrict__ noise, const float* __restrict__ nw,
                                                              const float4* __restrict__ bias, float4* __restrict__ y,
                                                              long long P, int C4, float slope, float gain, int round_out) {
    const int b = blockIdx.y;
    const long long per4 = P * C4;
    const float4* xb = x + (long long)b * per4;
    float4* yb = y + (long long)b * per4;
    const float w = (noise && nw) ? __ldg(nw) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const long long pp = i / C4;
        float4 t = xb[i];
        if (d) { const float4 m = __ldg(d + (long long)b * C4 + c); t.x *= m.x; t.y *= m.y; t.z *= m.z; t.w *= m.w; }
        if (noise) { const float nz = __ldg(noise + (long long)b * P + pp) * w; t.x += nz; t.y += nz; t.z += nz; t.w += nz; }
        if (bias) { const float4 bb = __ldg(bias + c); t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w; }
        t.x = (t.x > 0.f ? t.x : t.x * slope) * gain; t.y = (t.y > 0.f ? t.y : t.y * slope) * gain;
        t.z = (t.z > 0.f ? t.z : t.z * slope) * gain; t.w = (t.w > 0.f ? t.w : t.w * slope) * gain;
        if (round_out) { t.x = round_tf32(t.x); t.y = round_tf32(t.y); t.z = round_tf32(t.z); t.w = round_tf32(t.w); }
        yb[i] = t;
    }
}

// ------------------------------------------------------------------------------------------------ modulation
// y[b, p, c] = x[b * xbs + p * C + c] * s[b, c] * alpha          (xbs = 0 broadcasts ConstantInput over the batch)
__global__ void __launch_bounds__(kT) modulate_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ s,
                                                      float* __restrict__ y, int B, long long P, int C, float alpha,
                                                      int round_out) {
    const long long per = P * C, total = per * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const long long r = i - (long long)b * per;
        const int c = (int)(r % C);
        const float v = x[(long long)b * xbs + r] * __ldg(s + (long long)b * C + c) * alpha;
        y[i] = round_out ? round_tf32(v) : v;
    }
}

// out[b, c] += sum_p a[b, p, c] * w[b * wbs + p * C + c]   (out zeroed by the host); block = 32 channels x 8 p-lanes
__global__ void __launch_bounds__(kT) mul_reduce_kernel(const float* __restrict__ a, const float* __restrict__ w, long long wbs,
                                                        float* __restrict__ out, long long P, int C, long long p_per_z) {
    __shared__ float sm[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int lane_p = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const long long p0 = (long long)blockIdx.z * p_per_z;
    long long p1 = p0 + p_per_z;
    if (p1 > P) p1 = P;
    float acc = 0.f;
    if (c < C) {
        const float* ab = a + (long long)b * P * C + c;
        const float* wb = w + (long long)b * wbs + c;
        for (long long p = p0 + lane_p; p < p1; p += 8) acc = fmaf(ab[p * C], wb[p * C], acc);
    }
    sm[lane_p][threadIdx.x & 31] = acc;
    __syncthreads();
    if (lane_p == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
        atomicAdd(out + (long long)b * C + c, t);
    }
}

// y[b,p,c] = lrelu(x[b,p,c] * d[b,c] + noise[b,p] * nw[0] + bias[c]) * gain
__global__ void __launch_bounds__(kT) mod_epilogue_kernel(const float* __restrict__ x, const float* __restrict__ d,
                                                          const float* __restrict__ noise, const float* __restrict__ nw,
                                                          const float* __restrict__ bias, float* __restrict__ y, int B,
                                                          long long P, int C, float slope, float gain, int round_out) {
    const long long per = P * C, total = per * B;
    const float w = (noise && nw) ? __ldg(nw) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const long long r = i - (long long)b * per;
        const int c = (int)(r % C);
        const long long pp = r / C;
        float t = x[i];
        if (d) t *= __ldg(d + (long long)b * C + c);
        if (noise) t = fmaf(__ldg(noise + (long long)b * P + pp), w, t);
        if (bias) t += __ldg(bias + c);
        t = (t > 0.f ? t : t * slope) * gain;
        y[i] = round_out ? round_tf32(t) : t;
    }
}

// out[0] += sum_{rows} noise[row] * sum_c g[row, c]      (one warp per row; out zeroed by the host)
__global__ void __launch_bounds__(kT) noise_grad_kernel(const float* __restrict__ g, const float* __restrict__ noise,
                                                        float* __restrict__ out, long long rows, int C) {
    __shared__ float sm[kT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (long long r = (long long)blockIdx.x * (kT / 32) + warp; r < rows; r += (long long)gridDim.x * (kT / 32)) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += g[r * C + c];
        s = warp_sum(s);
        acc = fmaf(s, __ldg(noise + r), acc);
    }
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kT / 32; ++i) t += sm[i];
        atomicAdd(out, t);
    }
}

// ------------------------------------------------------------------------------------------------ minibatch stddev
// x is [B, F] (F = H*W*C, NHWC flattening), group size G = min(B, 4), M = B / G, sample b = g * M + m
// (discriminator.py:24-27 `input.view(group, -1, ...)`).  std[m] = mean_f sqrt(var_g(x[:, m, f]) + 1e-8).
constexpr float kStdEps = 1e-8f;

__global__ void __launch_bounds__(kT) stddev_fwd_kernel(const float* __restrict__ x, float* __restrict__ std, int G, int M,
                                                        long long F) {
    __shared__ float scratch[32];
    const int m = blockIdx.y;
    float acc[1] = {0.f};
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < F; f += (long long)gridDim.x * blockDim.x) {
        float v[4], mu = 0.f;
        for (int g = 0; g < G; ++g) { v[g] = x[((long long)g * M + m) * F + f]; mu += v[g]; }
        mu /= G;
        float var = 0.f;
        for (int g = 0; g < G; ++g) var += (v[g] - mu) * (v[g] - mu);
        acc[0] += sqrtf(var / G + kStdEps);
    }
    block_sum<1>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(std + m, acc[0] / (float)F);
}

// dx[g, m, f] = dstd[m] / F * (x - mu) / (G * s)
__global__ void __launch_bounds__(kT) stddev_bwd_kernel(const float* __restrict__ dstd, const float* __restrict__ x,
                                                        float* __restrict__ dx, int G, int M, long long F) {
    const int m = blockIdx.y;
    const float gs = __ldg(dstd + m) / (float)F;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < F; f += (long long)gridDim.x * blockDim.x) {
        float v[4], mu = 0.f;
        for (int g = 0; g < G; ++g) { v[g] = x[((long long)g * M + m) * F + f]; mu += v[g]; }
        mu /= G;
        float var = 0.f;
        for (int g = 0; g < G; ++g) var += (v[g] - mu) * (v[g] - mu);
        const float s = sqrtf(var / G + kStdEps);
        for (int g = 0; g < G; ++g) dx[((long long)g * M + m) * F + f] = gs * (v[g] - mu) / (G * s);
    }
}

// Backward of stddev_bwd w.r.t. (dstd, x) for an incoming cotangent gg[B, F] of dx:
//   d_dstd[m] = sum_{g,f} gg * c_g / (G s F)                                   with c_g = x_g - mu
//   d_x[g',m,f] = dstd[m] / (F G) * ((gg_{g'} - mean_g gg) / s - (sum_g gg_g c_g) c_{g'} / (G s^3))
__global__ void __launch_bounds__(kT) stddev_bwd_bwd_kernel(const float* __restrict__ gg, const float* __restrict__ dstd,
                                                            const float* __restrict__ x, float* __restrict__ d_dstd,
                                                            float* __restrict__ d_x, int G, int M, long long F) {
    __shared__ float scratch[32];
    const int m = blockIdx.y;
    const float ds = __ldg(dstd + m);
    float acc[1] = {0.f};
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < F; f += (long long)gridDim.x * blockDim.x) {
        float v[4], q[4], mu = 0.f, qm = 0.f;
        for (int g = 0; g < G; ++g) {
            v[g] = x[((long long)g * M + m) * F + f]; mu += v[g];
            q[g] = gg[((long long)g * M + m) * F + f]; qm += q[g];
        }
        mu /= G; qm /= G;
        float var = 0.f, a = 0.f;
        for (int g = 0; g < G; ++g) { const float c = v[g] - mu; var += c * c; a += q[g] * c; }
        const float s = sqrtf(var / G + kStdEps);
        acc[0] += a / (G * s);
        const float k = ds / ((float)F * G);
        for (int g = 0; g < G; ++g)
            d_x[((long long)g * M + m) * F + f] = k * ((q[g] - qm) / s - a * (v[g] - mu) / (G * s * s * s));
    }
    block_sum<1>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(d_dstd + m, acc[0] / (float)F);
}

// y[b, p, 0:C] = x[b, p, :], y[b, p, C] = std[b % M], y[b, p, C+1:Cp] = 0
__global__ void __launch_bounds__(kT) stddev_concat_kernel(const float* __restrict__ x, const float* __restrict__ std,
                                                           float* __restrict__ y, int B, int M, long long P, int C, int Cp,
                                                           int round_out) {
    const long long total = (long long)B * P * Cp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const long long row = i / Cp;
        const int b = (int)(row / P);
        float v = 0.f;
        if (c < C) v = x[row * C + c];
        else if (c == C) v = __ldg(std + (b % M));
        y[i] = round_out ? round_tf32(v) : v;
    }
}

// dx[b, p, :] = dy[b, p, 0:C];  dstd[m] += sum_{g, p} dy[g*M+m, p, C]      (dstd zeroed by the host)
__global__ void __launch_bounds__(kT) stddev_split_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                          float* __restrict__ dstd, int B, int M, long long P, int C, int Cp) {
    const long long total = (long long)B * P * Cp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const long long row = i / Cp;
        if (c < C) dx[row * C + c] = dy[i];
        else if (c == C) atomicAdd(dstd + ((int)(row / P) % M), dy[i]);
    }
}

// ------------------------------------------------------------------------------------------------ layout changes
// y[b, h, w, c] = c < 3 ? x[b, c, h, w] * scale + shift : 0                    (NCHW image -> padded NHWC)
__global__ void __launch_bounds__(kT) rgb_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, long long HW,
                                                         int cpad, float scale, float shift, int round_out) {
    const long long total = (long long)B * HW * cpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpad);
        const long long row = i / cpad;
        float v = 0.f;
        if (c < 3) {
            const long long b = row / HW, p = row - b * HW;
            v = fmaf(x[(b * 3 + c) * HW + p], scale, shift);
            if (round_out) v = round_tf32(v);
        }
        y[i] = v;
    }
}

// out[b, c, h, w] = src[b, h, w, c] * scale (+ res[b, c, h, w]),  c < 3        (padded NHWC -> NCHW image)
__global__ void __launch_bounds__(kT) nhwc_to_rgb_kernel(const float* __restrict__ src, const float* __restrict__ res,
                                                         float* __restrict__ out, int B, long long HW, int cpad, float scale) {
    const long long total = (long long)B * 3 * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i % HW;
        const long long bc = i / HW;
        const int c = (int)(bc % 3);
        const long long b = bc / 3;
        float v = src[(b * HW + p) * cpad + c] * scale;
        if (res) v += res[i];
        out[i] = v;
    }
}

// PixelNorm (layers.py:15-20): y[r, :] = x[r, :] * rsqrt(mean(x[r, :]^2) + 1e-8); one warp per row
__global__ void __launch_bounds__(kT) pixelnorm_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int d,
                                                       int round_out) {
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * (kT / 32) + (threadIdx.x >> 5); r < rows; r += gridDim.x * (kT / 32)) {
        float s = 0.f;
        for (int c = lane; c < d; c += 32) { const float v = x[(long long)r * d + c]; s = fmaf(v, v, s); }
        s = warp_sum(s);
        const float k = rsqrtf(s / (float)d + 1e-8f);
        for (int c = lane; c < d; c += 32) {
            const float v = x[(long long)r * d + c] * k;
            y[(long long)r * d + c] = round_out ? round_tf32(v) : v;
        }
    }
}

// out[b] = sum_i x[b, i]^2      (one CTA per (row, chunk); out zeroed by the host)
__global__ void __launch_bounds__(kT) row_sqsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
    __shared__ float scratch[32];
    const float* xb = x + (long long)blockIdx.y * n;
    float acc[1] = {0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc[0] = fmaf(xb[i], xb[i], acc[0]);
    block_sum<1>(acc, scratch);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, acc[0]);
}

// y[b, i] = x[b, i] * s[b] * alpha
__global__ void __launch_bounds__(kT) row_scale_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                       float* __restrict__ y, long long n, float alpha) {
    const float k = __ldg(s + blockIdx.y) * alpha;
    const float* xb = x + (long long)blockIdx.y * n;
    float* yb = y + (long long)blockIdx.y * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        yb[i] = xb[i] * k;
}

// out = alpha * a + beta * b + gamma     (b optional)
__global__ void __launch_bounds__(kT) axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                   long long n, float alpha, float beta, float gamma, int round_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = fmaf(alpha, a[i], gamma);
        if (b) v = fmaf(beta, b[i], v);
        out[i] = round_out ? round_tf32(v) : v;
    }
}

// y[row, 0:C] = x[row, :], y[row, C:Cp] = 0   (zero-extends the channel dimension: the tensor-core weight-gradient kernels
// tile the output channels by 128, layers with 32 / 64 channels present their dY zero-padded)
__global__ void __launch_bounds__(kT) pad_channels_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows,
                                                          int C4, int Cp4) {
    const long long total = rows * Cp4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp4);
        const long long r = i / Cp4;
        y4[i] = c < C4 ? __ldg(x4 + r * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------ EMA of G
constexpr int kMaxEma = 64;
constexpr int kEmaElemsPerCta = kT * 16;
struct EmaBatch {
    float* dst[kMaxEma];
    const float* src[kMaxEma];
    long long numel[kMaxEma];
    int cta_begin[kMaxEma + 1];
    int n;
    float decay;
};

__global__ void __launch_bounds__(kT) ema_kernel(const __grid_constant__ EmaBatch e) {
    int l = 0;
#pragma unroll 1
    while (l + 1 < e.n && (int)blockIdx.x >= e.cta_begin[l + 1]) ++l;
    const long long base = (long long)(blockIdx.x - e.cta_begin[l]) * kEmaElemsPerCta;
    const long long n = e.numel[l];
    float* __restrict__ d = e.dst[l];
    const float* __restrict__ s = e.src[l];
    for (int it = 0; it < 16; ++it) {
        const long long i = base + (long long)it * kT + threadIdx.x;
        if (i < n) d[i] = fmaf(e.decay, d[i], (1.f - e.decay) * s[i]);
    }
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int cb200_upfirdn2d(const float* x, const long long* x_strides, float* y, const long long* y_strides,
                               const float* fir, int N, int C, int Hi, int Wi, int Ho, int Wo, int up, int down, int pad_x0,
                               int pad_y0, int kh, int kw, int flip, float gain, int c_fast, int round_out, void* stream) {
    CB200_CHECK_ARG(N > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "upfirdn2d: empty tensor");
    CB200_CHECK_ARG(up >= 1 && down >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 64, "upfirdn2d: up/down >= 1, FIR at most 64 taps");
    UpfirdnParams p;
    p.x = x; p.y = y; p.k = fir;
    p.xs_n = x_strides[0]; p.xs_c = x_strides[1]; p.xs_h = x_strides[2]; p.xs_w = x_strides[3];
    p.ys_n = y_strides[0]; p.ys_c = y_strides[1]; p.ys_h = y_strides[2]; p.ys_w = y_strides[3];
    p.N = N; p.C = C; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo; p.up = up; p.down = down; p.px0 = pad_x0; p.py0 = pad_y0;
    p.kh = kh; p.kw = kw; p.flip = flip; p.c_fast = c_fast; p.round_out = round_out; p.gain = gain;
    const bool nhwc_contig = c_fast && C % 4 == 0 && p.xs_c == 1 && p.ys_c == 1 && p.xs_w == C && p.ys_w == C &&
                             p.xs_h == (long long)Wi * C && p.ys_h == (long long)Wo * C && p.xs_n == (long long)Hi * Wi * C &&
                             p.ys_n == (long long)Ho * Wo * C &&
                             ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (nhwc_contig) {
        long long tiles = (long long)N * ((Ho + kTile - 1) / kTile) * ((Wo + kTile - 1) / kTile);
        if (tiles > 148LL * 64) tiles = 148LL * 64;        // (capping at 8 CTAs / SM measured 20 % slower: fewer loads in flight)
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (kh == 4 && kw == 4 && up == 1) upfirdn2d_nhwc4_kernel<1, 4><<<(unsigned)tiles, kT, 0, st>>>(p);
        else if (kh == 4 && kw == 4 && up == 2) upfirdn2d_nhwc4_kernel<2, 4><<<(unsigned)tiles, kT, 0, st>>>(p);
        else if (up == 1) upfirdn2d_nhwc4_kernel<1, 0><<<(unsigned)tiles, kT, 0, st>>>(p);
        else upfirdn2d_nhwc4_kernel<0, 0><<<(unsigned)tiles, kT, 0, st>>>(p);
    }
    else
        upfirdn2d_kernel<<<grid_for((long long)N * C * Ho * Wo), kT, 0, static_cast<cudaStream_t>(stream)>>>(p);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("upfirdn2d");
    return CB200_OK;
}

extern "C" int cb200_patch_s2_gather(const float* x, float* u, int B, int Ho, int Wo, int C, int round_out, void* stream) {
    CB200_CHECK_ARG(B > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 4 == 0, "patch_s2_gather: C must be a positive multiple of 4");
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(u)) & 15) == 0, "patch_s2_gather: 16-byte alignment");
    patch_s2_gather_kernel<<<grid_for((long long)B * Ho * Wo * 9 * (C / 4)), kT, 0, static_cast<cudaStream_t>(stream)>>>(
        x, u, B, Ho, Wo, C / 4, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("patch_s2_gather");
    return CB200_OK;
}

extern "C" int cb200_patch_s2_scatter(const float* u, float* x, int B, int Ho, int Wo, int C, int round_out, void* stream) {
    CB200_CHECK_ARG(B > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 4 == 0, "patch_s2_scatter: C must be a positive multiple of 4");
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(u)) & 15) == 0, "patch_s2_scatter: 16-byte alignment");
    patch_s2_scatter_kernel<<<grid_for((long long)B * (2 * Ho + 1) * (2 * Wo + 1) * (C / 4)), kT, 0,
                              static_cast<cudaStream_t>(stream)>>>(u, x, B, Ho, Wo, C / 4, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("patch_s2_scatter");
    return CB200_OK;
}

extern "C" int cb200_bias_act(const float* x, const float* bias, const float* ref, const float* res, float* y, long long n,
                              int C, int mode, float slope, float gain, int round_out, void* stream) {
    CB200_CHECK_ARG(n > 0 && C > 0 && n % C == 0, "bias_act: element count must be a positive multiple of C");
    CB200_CHECK_ARG(mode == 0 || (mode == 1 && ref != nullptr), "bias_act: mode 1 needs the reference activation");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (C % 4 == 0 && al16(x) && al16(y) && al16(bias) && al16(ref) && al16(res)) {
        bias_act_vec_kernel<<<grid_for(n / 4), kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(bias), reinterpret_cast<const float4*>(ref),
            reinterpret_cast<const float4*>(res), reinterpret_cast<float4*>(y), n / 4, C / 4, mode, slope, gain, round_out);
    } else {
        bias_act_kernel<<<grid_for(n), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, bias, ref, res, y, n, C, mode, slope, gain,
                                                                                   round_out);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("bias_act");
    return CB200_OK;
}

extern "C" int cb200_modulate(const float* x, long long x_batch_stride, const float* s, float* y, int B, long long P, int C,
                              float alpha, int round_out, void* stream) {
    CB200_CHECK_ARG(B > 0 && P > 0 && C > 0, "modulate: empty tensor");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (C % 4 == 0 && x_batch_stride % 4 == 0 && B <= 65535 && al16(x) && al16(y) && al16(s)) {
        const long long per4 = P * (C / 4);
        dim3 grid(grid_for(per4, kT, (148 * 16 + B - 1) / B), B);
        modulate_vec_kernel<<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(x), x_batch_stride / 4, reinterpret_cast<const float4*>(s),
            reinterpret_cast<float4*>(y), per4, C / 4, alpha, round_out);
    } else {
        modulate_kernel<<<grid_for((long long)B * P * C), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, x_batch_stride, s, y, B, P,
                                                                                                     C, alpha, round_out);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("modulate");
    return CB200_OK;
}

extern "C" int cb200_mul_reduce(const float* a, const float* w, long long w_batch_stride, float* out, int B, long long P, int C,
                                void* stream) {
    CB200_CHECK_ARG(B > 0 && P > 0 && C > 0 && B <= 65535, "mul_reduce: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C, st);
    if (e != cudaSuccess) { cb200_set_error("mul_reduce: memset: %s", cudaGetErrorString(e)); return (int)e; }
    long long z = (P + 511) / 512;
    if (z > 64) z = 64;
    const long long p_per_z = (P + z - 1) / z;
    dim3 grid((C + 31) / 32, B, (unsigned)z);
    mul_reduce_kernel<<<grid, kT, 0, st>>>(a, w, w_batch_stride, out, P, C, p_per_z);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("mul_reduce");
    return CB200_OK;
}

extern "C" int cb200_mod_epilogue(const float* x, const float* demod, const float* noise, const float* noise_weight,
                                  const float* bias, float* y, int B, long long P, int C, float slope, float gain, int round_out,
                                  void* stream) {
    CB200_CHECK_ARG(B > 0 && P > 0 && C > 0, "mod_epilogue: empty tensor");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (C % 4 == 0 && B <= 65535 && al16(x) && al16(y) && al16(demod) && al16(bias)) {
        dim3 grid(grid_for(P * (C / 4), kT, (148 * 16 + B - 1) / B), B);
        mod_epilogue_vec_kernel<<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(demod), noise, noise_weight,
            reinterpret_cast<const float4*>(bias), reinterpret_cast<float4*>(y), P, C / 4, slope, gain, round_out);
    } else {
        mod_epilogue_kernel<<<grid_for((long long)B * P * C), kT, 0, static_cast<cudaStream_t>(stream)>>>(
            x, demod, noise, noise_weight, bias, y, B, P, C, slope, gain, round_out);
    }
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("mod_epilogue");
    return CB200_OK;
}

extern "C" int cb200_noise_grad(const float* g, const float* noise, float* out1, long long rows, int C, void* stream) {
    CB200_CHECK_ARG(rows > 0 && C > 0, "noise_grad: empty tensor");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out1, 0, sizeof(float), st);
    if (e != cudaSuccess) { cb200_set_error("noise_grad: memset: %s", cudaGetErrorString(e)); return (int)e; }
    noise_grad_kernel<<<grid_for(rows, kT / 32, 148 * 4), kT, 0, st>>>(g, noise, out1, rows, C);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("noise_grad");
    return CB200_OK;
}

static int stddev_shape(int B, int* G, int* M) {
    *G = B < 4 ? B : 4;
    if (B <= 0 || B % *G != 0) return 1;
    *M = B / *G;
    return *M > 65535;
}

extern "C" int cb200_stddev_fwd(const float* x, float* std, int B, long long F, void* stream) {
    int G, M;
    CB200_CHECK_ARG(!stddev_shape(B, &G, &M) && F > 0, "stddev_fwd: batch %d must be a multiple of min(batch, 4)", B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(std, 0, sizeof(float) * M, st);
    if (e != cudaSuccess) { cb200_set_error("stddev_fwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(grid_for(F, kT, 64), M);
    stddev_fwd_kernel<<<grid, kT, 0, st>>>(x, std, G, M, F);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("stddev_fwd");
    return CB200_OK;
}

extern "C" int cb200_stddev_bwd(const float* dstd, const float* x, float* dx, int B, long long F, void* stream) {
    int G, M;
    CB200_CHECK_ARG(!stddev_shape(B, &G, &M) && F > 0, "stddev_bwd: batch %d must be a multiple of min(batch, 4)", B);
    dim3 grid(grid_for(F, kT, 64), M);
    stddev_bwd_kernel<<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(dstd, x, dx, G, M, F);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("stddev_bwd");
    return CB200_OK;
}

extern "C" int cb200_stddev_bwd_bwd(const float* gg, const float* dstd, const float* x, float* d_dstd, float* d_x, int B,
                                    long long F, void* stream) {
    int G, M;
    CB200_CHECK_ARG(!stddev_shape(B, &G, &M) && F > 0, "stddev_bwd_bwd: batch %d must be a multiple of min(batch, 4)", B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(d_dstd, 0, sizeof(float) * M, st);
    if (e != cudaSuccess) { cb200_set_error("stddev_bwd_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(grid_for(F, kT, 64), M);
    stddev_bwd_bwd_kernel<<<grid, kT, 0, st>>>(gg, dstd, x, d_dstd, d_x, G, M, F);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("stddev_bwd_bwd");
    return CB200_OK;
}

extern "C" int cb200_stddev_concat(const float* x, const float* std, float* y, int B, long long P, int C, int Cp, int round_out,
                                   void* stream) {
    int G, M;
    CB200_CHECK_ARG(!stddev_shape(B, &G, &M) && P > 0 && C > 0 && Cp > C, "stddev_concat: bad shape");
    stddev_concat_kernel<<<grid_for((long long)B * P * Cp), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, std, y, B, M, P, C, Cp,
                                                                                                       round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("stddev_concat");
    return CB200_OK;
}

extern "C" int cb200_stddev_split(const float* dy, float* dx, float* dstd, int B, long long P, int C, int Cp, void* stream) {
    int G, M;
    CB200_CHECK_ARG(!stddev_shape(B, &G, &M) && P > 0 && C > 0 && Cp > C, "stddev_split: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(dstd, 0, sizeof(float) * M, st);
    if (e != cudaSuccess) { cb200_set_error("stddev_split: memset: %s", cudaGetErrorString(e)); return (int)e; }
    stddev_split_kernel<<<grid_for((long long)B * P * Cp), kT, 0, st>>>(dy, dx, dstd, B, M, P, C, Cp);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("stddev_split");
    return CB200_OK;
}

extern "C" int cb200_rgb_to_nhwc(const float* x, float* y, int B, int H, int W, int cpad, float scale, float shift, int round_out,
                                 void* stream) {
    CB200_CHECK_ARG(B > 0 && H > 0 && W > 0 && cpad >= 3, "rgb_to_nhwc: bad shape");
    rgb_to_nhwc_kernel<<<grid_for((long long)B * H * W * cpad), kT, 0, static_cast<cudaStream_t>(stream)>>>(
        x, y, B, (long long)H * W, cpad, scale, shift, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("rgb_to_nhwc");
    return CB200_OK;
}

extern "C" int cb200_nhwc_to_rgb(const float* src, const float* res, float* out, int B, int H, int W, int cpad, float scale,
                                 void* stream) {
    CB200_CHECK_ARG(B > 0 && H > 0 && W > 0 && cpad >= 3, "nhwc_to_rgb: bad shape");
    nhwc_to_rgb_kernel<<<grid_for((long long)B * 3 * H * W), kT, 0, static_cast<cudaStream_t>(stream)>>>(
        src, res, out, B, (long long)H * W, cpad, scale);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("nhwc_to_rgb");
    return CB200_OK;
}

extern "C" int cb200_pixelnorm(const float* x, float* y, int rows, int d, int round_out, void* stream) {
    CB200_CHECK_ARG(rows > 0 && d > 0, "pixelnorm: empty input");
    pixelnorm_kernel<<<grid_for(rows, kT / 32), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rows, d, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("pixelnorm");
    return CB200_OK;
}

extern "C" int cb200_row_sqsum(const float* x, float* out, int B, long long n, void* stream) {
    CB200_CHECK_ARG(B > 0 && B <= 65535 && n > 0, "row_sqsum: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * B, st);
    if (e != cudaSuccess) { cb200_set_error("row_sqsum: memset: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(grid_for(n, kT * 4, 64), B);
    row_sqsum_kernel<<<grid, kT, 0, st>>>(x, out, n);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("row_sqsum");
    return CB200_OK;
}

extern "C" int cb200_row_scale(const float* x, const float* s, float* y, int B, long long n, float alpha, void* stream) {
    CB200_CHECK_ARG(B > 0 && B <= 65535 && n > 0, "row_scale: bad shape");
    dim3 grid(grid_for(n, kT * 4, 64), B);
    row_scale_kernel<<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(x, s, y, n, alpha);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("row_scale");
    return CB200_OK;
}

extern "C" int cb200_axpby(const float* a, const float* b, float* out, long long n, float alpha, float beta, float gamma,
                           int round_out, void* stream) {
    CB200_CHECK_ARG(n > 0, "axpby: empty input");
    axpby_kernel<<<grid_for(n), kT, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n, alpha, beta, gamma, round_out);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("axpby");
    return CB200_OK;
}

extern "C" int cb200_pad_channels(const float* x, float* y, long long rows, int C, int Cp, void* stream) {
    CB200_CHECK_ARG(rows > 0 && C > 0 && Cp >= C && C % 4 == 0 && Cp % 4 == 0, "pad_channels: C, Cp must be multiples of 4, Cp >= C");
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "pad_channels: 16-byte alignment");
    pad_channels_kernel<<<grid_for(rows * (Cp / 4)), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rows, C / 4, Cp / 4);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH("pad_channels");
    return CB200_OK;
}

extern "C" int cb200_ema_lerp(const struct cb200_ema_tensor* tensors, int n, float decay, void* stream) {
    CB200_CHECK_ARG(n >= 0, "ema_lerp: negative tensor count");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int base = 0; base < n; base += kMaxEma) {
        EmaBatch e;
        e.n = (n - base < kMaxEma) ? n - base : kMaxEma;
        e.decay = decay;
        int ctas = 0;
        for (int i = 0; i < e.n; ++i) {
            const cb200_ema_tensor& t = tensors[base + i];
            CB200_CHECK_ARG(t.numel > 0 && t.dst && t.src, "ema_lerp: empty tensor in the table");
            e.dst[i] = t.dst; e.src[i] = t.src; e.numel[i] = t.numel;
            e.cta_begin[i] = ctas;
            ctas += (int)((t.numel + kEmaElemsPerCta - 1) / kEmaElemsPerCta);
        }
        e.cta_begin[e.n] = ctas;
        ema_kernel<<<ctas, kT, 0, st>>>(e);
        CB200_COUNT_LAUNCH();
        CB200_CHECK_LAUNCH("ema_lerp");
    }
    return CB200_OK;
}
