// tcgen05 / TMA "tap GEMM" for sm_100a: one warp-specialised kernel that covers every dense
// contraction whose A operand is a (shifted) box of an NHWC tensor:
//
//     out[row(m), n] = epilogue( sum_{tap t} sum_{c} A[m shifted by tap t, c] * Bw[n, t*C + c] )
//
//   * 3x3 stride-1 convolution forward and its data-gradient            (SNDCGAN D layers 3,5,7)
//   * 4x4 stride-2 convolution forward (space-to-depth view, no copies)  (layers 2,4,6)
//   * 4x4 stride-2 data-gradient == ConvTranspose2d(4,2,1) forward, as 4 output-parity classes
//     of 2x2-tap convolutions (blockIdx.z = class)                       (D dgrad, G forward)
//   * plain C = A * B^T GEMMs (the MLP heads)                            (1 tap, 2-D A)
//
// Replaces the cuDNN implicit-GEMM / cuBLAS calls behind nn.Conv2d / nn.Linear of the reference
// (models/gan/sndcgan.py:91-109, models/gan/base.py:14-35,92-101).
//
// Design (Blackwell-native): operands are fp32 in HBM and are fed to the tensor core as TF32
// (kind::tf32, fp32 accumulate in TMEM).  A tile = 128 output pixels x 32 channels (one 128-byte
// swizzle span) is fetched by ONE TMA box load per k-block; im2col never exists - the tap shift is
// a coordinate offset of the box and zero padding is TMA out-of-bounds fill.  B tile = BN x 32 of
// the K-major weight matrix.  Warp roles: warp0 = TMA producer, warp1 = MMA issuer (single elected
// thread), warp2 = TMEM allocator, warps4-7 = epilogue (tcgen05.ld -> bias / LeakyReLU / lrelu'
// mask / TF32 rounding -> global).  smem ring of kStages (full/empty mbarriers); tcgen05.commit
// releases ring slots and signals the epilogue.
//
// Roofline: tensor pipe.  kind::tf32 M=128,N=BN,K=8 per instruction; algorithmic FLOPs =
// 2*M*N*K_total per launch (DESIGN.md lists them per layer).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                       // fp32 elements = 128 bytes = one swizzle span
constexpr int kATileBytes = kBlockM * kBlockK * 4;   // 16 KB
constexpr int kMaxTaps = 16;
constexpr int kThreads = 256;

struct TapGemmParams {
    CUtensorMap tmap_a;
    CUtensorMap tmap_b;
    int box[4];           // A box extent along spatial dims 1..4 (product == 128)
    int tiles[4];         // tile counts along dims 1..4
    int extent[4];        // valid extent along dims 1..4 (rows beyond are not stored)
    long long ostride[4]; // output-row stride of each dim
    long long cls_off[4]; // output-row offset of each class
    int tap[4][kMaxTaps][5];   // [class][tap] -> {c_add, d1, d2, d3, d4}
    int ntaps, cblocks;   // k-blocks = ntaps * cblocks
    int ldo;              // floats between consecutive output rows
    int b_rows_per_cls;   // row offset into the B map per class
    float* out;
    const float* bias;    // [N] or null
    const float* dact;    // same addressing as out; multiplies by (dact > 0 ? 1 : slope); or null
    float slope;          // LeakyReLU slope applied after bias (1.0 = identity) when dact == null
    int round_out;        // round outputs to TF32 (they feed another tensor-core GEMM)
    float* colsum;        // [N] or null: += column sums of the stored outputs (bias gradient of the producing layer)
    int n_total;          // N (columns of the whole problem)
    // split-K (single-CTA kernel only): gridDim.z = classes * splits; split s accumulates k-blocks
    // [s * kb_per_split, (s+1) * kb_per_split) and parks its raw 128 x BN partial in `ws`; the LAST split to arrive at a
    // tile (per-tile arrival counter) adds the partials in split order and runs the ordinary epilogue.
    int splits, kb_per_split, classes;
    float* ws;            // [splits][units][128 * BN] fp32, units = gridDim.x * gridDim.y * classes
    int* counters;        // [units], zero between launches (the finishing CTA resets its counter)
    int debug;            // profiling experiments only (env CB200_TAPGEMM_DEBUG): 1 no stores, 2 no MMA, 4 no A loads, 8 no B loads
};

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int kBTileBytes = BN * kBlockK * 4;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// Epilogue of one 128-row tile.  tcgen05.ld hands every thread one accumulator ROW (32 columns at a time); writing
// rows straight to global memory would make each warp store touch 32 different lines (measured: ~30 % of the
// kernel).  Instead each epilogue warp transposes its 32x32 block through shared memory (the pipeline's stage-0
// buffer, free once tmem_full has fired) so that 8 lanes cover one contiguous 128-byte row segment: bias /
// LeakyReLU / lrelu' mask / TF32 rounding are applied in the coalesced domain, loads of `dact` included.
constexpr int kEpiStride = 36;                        // floats per staged row: 16-byte aligned, conflict-free float4
constexpr int kEpiWarpFloats = 32 * kEpiStride;       // 4.5 KB per epilogue warp

// MODE 0: accumulators from TMEM -> epilogue -> global.  MODE 1 (split-K): raw accumulators -> `ws_tile` (row-major
// 128 x BN, coalesced).  MODE 2 (split-K finisher): sum of `ws_splits` partials (stride `ws_stride` floats) -> epilogue.
template <int BN, int MODE = 0>
__device__ __forceinline__ void epilogue_rows_nowait(const TapGemmParams& p, const int (&base)[4], int cls, int n0,
                                                     uint32_t tmem_base, int warp, int lane, float* stage_buf,
                                                     float* cs_smem = nullptr, float* ws_tile = nullptr,
                                                     int ws_splits = 0, size_t ws_stride = 0) {
    const int q = warp & 3;
    float* st = stage_buf + q * kEpiWarpFloats;
    int r = q * 32 + lane;
    bool valid = true;
    long long orow = p.cls_off[cls];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        int rd = r % p.box[d];
        r /= p.box[d];
        int x = base[d] + rd;
        valid = valid && (x < p.extent[d]);
        orow += (long long)x * p.ostride[d];
    }
    const long long my_off = valid ? orow * p.ldo : -1;        // element offset of this lane's row, -1 = masked
    const int sub = lane >> 3;                                  // row within a group of 4
    const int col4 = (lane & 7) * 4;                            // first of this lane's 4 columns within the chunk
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        if constexpr (MODE != 2) {
            uint32_t v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            tc::tmem_ld_wait();
            if (p.debug & 1) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(st + lane * kEpiStride + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            __syncwarp();
        }
        if constexpr (MODE == 1) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = it * 4 + sub;
                *reinterpret_cast<float4*>(ws_tile + (size_t)(q * 32 + row) * BN + c0 + col4) =
                    *reinterpret_cast<const float4*>(st + row * kEpiStride + col4);
            }
            __syncwarp();
            continue;
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + col4));
        long long offs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + sub);
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);       // column sums of this lane's 4 columns over its 8 rows
        float4 act[8];
        if (p.dact) {                      // all eight independent loads in flight before the first store
#pragma unroll
            for (int it = 0; it < 8; ++it)
                act[it] = (offs[it] >= 0) ? __ldg(reinterpret_cast<const float4*>(p.dact + offs[it] + n0 + c0 + col4))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + sub;
            float4 o;
            if constexpr (MODE == 2) {
                o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (offs[it] >= 0) {
                    const float* src = ws_tile + (size_t)(q * 32 + row) * BN + c0 + col4;
                    int sp = 0;
                    for (; sp + 4 <= ws_splits; sp += 4) {          // four independent loads in flight, fixed summation order
                        const float4 t0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sp * ws_stride));
                        const float4 t1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sp + 1) * ws_stride));
                        const float4 t2 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sp + 2) * ws_stride));
                        const float4 t3 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(sp + 3) * ws_stride));
                        o.x += t0.x; o.y += t0.y; o.z += t0.z; o.w += t0.w;
                        o.x += t1.x; o.y += t1.y; o.z += t1.z; o.w += t1.w;
                        o.x += t2.x; o.y += t2.y; o.z += t2.z; o.w += t2.w;
                        o.x += t3.x; o.y += t3.y; o.z += t3.z; o.w += t3.w;
                    }
                    for (; sp < ws_splits; ++sp) {
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sp * ws_stride));
                        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                    }
                }
            } else {
                o = *reinterpret_cast<const float4*>(st + row * kEpiStride + col4);
            }
            if (offs[it] >= 0) {
                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                if (p.dact) {
                    const float4 a = act[it];
                    o.x *= (a.x > 0.f) ? 1.f : p.slope; o.y *= (a.y > 0.f) ? 1.f : p.slope;
                    o.z *= (a.z > 0.f) ? 1.f : p.slope; o.w *= (a.w > 0.f) ? 1.f : p.slope;
                } else {
                    o.x = (o.x > 0.f) ? o.x : o.x * p.slope; o.y = (o.y > 0.f) ? o.y : o.y * p.slope;
                    o.z = (o.z > 0.f) ? o.z : o.z * p.slope; o.w = (o.w > 0.f) ? o.w : o.w * p.slope;
                }
                if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                *reinterpret_cast<float4*>(p.out + offs[it] + n0 + c0 + col4) = o;
                cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
            }
        }
        if (p.colsum) {
            // fold the four row groups of the warp (lane bits 3, 4); lanes 0..7 then hold 32-row sums of 4 columns
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8);  cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8);  cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
            if (sub == 0) {
                if (cs_smem) {          // persistent kernel: this WARP's private partial sums, flushed once at the end
                    float4* dst = reinterpret_cast<float4*>(cs_smem + n0 + c0 + col4);
                    float4 t = *dst;
                    t.x += cs.x; t.y += cs.y; t.z += cs.z; t.w += cs.w;
                    *dst = t;
                } else {
                    float* dst = p.colsum + n0 + c0 + col4;
                    atomicAdd(dst + 0, cs.x); atomicAdd(dst + 1, cs.y); atomicAdd(dst + 2, cs.z); atomicAdd(dst + 3, cs.w);
                }
            }
        }
        __syncwarp();
    }
}


// Split-K tail of one CTA (the four epilogue warps): park the raw 128 x BN partial, count the arrival, and let the LAST
// split of this output tile add all partials in split order and run the fused epilogue (deterministic).  `flag` = one
// shared-memory word of the CTA.
template <int BN>
__device__ __forceinline__ void splitk_epilogue(const TapGemmParams& p, const int (&base)[4], int cls, int n0, int split,
                                                uint32_t tmem_base, uint64_t* tmem_full_bar, int warp, int lane,
                                                float* stage_buf, uint32_t* flag) {
    const int units = (int)(gridDim.x * gridDim.y) * p.classes;
    const int unit = (cls * (int)gridDim.y + (int)blockIdx.y) * (int)gridDim.x + (int)blockIdx.x;
    const size_t tile_floats = (size_t)kBlockM * BN;
    tc::mbar_wait(tmem_full_bar, 0);
    tc::fence_after_sync();
    epilogue_rows_nowait<BN, 1>(p, base, cls, n0, tmem_base, warp, lane, stage_buf, nullptr,
                                p.ws + ((size_t)split * units + unit) * tile_floats);
    __threadfence();                                       // partial visible device-wide before the arrival
    asm volatile("bar.sync 1, 128;" ::: "memory");         // the four epilogue warps
    if (threadIdx.x == 128) {
        const int old = atomicAdd(p.counters + unit, 1);
        const int last = (old == p.splits - 1) ? 1 : 0;
        if (last) p.counters[unit] = 0;                    // every split has arrived: re-arm for the next launch
        *flag = (uint32_t)last;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (*flag) {
        __threadfence();
        epilogue_rows_nowait<BN, 2>(p, base, cls, n0, tmem_base, warp, lane, stage_buf, nullptr,
                                    p.ws + (size_t)unit * tile_floats, p.splits, (size_t)units * tile_floats);
    }
}

template <int BN>
__device__ __forceinline__ void epilogue_rows(const TapGemmParams& p, const int (&base)[4], int cls, int n0,
                                              uint32_t tmem_base, uint64_t* tmem_full_bar, int warp, int lane,
                                              float* stage_buf) {
    tc::mbar_wait(tmem_full_bar, 0);
    tc::fence_after_sync();
    epilogue_rows_nowait<BN>(p, base, cls, n0, tmem_base, warp, lane, stage_buf);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 2)    // 2 CTAs per SM (<= 128 registers): the split-K grid is sized for it
tap_gemm_kernel(const __grid_constant__ TapGemmParams p) {
    using L = SmemLayout<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cls = (int)blockIdx.z % p.classes;
    const int split = (int)blockIdx.z / p.classes;
    const int n0 = blockIdx.y * BN;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.ntaps * p.cblocks, kb_begin + p.kb_per_split);

    // tile -> base coordinates along the 4 spatial dims
    int base[4];
    {
        int t = blockIdx.x;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            base[d] = (t % p.tiles[d]) * p.box[d];
            t /= p.tiles[d];
        }
    }

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.tmap_a);
        tc::prefetch_tmap(&p.tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, BN);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const int t = kb / p.cblocks;
                const int cb = kb - t * p.cblocks;
                tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * L::kStageBytes;
                uint8_t* sb = sa + kATileBytes;
                tc::mbar_expect_tx(&full_bar[stage], ((p.debug & 4) ? 0 : kATileBytes) + ((p.debug & 8) ? 0 : L::kBTileBytes));
                const int* tp = p.tap[cls][t];
                if (!(p.debug & 4))
                    tc::tma_load_5d(sa, &p.tmap_a, &full_bar[stage], cb * kBlockK + tp[0], base[0] + tp[1],
                                    base[1] + tp[2], base[2] + tp[3], base[3] + tp[4]);
                if (!(p.debug & 8))
                    tc::tma_load_2d(sb, &p.tmap_b, &full_bar[stage], kb * kBlockK, cls * p.b_rows_per_cls + n0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc::idesc_tf32(kBlockM, BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
            tc::mbar_wait(&full_bar[stage], phase);
            tc::fence_after_sync();
            if (tc::elect_one()) {
                const uint32_t sa = tc::smem_u32(smem + stage * L::kStageBytes);
                const uint32_t sb = sa + kATileBytes;
                if (!(p.debug & 2)) {
#pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k) {
                        const uint64_t adesc = tc::smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t bdesc = tc::smem_desc_sw128(sb + k * 32, 16, 1024);
                        tc::mma_tf32(tmem_base, adesc, bdesc, idesc, (kb > kb_begin || k) ? 1u : 0u);
                    }
                    tc::mma_commit(&empty_bar[stage]);
                    if (kb == kb_end - 1) tc::mma_commit(tmem_full_bar);
                } else {
                    tc::mbar_arrive(&empty_bar[stage]);
                    if (kb == kb_end - 1) tc::mbar_arrive(tmem_full_bar);
                }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4) {
        if (p.splits <= 1) {
            epilogue_rows<BN>(p, base, cls, n0, tmem_base, tmem_full_bar, warp, lane, reinterpret_cast<float*>(smem));
        } else {
            splitk_epilogue<BN>(p, base, cls, n0, split, tmem_base, tmem_full_bar, warp, lane, reinterpret_cast<float*>(smem),
                                tmem_slot + 1);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, BN);
    }
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster own two consecutive 128-row M tiles and issue ONE
// tcgen05.mma with M = 256; each CTA stages its own A tile and HALF of the B tile (BN/2 weight rows), so the
// shared-memory / L2 operand traffic per FLOP is halved with respect to the single-CTA kernel.  The pair
// leader (cluster rank 0) owns the `full` barriers (they collect the TMA bytes of both CTAs) and issues the
// MMAs; tcgen05.commit is multicast to the `empty` / `tmem_full` barriers of both CTAs.
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct SmemLayout2 {
    static constexpr int kBTileBytes = (BN / 2) * kBlockK * 4;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tap_gemm2_kernel(const __grid_constant__ TapGemmParams p) {
    using L = SmemLayout2<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;
    const int cls = (int)blockIdx.z % p.classes;
    const int split = (int)blockIdx.z / p.classes;
    const int n0 = blockIdx.y * BN;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.ntaps * p.cblocks, kb_begin + p.kb_per_split);

    int base[4];
    {
        int t = blockIdx.x;           // grid.x is padded to an even count; surplus tiles fall outside `extent`
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            base[d] = (t % p.tiles[d]) * p.box[d];
            t /= p.tiles[d];
        }
        if (t > 0) base[3] = p.extent[3] + p.box[3] * t;     // beyond the last tile: loads are OOB zeros, stores masked
    }

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.tmap_a);
        tc::prefetch_tmap(&p.tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc2(tmem_slot, BN);
    tc::fence_before_sync();
    tc::cluster_sync_all();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const int t = kb / p.cblocks;
                const int cb = kb - t * p.cblocks;
                tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * L::kStageBytes;
                uint8_t* sb = sa + kATileBytes;
                const uint32_t bar = tc::map_to_cta(tc::smem_u32(&full_bar[stage]), 0);
                if (leader) tc::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                const int* tp = p.tap[cls][t];
                tc::tma2_load_5d(sa, &p.tmap_a, bar, cb * kBlockK + tp[0], base[0] + tp[1], base[1] + tp[2],
                                 base[2] + tp[3], base[3] + tp[4]);
                tc::tma2_load_2d(sb, &p.tmap_b, bar, kb * kBlockK, cls * p.b_rows_per_cls + n0 + (int)rank * (BN / 2));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && leader) {
        constexpr uint32_t idesc = tc::idesc_tf32(256, BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
            tc::mbar_wait(&full_bar[stage], phase);
            tc::fence_after_sync();
            if (tc::elect_one()) {
                const uint32_t sa = tc::smem_u32(smem + stage * L::kStageBytes);
                const uint32_t sb = sa + kATileBytes;
#pragma unroll
                for (int k = 0; k < kBlockK / 8; ++k) {
                    const uint64_t adesc = tc::smem_desc_sw128(sa + k * 32, 16, 1024);
                    const uint64_t bdesc = tc::smem_desc_sw128(sb + k * 32, 16, 1024);
                    tc::mma2_tf32(tmem_base, adesc, bdesc, idesc, (kb > kb_begin || k) ? 1u : 0u);
                }
                tc::mma2_commit_multicast(&empty_bar[stage]);
                if (kb == kb_end - 1) tc::mma2_commit_multicast(tmem_full_bar);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4) {
        if (p.splits <= 1) {
            epilogue_rows<BN>(p, base, cls, n0, tmem_base, tmem_full_bar, warp, lane, reinterpret_cast<float*>(smem));
        } else {
            splitk_epilogue<BN>(p, base, cls, n0, split, tmem_base, tmem_full_bar, warp, lane, reinterpret_cast<float*>(smem),
                                tmem_slot + 1);
        }
    }
    tc::fence_before_sync();
    tc::cluster_sync_all();
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc2(tmem_base, BN);
    }
}


// ------------------------------------------------------------------------------------------------
// Persistent CTA-pair variant: one cluster of two CTAs per SM pair walks a strided list of output super-tiles.
//   * super-tile = 2*MT consecutive 128-row M tiles (MT per CTA) x BN columns; every k-block stages MT A boxes and
//     BN/2 weight rows per CTA -> (MT*16 KB + BN*64 B) per (MT x 128 x BN x 32) MACs: 2.0x (BN=256, MT=1) resp.
//     1.6x (BN=128, MT=2) fewer shared-memory bytes per FLOP than the 128x128 single-CTA tile, which is what the
//     measured ~77 B/clk/SM of L2->SM ingest requires for TF32 operands;
//   * the smem ring runs continuously across tiles (no pipeline drain between tiles);
//   * TMEM holds TWO accumulator sets (2 * MT * BN = 512 columns): the epilogue of tile i (TMEM -> smem transpose
//     -> coalesced stores) overlaps the MMAs of tile i+1 (tmem_full / tmem_empty barriers, the latter collecting
//     the arrivals of the epilogue warps of BOTH CTAs at the pair leader).
// ------------------------------------------------------------------------------------------------
constexpr int kColsumMax = 512;     // widest N whose fused column sums are kept per CTA in shared memory

template <int BN, int MT, int STAGES>
struct SmemLayoutP {
    static constexpr int kATile = MT * kATileBytes;
    static constexpr int kBTile = (BN / 2) * kBlockK * 4;
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kEpiOffset = STAGES * kStageBytes;
    static constexpr int kCsCols = (BN <= 64) ? 64 : kColsumMax;                    // columns kept per epilogue warp
    static constexpr int kColsumOffset = kEpiOffset + 4 * kEpiWarpFloats * 4;       // [4 warps][kCsCols] floats
    static constexpr int kBarOffset = kColsumOffset + 4 * kCsCols * 4;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16 + 1024;
};

struct PersistSched {
    int m_tiles;          // 128-row tiles per class
    int n_tiles;          // N / BN
    int classes;
    int super_per_class;  // ceil(m_tiles / (2*MT))
    int total_units;      // super_per_class * n_tiles * classes
    int num_pairs;        // gridDim.x / 2
};

template <int BN, int MT, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tap_gemm_persist_kernel(const __grid_constant__ TapGemmParams p, const __grid_constant__ PersistSched sc) {
    using L = SmemLayoutP<BN, MT, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* epi_buf = reinterpret_cast<float*>(smem + L::kEpiOffset);
    // fused column sums: one private [kCsCols] slice per epilogue warp (plain read-modify-write, no atomics)
    float* cs_smem = (p.colsum && p.n_total <= L::kCsCols) ? reinterpret_cast<float*>(smem + L::kColsumOffset) : nullptr;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] (used on the leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = blockIdx.x >> 1;
    const int num_kb = p.ntaps * p.cblocks;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.tmap_a);
        tc::prefetch_tmap(&p.tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tmem_full_bar[b], 1);
            tc::mbar_init(&tmem_empty_bar[b], 8);       // 4 epilogue warps x 2 CTAs
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc2(tmem_slot, 2 * MT * BN);
    if (cs_smem)
        for (int i = threadIdx.x; i < 4 * L::kCsCols; i += kThreads) cs_smem[i] = 0.f;
    tc::fence_before_sync();
    tc::cluster_sync_all();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // work unit -> (class, n tile, first m tile of this CTA)
    auto decode = [&](int unit, int& cls, int& n0, int& mt0) {
        const int nt = unit % sc.n_tiles;
        const int rest = unit / sc.n_tiles;
        const int sup = rest % sc.super_per_class;
        cls = rest / sc.super_per_class;
        n0 = nt * BN;
        mt0 = (sup * 2 + (int)rank) * MT;
    };
    auto tile_base = [&](int mtile, int (&base)[4]) {
        int t = mtile;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            base[d] = (t % p.tiles[d]) * p.box[d];
            t /= p.tiles[d];
        }
        if (mtile >= sc.m_tiles) base[3] = p.extent[3] + p.box[3] * (1 + t);   // surplus tile: OOB loads, masked stores
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = pair_id; unit < sc.total_units; unit += sc.num_pairs) {
                int cls, n0, mt0;
                decode(unit, cls, n0, mt0);
                int base[MT][4];
#pragma unroll
                for (int m = 0; m < MT; ++m) tile_base(mt0 + m, base[m]);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int t = kb / p.cblocks;
                    const int cb = kb - t * p.cblocks;
                    tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kATile;
                    const uint32_t bar = tc::map_to_cta(tc::smem_u32(&full_bar[stage]), 0);
                    if (leader) tc::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    const int* tp = p.tap[cls][t];
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tc::tma2_load_5d(sa + m * kATileBytes, &p.tmap_a, bar, cb * kBlockK + tp[0], base[m][0] + tp[1],
                                         base[m][1] + tp[2], base[m][2] + tp[3], base[m][3] + tp[4]);
                    tc::tma2_load_2d(sb, &p.tmap_b, bar, kb * kBlockK, cls * p.b_rows_per_cls + n0 + (int)rank * (BN / 2));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && leader) {
        constexpr uint32_t idesc = tc::idesc_tf32(256, BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int unit = pair_id; unit < sc.total_units; unit += sc.num_pairs, ++it) {
            const int ab = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            tc::mbar_wait_cluster(&tmem_empty_bar[ab], (use & 1) ^ 1);     // epilogues of both CTAs drained this set
            tc::fence_after_sync();
            for (int kb = 0; kb < num_kb; ++kb) {
                tc::mbar_wait(&full_bar[stage], phase);
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    const uint32_t sa = tc::smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t sb = sa + L::kATile;
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 8; ++k) {
                            const uint64_t adesc = tc::smem_desc_sw128(sa + m * kATileBytes + k * 32, 16, 1024);
                            const uint64_t bdesc = tc::smem_desc_sw128(sb + k * 32, 16, 1024);
                            tc::mma2_tf32(tmem_base + (uint32_t)((ab * MT + m) * BN), adesc, bdesc, idesc, (kb | k) ? 1u : 0u);
                        }
                    }
                    tc::mma2_commit_multicast(&empty_bar[stage]);
                    if (kb == num_kb - 1) tc::mma2_commit_multicast(&tmem_full_bar[ab]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        int it = 0;
        for (int unit = pair_id; unit < sc.total_units; unit += sc.num_pairs, ++it) {
            const int ab = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            int cls, n0, mt0;
            decode(unit, cls, n0, mt0);
            tc::mbar_wait(&tmem_full_bar[ab], use & 1);
            tc::fence_after_sync();
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                int base[4];
                tile_base(mt0 + m, base);
                epilogue_rows_nowait<BN>(p, base, cls, n0, tmem_base + (uint32_t)((ab * MT + m) * BN), warp, lane, epi_buf,
                                         cs_smem ? cs_smem + (warp & 3) * L::kCsCols : nullptr);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(&tmem_empty_bar[ab], 0);
        }
        if (cs_smem) {          // flush this CTA's partial column sums: one global atomic per column
            asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps
            for (int i = (int)threadIdx.x - 128; i < p.n_total; i += 128) {
                const float v = (cs_smem[i] + cs_smem[L::kCsCols + i]) + (cs_smem[2 * L::kCsCols + i] + cs_smem[3 * L::kCsCols + i]);
                if (v != 0.f) atomicAdd(p.colsum + i, v);
            }
        }
    }
    tc::fence_before_sync();
    tc::cluster_sync_all();
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc2(tmem_base, 2 * MT * BN);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// rank-5 fp32 tensor map with 128B swizzle; dims[0] is the contiguous one.
int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box) {
    auto fn = get_encode_fn();
    if (!fn) {
        cb200_set_error("cuTensorMapEncodeTiled driver entry point not available");
        return CB200_ERR_TMAP;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i];
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cb200_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu %llu %llu)",
                        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                        (unsigned long long)(rank > 4 ? dims[4] : 0));
        return CB200_ERR_TMAP;
    }
    return CB200_OK;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// split 128 output rows over (w, h, b): w fastest
void pick_spatial_box(int Wo, int Ho, int* wt, int* ht, int* bt) {
    *wt = Wo < kBlockM ? Wo : kBlockM;
    int rest = kBlockM / *wt;
    *ht = Ho < rest ? Ho : rest;
    *bt = rest / *ht;
}

// One-shot split-K workspace handed over by the caller right before a tap-GEMM call (cb200_tapgemm_workspace):
// thread-local, consumed by the next dispatch() of this thread.
struct SplitWs {
    float* ws = nullptr;
    long long bytes = 0;
    int* counters = nullptr;
    int n_counters = 0;
};
thread_local SplitWs g_split_ws;

// 0 = never split K, 1 = split when the tile list cannot fill the GPU (default); env CB200_TAPGEMM_SPLITK overrides.
int splitk_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CB200_TAPGEMM_SPLITK");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

// CTA-pair tiles + split-K for small tile lists: OFF by default - measured 1.5-2x SLOWER than the single-CTA kernel on
// every small shape (tools/bench_small_gemm.py: 4x4x512 at 64 images 54 -> 103 us); env CB200_TAPGEMM_SMALL_PAIR=1 enables
// it for A/B runs.
int small_pair_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CB200_TAPGEMM_SMALL_PAIR");
        mode = e ? atoi(e) : 0;
    }
    return mode;
}

// 1 (default) = deep shared-memory ring for grids of at most one CTA per SM; env CB200_TAPGEMM_DEEP=0 disables.
int deep_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CB200_TAPGEMM_DEEP");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

int plan_splits(const SplitWs& w, int units, int num_kb, int bn);

struct Epilogue {
    const float* bias;
    const float* dact;
    float slope;
    int round_out;
    float* colsum = nullptr;
};

// Split-K plan for a tile list that cannot fill the GPU: one CTA per SM (the deep-pipeline variant below), every split
// keeps >= 8 k-blocks.  Returns the number of splits (>= 1).  Measured on the per-rank shapes of the 8-GPU run
// (tools/bench_small_gemm.py, profiles/small_gemm_r2.md): splitting pays only for SHORT tile lists with a LONG reduction
// (heads forward 8192 -> 1536 at 64 / 192 rows: 85 -> 54 us; 4x4x512 layers at 64 images: 54 -> 42 us); with more tiles or
// a shorter K the partial-tile traffic and the serial finisher cost more than the extra SMs bring.
int plan_splits(const SplitWs& w, int units, int num_kb, int bn) {
    int splits = 1;
    if (w.ws && w.counters && splitk_mode() && units <= 48 && num_kb >= 96 && units <= w.n_counters) {
        splits = 148 / units;
        if (splits > num_kb / 8) splits = num_kb / 8;
        if (splits > 16) splits = 16;
        const long long per_split = (long long)units * kBlockM * bn * 4;
        if ((long long)splits * per_split > w.bytes) splits = (int)(w.bytes / per_split);
        if (splits < 1) splits = 1;
    }
    return splits;
}

template <int BN, int STAGES>
int launch_cfg(TapGemmParams& p, int m_tiles, int n_tiles, int classes, cudaStream_t st, const char* name) {
    using L = SmemLayout<BN, STAGES>;
    static bool configured_dev[64] = {};          // per device: the attribute belongs to the device's context
    bool& configured = cb200_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tap_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             L::kTotal);
        if (e != cudaSuccess) {
            cb200_set_error("%s: cudaFuncSetAttribute(smem=%d): %s", name, L::kTotal, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    dim3 grid(m_tiles, n_tiles, classes * p.splits);
    tap_gemm_kernel<BN, STAGES><<<grid, kThreads, L::kTotal, st>>>(p);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH(name);
    return CB200_OK;
}

// STAGES = ring depth when the grid is large (2 CTAs per SM), DEEP = ring depth when the whole grid is at most one CTA per
// SM: a lone CTA needs ~1 us of TMA latency covered by loads in flight (measured: 3 x 32 KB stages sustain one k-block
// per 0.36 us against 0.19 us of MMA time), so the small-grid variant spends the SM's whole shared memory on the ring.
template <int BN, int STAGES, int DEEP>
int launch(const TapGemmParams& p_in, int m_tiles, int n_tiles, int classes, cudaStream_t st, const char* name,
           const SplitWs& w) {
    TapGemmParams p = p_in;
    const int units = m_tiles * n_tiles * classes;
    const int num_kb = p.ntaps * p.cblocks;
    const int splits = plan_splits(w, units, num_kb, BN);
    p.kb_per_split = (num_kb + splits - 1) / splits;
    p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.classes = classes;
    p.ws = w.ws;
    p.counters = w.counters;
    const bool deep = units * p.splits <= 148 && deep_mode();
    static const bool verbose = getenv("CB200_TAPGEMM_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "[tapgemm] %s BN=%d m_tiles=%d n_tiles=%d classes=%d k-blocks=%d -> splits=%d (%d CTAs, %d stages)\n",
                name, BN, m_tiles, n_tiles, classes, num_kb, p.splits, units * p.splits, deep ? DEEP : STAGES);
    if (deep) return launch_cfg<BN, DEEP>(p, m_tiles, n_tiles, classes, st, name);
    return launch_cfg<BN, STAGES>(p, m_tiles, n_tiles, classes, st, name);
}

template <int BN, int STAGES>
int launch2(const TapGemmParams& p_in, int m_tiles, int n_tiles, int classes, cudaStream_t st, const char* name,
            const SplitWs& w = SplitWs()) {
    using L = SmemLayout2<BN, STAGES>;
    static bool configured_dev[64] = {};          // per device: the attribute belongs to the device's context
    bool& configured = cb200_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tap_gemm2_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             L::kTotal);
        if (e != cudaSuccess) {
            cb200_set_error("%s: cudaFuncSetAttribute(smem=%d): %s", name, L::kTotal, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    TapGemmParams p = p_in;
    const int m_pad = (m_tiles + 1) & ~1;
    const int units = m_pad * n_tiles * classes;             // CTAs per split (partials / counters are per CTA)
    const int num_kb = p.ntaps * p.cblocks;
    const int splits = plan_splits(w, units, num_kb, BN);
    p.kb_per_split = (num_kb + splits - 1) / splits;
    p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.classes = classes;
    p.ws = w.ws;
    p.counters = w.counters;
    static const bool verbose = getenv("CB200_TAPGEMM_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "[tapgemm] %s pair BN=%d m_tiles=%d n_tiles=%d classes=%d k-blocks=%d -> splits=%d (%d CTAs)\n", name,
                BN, m_tiles, n_tiles, classes, num_kb, p.splits, units * p.splits);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(m_pad, n_tiles, classes * p.splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tap_gemm2_kernel<BN, STAGES>, p);
    CB200_COUNT_LAUNCH();
    if (e != cudaSuccess) {
        cb200_set_error("%s: cluster launch failed: %s", name, cudaGetErrorString(e));
        return (int)e;
    }
    CB200_CHECK_LAUNCH(name);
    return CB200_OK;
}

template <int BN, int MT, int STAGES>
int launch_persist(const TapGemmParams& p, int m_tiles, int n_tiles, int classes, cudaStream_t st, const char* name) {
    using L = SmemLayoutP<BN, MT, STAGES>;
    static bool configured_dev[64] = {};          // per device: the attribute belongs to the device's context
    bool& configured = cb200_device_flag(configured_dev);
    static int sms = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tap_gemm_persist_kernel<BN, MT, STAGES>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) {
            cb200_set_error("%s: cudaFuncSetAttribute(smem=%d): %s", name, L::kTotal, cudaGetErrorString(e));
            return (int)e;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        configured = true;
    }
    PersistSched sc;
    sc.m_tiles = m_tiles; sc.n_tiles = n_tiles; sc.classes = classes;
    sc.super_per_class = (m_tiles + 2 * MT - 1) / (2 * MT);
    sc.total_units = sc.super_per_class * n_tiles * classes;
    sc.num_pairs = sms / 2 < sc.total_units ? sms / 2 : sc.total_units;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * sc.num_pairs, 1, 1);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tap_gemm_persist_kernel<BN, MT, STAGES>, p, sc);
    CB200_COUNT_LAUNCH();
    if (e != cudaSuccess) {
        cb200_set_error("%s: persistent cluster launch failed: %s", name, cudaGetErrorString(e));
        return (int)e;
    }
    CB200_CHECK_LAUNCH(name);
    return CB200_OK;
}

// 0 = off, 1 = persistent CTA-pair kernels where the shape allows (env CB200_TAPGEMM_PERSIST)
int persist_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CB200_TAPGEMM_PERSIST");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

// 0 = single-CTA tiles only, 1 = CTA pairs where the shape allows (default); env CB200_TAPGEMM_PAIR overrides.
int pair_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("CB200_TAPGEMM_PAIR");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

// Rows of the weight matrix one CTA stages per k-block for an (N, m_tiles) problem - the B tensor map's box must
// be encoded with exactly this many rows.  Mirrors the kernel choice made in dispatch().
int b_box_rows(int N, int m_tiles) {
    if (persist_mode() && m_tiles >= 64 && N % 64 == 0) return (N % 256 == 0) ? 128 : (N % 128 == 0 ? 64 : 32);
    if (persist_mode() && m_tiles >= 64 && N == 32) return 16;
    if (pair_mode() && m_tiles >= 296) {
        if (N % 256 == 0) return 128;      // pair, BN = 256: half per CTA
        if (N % 128 == 0) return 64;       // pair, BN = 128
    }
    if (small_pair_mode() && m_tiles >= 2 && g_split_ws.ws) {        // the split-K CTA-pair path of dispatch()
        if (N % 256 == 0) return 128;
        if (N % 128 == 0 && m_tiles >= 4) return 64;
    }
    return (N % 128 == 0) ? 128 : (N % 64 == 0 ? 64 : 32);
}

int debug_flags() {
    static int flags = -1;
    if (flags < 0) {
        const char* e = getenv("CB200_TAPGEMM_DEBUG");
        flags = e ? atoi(e) : 0;
    }
    return flags;
}

int dispatch(const TapGemmParams& p_in, int N, int m_tiles, int classes, cudaStream_t st, const char* name) {
    TapGemmParams p = p_in;
    const SplitWs w = g_split_ws;
    g_split_ws = SplitWs();                 // one-shot
    p.debug = debug_flags();
    p.n_total = N;
    p.splits = 1; p.kb_per_split = p.ntaps * p.cblocks; p.classes = classes;
    if (p.colsum) {
        cudaError_t e = cudaMemsetAsync(p.colsum, 0, sizeof(float) * (size_t)N, st);
        if (e != cudaSuccess) { cb200_set_error("%s: colsum memset: %s", name, cudaGetErrorString(e)); return (int)e; }
    }
    if (persist_mode() && m_tiles >= 64 && N % 64 == 0) {
        if (N % 256 == 0) return launch_persist<256, 1, 5>(p, m_tiles, N / 256, classes, st, name);   // 5 x 32 KB stages
        if (N % 128 == 0) return launch_persist<128, 2, 4>(p, m_tiles, N / 128, classes, st, name);   // 4 x 40 KB stages
        return launch_persist<64, 4, 3>(p, m_tiles, N / 64, classes, st, name);                        // 3 x 68 KB stages
    }
    if (persist_mode() && m_tiles >= 64 && N == 32) {     // thin outputs (3 image channels padded to 32): A-supply bound
        return launch_persist<32, 4, 3>(p, m_tiles, 1, classes, st, name);                             // 3 x 66 KB stages
    }
    if (pair_mode() && m_tiles >= 296) {
        if (N % 256 == 0) return launch2<256, 4>(p, m_tiles, N / 256, classes, st, name);     // 32 KB / stage / CTA
        if (N % 128 == 0) return launch2<128, 4>(p, m_tiles, N / 128, classes, st, name);     // 24 KB / stage / CTA
    }
    // small tile lists (deep layers / heads at a small per-GPU batch): these GEMMs are bound by L2 -> SM operand traffic
    // (measured: the 128 x 128 tile tops out at 260-380 TFLOP/s however many CTAs run), so take the 256 x 256 CTA-pair
    // tile (half the operand bytes per FLOP) and fill the GPU by splitting K
    if (small_pair_mode() && m_tiles >= 2 && w.ws) {
        if (N % 256 == 0) return launch2<256, 4>(p, m_tiles, N / 256, classes, st, name, w);
        if (N % 128 == 0 && m_tiles >= 4) return launch2<128, 4>(p, m_tiles, N / 128, classes, st, name, w);
    }
    if (N % 128 == 0) return launch<128, 3, 6>(p, m_tiles, N / 128, classes, st, name, w);   // 96 KB (2 CTAs / SM) | 192 KB
    if (N % 64 == 0) return launch<64, 4, 8>(p, m_tiles, N / 64, classes, st, name, w);      // 96 KB | 192 KB
    if (N % 32 == 0) return launch<32, 4, 8>(p, m_tiles, N / 32, classes, st, name, w);      // 80 KB | 160 KB
    cb200_set_error("%s: N=%d must be a multiple of 32", name, N);
    return CB200_ERR_ARG;
}

int check_ptr16(const void* p, const char* what) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) {
        cb200_set_error("%s must be 16-byte aligned", what);
        return CB200_ERR_ARG;
    }
    return CB200_OK;
}

// Common set-up for an NHWC tensor [Bn, Hs, Ws, C] traversed with unit spatial stride.
int conv_like(const float* src, int Bn, int Hs, int Ws, int C,           // A tensor (NHWC)
              int Wo, int Ho,                                             // positions per image along w / h
              const float* wmat, int N, int ntaps, const int (*taps)[2],  // weight [N(*classes), ntaps*C]; taps = {dh, dw}
              int classes, const int (*cls_taps)[4][2],                   // optional per-class taps (classes == 4)
              float* out, int ldo, long long os_w, long long os_h, long long os_b, const long long* cls_off,
              const Epilogue& ep, cudaStream_t st, const char* name) {
    CB200_CHECK_ARG(C % kBlockK == 0, "%s: channel count %d must be a multiple of 32", name, C);
    CB200_CHECK_ARG(is_pow2(Wo) && is_pow2(Ho), "%s: spatial size %dx%d must be powers of two", name, Ho, Wo);
    if (int e = check_ptr16(src, "activation pointer")) return e;
    if (int e = check_ptr16(wmat, "weight pointer")) return e;
    if (int e = check_ptr16(out, "output pointer")) return e;
    TapGemmParams p;
    memset(&p, 0, sizeof(p));
    int wt, ht, bt;
    pick_spatial_box(Wo, Ho, &wt, &ht, &bt);
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)Bn, 1};
    uint64_t strides[5] = {4, (uint64_t)C * 4, (uint64_t)Ws * C * 4, (uint64_t)Hs * Ws * C * 4,
                           (uint64_t)Bn * Hs * Ws * C * 4};
    uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)wt, (uint32_t)ht, (uint32_t)bt, 1};
    if (int e = encode_map(&p.tmap_a, src, 5, dims, strides, box)) return e;
    p.box[0] = wt; p.box[1] = ht; p.box[2] = bt; p.box[3] = 1;
    p.tiles[0] = (Wo + wt - 1) / wt; p.tiles[1] = (Ho + ht - 1) / ht; p.tiles[2] = (Bn + bt - 1) / bt; p.tiles[3] = 1;
    {
        const int Ktot = ntaps * C;
        uint64_t bdims[2] = {(uint64_t)Ktot, (uint64_t)N * classes};
        uint64_t bstr[2] = {4, (uint64_t)Ktot * 4};
        uint32_t bbox[2] = {(uint32_t)kBlockK, (uint32_t)b_box_rows(N, p.tiles[0] * p.tiles[1] * p.tiles[2])};
        if (int e = encode_map(&p.tmap_b, wmat, 2, bdims, bstr, bbox)) return e;
    }
    p.extent[0] = Wo; p.extent[1] = Ho; p.extent[2] = Bn; p.extent[3] = 1;
    p.ostride[0] = os_w; p.ostride[1] = os_h; p.ostride[2] = os_b; p.ostride[3] = 0;
    for (int c = 0; c < classes; ++c) {
        p.cls_off[c] = cls_off ? cls_off[c] : 0;
        for (int t = 0; t < ntaps; ++t) {
            const int* hw = cls_taps ? cls_taps[c][t] : taps[t];
            p.tap[c][t][0] = 0;
            p.tap[c][t][1] = hw[1];   // dw
            p.tap[c][t][2] = hw[0];   // dh
            p.tap[c][t][3] = 0;
            p.tap[c][t][4] = 0;
        }
    }
    p.ntaps = ntaps; p.cblocks = C / kBlockK;
    p.ldo = ldo; p.b_rows_per_cls = N;
    p.out = out; p.bias = ep.bias; p.dact = ep.dact; p.slope = ep.slope; p.round_out = ep.round_out;
    p.colsum = ep.colsum;
    const int m_tiles = p.tiles[0] * p.tiles[1] * p.tiles[2];
    return dispatch(p, N, m_tiles, classes, st, name);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------

// Optional split-K workspace for the NEXT cb200_gemm_nt_tf32 / cb200_conv2d_nhwc_fwd / cb200_conv2d_nhwc_dgrad call made
// by this host thread (one-shot).  ws: device scratch of `bytes` bytes (any contents); counters: `n_counters` int32 that
// are ZERO and stay zero between launches (the kernel re-arms them) - one persistent buffer per (device, stream).
// Without it (or when the tile list already fills the GPU) the kernels run unsplit.  The caller keeps `ws` alive until
// the launch has been enqueued on its stream (stream-ordered reuse afterwards is fine).
extern "C" int cb200_tapgemm_workspace(void* ws, long long bytes, int* counters, int n_counters) {
    g_split_ws.ws = static_cast<float*>(ws);
    g_split_ws.bytes = ws ? bytes : 0;
    g_split_ws.counters = counters;
    g_split_ws.n_counters = counters ? n_counters : 0;
    return CB200_OK;
}

// out[M, N] (row stride ldo) = epi( A[M, K] (row stride lda) * Bw[N, K]^T (row stride ldb) + bias ), TF32 tensor cores.
// epi = LeakyReLU(slope) when dact == NULL, else multiply by lrelu'(dact[M,N]) (dact shares ldo with out).
extern "C" int cb200_gemm_nt_tf32(const float* a, long long lda, const float* bw, long long ldb, const float* bias,
                                  const float* dact, float* out, long long ldo, int M, int N, int K, float slope,
                                  int round_out, float* colsum, void* stream) {
    CB200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_nt: empty problem");
    CB200_CHECK_ARG(K % kBlockK == 0, "gemm_nt: K=%d must be a multiple of 32", K);
    CB200_CHECK_ARG(lda % 4 == 0 && ldo % 4 == 0 && ldb % 4 == 0, "gemm_nt: lda/ldb/ldo must be multiples of 4 floats");
    if (int e = check_ptr16(a, "gemm_nt: A")) return e;
    if (int e = check_ptr16(bw, "gemm_nt: B")) return e;
    if (int e = check_ptr16(out, "gemm_nt: out")) return e;
    TapGemmParams p;
    memset(&p, 0, sizeof(p));
    uint64_t dims[5] = {(uint64_t)K, (uint64_t)M, 1, 1, 1};
    uint64_t strides[5] = {4, (uint64_t)lda * 4, (uint64_t)lda * 4 * M, (uint64_t)lda * 4 * M, (uint64_t)lda * 4 * M};
    uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)kBlockM, 1, 1, 1};
    if (int e = encode_map(&p.tmap_a, a, 5, dims, strides, box)) return e;
    uint64_t bdims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t bstr[2] = {4, (uint64_t)ldb * 4};
    uint32_t bbox[2] = {(uint32_t)kBlockK, (uint32_t)b_box_rows(N, (M + kBlockM - 1) / kBlockM)};
    if (int e = encode_map(&p.tmap_b, bw, 2, bdims, bstr, bbox)) return e;
    p.box[0] = kBlockM; p.box[1] = 1; p.box[2] = 1; p.box[3] = 1;
    p.tiles[0] = (M + kBlockM - 1) / kBlockM; p.tiles[1] = 1; p.tiles[2] = 1; p.tiles[3] = 1;
    p.extent[0] = M; p.extent[1] = 1; p.extent[2] = 1; p.extent[3] = 1;
    p.ostride[0] = 1;
    p.ntaps = 1; p.cblocks = K / kBlockK;
    p.ldo = (int)ldo; p.b_rows_per_cls = 0;
    p.out = out; p.bias = bias; p.dact = dact; p.slope = slope; p.round_out = round_out;
    p.colsum = colsum;
    return dispatch(p, N, p.tiles[0], 1, static_cast<cudaStream_t>(stream), "gemm_nt_tf32");
}

// y[B,Ho,Wo,Cout] = lrelu(conv(x[B,H,W,Cin], w) + bias), NHWC, w as GEMM matrix [Cout, ks*ks*Cin] (tap-major).
// Supported: (ks=3, stride=1, pad=1) and (ks=4, stride=2, pad=1).
extern "C" int cb200_conv2d_nhwc_fwd(const float* x, const float* wmat, const float* bias, float* y, int B, int H,
                                     int W, int Cin, int Cout, int ks, int stride, float slope, int round_out,
                                     void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Epilogue ep{bias, nullptr, slope, round_out};
    if (ks == 3 && stride == 1) {
        int taps[9][2];
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) { taps[kh * 3 + kw][0] = kh - 1; taps[kh * 3 + kw][1] = kw - 1; }
        return conv_like(x, B, H, W, Cin, W, H, wmat, Cout, 9, taps, 1, nullptr, y, Cout, 1, W, (long long)H * W,
                         nullptr, ep, st, "conv3x3s1_fwd");
    }
    if (ks == 4 && stride == 2) {
        // space-to-depth view of x: dims (2C, W/2, 2, H/2, B); tap (kh,kw) -> (dh,ph),(dw,pw)
        CB200_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "conv4x4s2_fwd: H, W must be even");
        CB200_CHECK_ARG(Cin % kBlockK == 0, "conv4x4s2_fwd: Cin=%d must be a multiple of 32", Cin);
        const int Ho = H / 2, Wo = W / 2;
        CB200_CHECK_ARG(is_pow2(Wo) && is_pow2(Ho), "conv4x4s2_fwd: output size must be powers of two");
        if (int e = check_ptr16(x, "conv4x4s2_fwd: x")) return e;
        if (int e = check_ptr16(wmat, "conv4x4s2_fwd: w")) return e;
        if (int e = check_ptr16(y, "conv4x4s2_fwd: y")) return e;
        TapGemmParams p;
        memset(&p, 0, sizeof(p));
        int wt, ht, bt;
        pick_spatial_box(Wo, Ho, &wt, &ht, &bt);
        uint64_t dims[5] = {(uint64_t)2 * Cin, (uint64_t)Wo, 2, (uint64_t)Ho, (uint64_t)B};
        uint64_t strides[5] = {4, (uint64_t)2 * Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)2 * W * Cin * 4,
                               (uint64_t)H * W * Cin * 4};
        uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)wt, 1, (uint32_t)ht, (uint32_t)bt};
        if (int e = encode_map(&p.tmap_a, x, 5, dims, strides, box)) return e;
        const int Ktot = 16 * Cin;
        uint64_t bdims[2] = {(uint64_t)Ktot, (uint64_t)Cout};
        uint64_t bstr[2] = {4, (uint64_t)Ktot * 4};
        const int mt = ((Wo + wt - 1) / wt) * ((Ho + ht - 1) / ht) * ((B + bt - 1) / bt);
        uint32_t bbox[2] = {(uint32_t)kBlockK, (uint32_t)b_box_rows(Cout, mt)};
        if (int e = encode_map(&p.tmap_b, wmat, 2, bdims, bstr, bbox)) return e;
        p.box[0] = wt; p.box[1] = 1; p.box[2] = ht; p.box[3] = bt;
        p.tiles[0] = (Wo + wt - 1) / wt; p.tiles[1] = 1; p.tiles[2] = (Ho + ht - 1) / ht; p.tiles[3] = (B + bt - 1) / bt;
        p.extent[0] = Wo; p.extent[1] = 1; p.extent[2] = Ho; p.extent[3] = B;
        p.ostride[0] = 1; p.ostride[1] = 0; p.ostride[2] = Wo; p.ostride[3] = (long long)Ho * Wo;
        static const int dpar[4][2] = {{-1, 1}, {0, 0}, {0, 1}, {1, 0}};   // k -> (delta, parity)
        for (int kh = 0; kh < 4; ++kh)
            for (int kw = 0; kw < 4; ++kw) {
                int* tp = p.tap[0][kh * 4 + kw];
                tp[0] = dpar[kw][1] * Cin;   // channel offset selects the w-parity half
                tp[1] = dpar[kw][0];         // dw (in half-res units)
                tp[2] = dpar[kh][1];         // h parity
                tp[3] = dpar[kh][0];         // dh
                tp[4] = 0;
            }
        p.ntaps = 16; p.cblocks = Cin / kBlockK;
        p.ldo = Cout; p.b_rows_per_cls = 0;
        p.out = y; p.bias = bias; p.dact = nullptr; p.slope = slope; p.round_out = round_out;
        return dispatch(p, Cout, p.tiles[0] * p.tiles[2] * p.tiles[3], 1, st, "conv4x4s2_fwd");
    }
    cb200_set_error("conv2d_nhwc_fwd: unsupported kernel/stride %d/%d", ks, stride);
    return CB200_ERR_ARG;
}

// dx[B,H,W,Cin] = dgrad(dy[B,Ho,Wo,Cout], w) (* lrelu'(act_in) if act_in != null) (+ bias_out, lrelu if given).
//   ks=3,stride=1: wmat_t = [Cin, 3*3*Cout] with tap index (kh*3+kw) holding W[co,ci,kh,kw]
//   ks=4,stride=2: wmat_t = [4 classes (ph*2+pw)][Cin][4 taps][Cout]  (see pack_weights in weights.cu)
// This is also ConvTranspose2d(4,2,1) / ConvTranspose2d(3,1,1) *forward* (G_SNDCGAN) with bias_out / slope.
extern "C" int cb200_conv2d_nhwc_dgrad(const float* dy, const float* wmat_t, const float* act_in, const float* bias_out,
                                       float* dx, int B, int H, int W, int Cin, int Cout, int ks, int stride,
                                       float slope, int round_out, float* colsum, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Epilogue ep{bias_out, act_in, slope, round_out, colsum};
    if (ks == 3 && stride == 1) {
        int taps[9][2];
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) { taps[kh * 3 + kw][0] = 1 - kh; taps[kh * 3 + kw][1] = 1 - kw; }
        return conv_like(dy, B, H, W, Cout, W, H, wmat_t, Cin, 9, taps, 1, nullptr, dx, Cin, 1, W, (long long)H * W,
                         nullptr, ep, st, "conv3x3s1_dgrad");
    }
    if (ks == 4 && stride == 2) {
        CB200_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "conv4x4s2_dgrad: H, W must be even");
        const int Ho = H / 2, Wo = W / 2;
        // output parity p: taps (k, delta on the dy grid): p=0 -> (1,0),(3,-1);  p=1 -> (0,+1),(2,0)
        static const int dsel[2][2] = {{0, -1}, {1, 0}};   // [parity][j] -> delta
        int cls_taps[4][4][2];
        long long cls_off[4];
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                int c = ph * 2 + pw;
                cls_off[c] = (long long)ph * W + pw;
                for (int jh = 0; jh < 2; ++jh)
                    for (int jw = 0; jw < 2; ++jw) {
                        cls_taps[c][jh * 2 + jw][0] = dsel[ph][jh];
                        cls_taps[c][jh * 2 + jw][1] = dsel[pw][jw];
                    }
            }
        return conv_like(dy, B, Ho, Wo, Cout, Wo, Ho, wmat_t, Cin, 4, nullptr, 4, cls_taps, dx, Cin, 2, 2LL * W,
                         (long long)H * W, cls_off, ep, st, "conv4x4s2_dgrad");
    }
    cb200_set_error("conv2d_nhwc_dgrad: unsupported kernel/stride %d/%d", ks, stride);
    return CB200_ERR_ARG;
}
