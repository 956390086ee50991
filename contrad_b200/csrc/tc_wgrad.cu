// tcgen05 / TMA weight-gradient GEMM for sm_100a (the third leg of every conv / linear layer):
//
//     dW_hat[co, t*Cin + ci] = sum_{pixels p} dY[p, co] * X[p shifted by tap t, ci]
//
// i.e. D[M'=Cout, N'=Cin] = A'^T B' with the REDUCTION over pixels, which is the slow (row) dimension
// of both NHWC operands.  Both operands are therefore fed to tcgen05.mma as MN-major shared-memory
// tiles (descriptor major bits = 1): a TMA box of [KP pixels x 32 channels] lands as KP rows of one
// 128-byte span.  For 32-bit MN-major operands the tensor core only accepts the SWIZZLE_128B_BASE32B
// layout (32-byte swizzle atoms), which TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B -
// so no transpose ever happens in memory.  The tap shift / zero padding is again a TMA coordinate
// offset + out-of-bounds fill.  Pixels are split across CTAs (split-K); partial tiles are reduced
// with fp32 atomics (red.global.add) into a zeroed dW_hat.
//
// Replaces the cuDNN wgrad / cuBLAS calls of autograd for nn.Conv2d / nn.Linear in the reference
// (models/gan/sndcgan.py:91-109, models/gan/base.py:14-35,92-101).
// Roofline: tensor pipe; algorithmic FLOPs = 2 * pixels * Cout * taps * Cin per launch.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cudaTypedefs.h>
#include <string.h>

namespace {

constexpr int kTileM = 128;     // Cout rows per CTA
constexpr int kThreads = 256;
constexpr int kMaxTaps = 16;

struct WgradParams {
    CUtensorMap tmap_dy;      // (Cout, d1..d4)
    CUtensorMap tmap_x;       // (C,    d1..d4)
    int box[4];               // pixel box along dims 1..4 (product == KP)
    int tiles[4];             // pixel-tile counts along dims 1..4
    int tap[kMaxTaps][5];     // {c_add, d1, d2, d3, d4} offsets applied to the X box
    int ntaps;
    int cin_tiles;            // Cin / BN
    int total_btiles;         // ntaps * cin_tiles  ("B tiles": one (tap, Cin tile) each)
    int total_ptiles;         // product of tiles[]
    int ptiles_per_split;
    int ldw;                  // floats per dW_hat row (= ntaps * Cin for convs)
    float* dw;
};

// One pipeline stage = the dY box (128 channels x KP pixels) + NB X boxes (BN channels x KP pixels) that SHARE it:
// NB (tap, Cin-tile) products accumulate into NB*BN TMEM columns, so the dY bytes are amortised over NB MMAs.
template <int BN, int NB, int KP, int STAGES>
struct WgSmem {
    static constexpr int kChunkBytes = KP * 128;                      // one [KP x 32 fp32] box
    static constexpr int kABytes = (kTileM / 32) * kChunkBytes;
    static constexpr int kBBytes = (BN / 32) * kChunkBytes;
    static constexpr int kStageBytes = kABytes + NB * kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;
    static constexpr int kTmemCols = (NB * BN <= 32) ? 32 : (NB * BN <= 64) ? 64 : (NB * BN <= 128) ? 128 : (NB * BN <= 256) ? 256 : 512;
};

template <int BN, int NB, int KP, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ WgradParams p) {
    using L = WgSmem<BN, NB, KP, STAGES>;
    static_assert(NB * BN <= 512, "accumulators exceed TMEM");
    static_assert(L::kStageBytes >= 4 * 32 * 36 * 4, "stage 0 doubles as the epilogue staging buffer");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int bt0 = blockIdx.y * NB;                       // first B tile of this CTA
    const int nb = min(NB, p.total_btiles - bt0);          // valid B tiles (>= 1)
    const int co0 = blockIdx.z * kTileM;
    const int pt_begin = split * p.ptiles_per_split;
    const int pt_end = min(p.total_ptiles, pt_begin + p.ptiles_per_split);
    const int num_kb = pt_end - pt_begin;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.tmap_dy);
        tc::prefetch_tmap(&p.tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(tmem_full_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, L::kTmemCols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (num_kb > 0) {
        if (warp == 0) {
            if (tc::elect_one()) {
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    int t = pt_begin + kb;
                    int c[4];
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        c[d] = (t % p.tiles[d]) * p.box[d];
                        t /= p.tiles[d];
                    }
                    tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    tc::mbar_expect_tx(&full_bar[stage], L::kABytes + nb * L::kBBytes);
#pragma unroll
                    for (int j = 0; j < kTileM / 32; ++j)
                        tc::tma_load_5d(sa + j * L::kChunkBytes, &p.tmap_dy, &full_bar[stage], co0 + j * 32, c[0], c[1],
                                        c[2], c[3]);
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (b < nb) {
                            const int bt = bt0 + b;
                            const int tap_id = bt / p.cin_tiles;
                            const int ci0 = (bt - tap_id * p.cin_tiles) * BN;
                            const int* tp = p.tap[tap_id];
                            uint8_t* sb = sa + L::kABytes + b * L::kBBytes;
#pragma unroll
                            for (int j = 0; j < BN / 32; ++j)
                                tc::tma_load_5d(sb + j * L::kChunkBytes, &p.tmap_x, &full_bar[stage],
                                                ci0 + j * 32 + tp[0], c[0] + tp[1], c[1] + tp[2], c[2] + tp[3],
                                                c[3] + tp[4]);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            constexpr uint32_t idesc = tc::idesc_tf32(kTileM, BN, 1, 1);     // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                tc::mbar_wait(&full_bar[stage], phase);
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    const uint32_t sa = tc::smem_u32(smem + stage * L::kStageBytes);
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (b < nb) {
                            const uint32_t sb = sa + L::kABytes + b * L::kBBytes;
#pragma unroll
                            for (int g = 0; g < KP / 8; ++g) {
                                // MN-major TF32: LBO = stride between 32-element MN chunks, SBO = between 4-row K atoms
                                const uint64_t adesc = tc::smem_desc_sw128_base32(sa + g * 1024, L::kChunkBytes, 512);
                                const uint64_t bdesc = tc::smem_desc_sw128_base32(sb + g * 1024, L::kChunkBytes, 512);
                                tc::mma_tf32(tmem_base + b * BN, adesc, bdesc, idesc, (kb | g) ? 1u : 0u);
                            }
                        }
                    }
                    tc::mma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) tc::mma_commit(tmem_full_bar);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        } else if (warp >= 4) {
            // Each thread receives one accumulator ROW (co) from TMEM; the 32x32 block of a warp is transposed through
            // shared memory (stage-0 buffer, idle once tmem_full fired) so that 8 lanes issue one contiguous 128-byte
            // vector reduction (red.global.add.v4.f32) instead of 32 scattered scalar atomics per row.
            const int q = warp & 3;
            float* st = reinterpret_cast<float*>(smem) + q * (32 * 36);
            const int sub = lane >> 3, col4 = (lane & 7) * 4;
            tc::mbar_wait(tmem_full_bar, 0);
            tc::fence_after_sync();
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                const int bt = bt0 + b;
                const int tap_id = bt / p.cin_tiles;
                const int ci0 = (bt - tap_id * p.cin_tiles) * BN;
                float* dbase = p.dw + (long long)(co0 + q * 32) * p.ldw + (long long)tap_id * (p.cin_tiles * BN) + ci0;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + b * BN + c0, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(st + lane * 36 + j) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                        __uint_as_float(v[j + 3]));
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + sub;
                        const float4 o = *reinterpret_cast<const float4*>(st + row * 36 + col4);
                        atomicAdd(reinterpret_cast<float4*>(dbase + (long long)row * p.ldw + c0 + col4), o);
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, L::kTmemCols);
    }
}

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

int encode5(CUtensorMap* m, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
            const char* what) {
    auto fn = get_encode_fn();
    if (!fn) {
        cb200_set_error("cuTensorMapEncodeTiled driver entry point not available");
        return CB200_ERR_TMAP;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < 5; ++i) {
        gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i];
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cb200_set_error("wgrad: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
        return CB200_ERR_TMAP;
    }
    return CB200_OK;
}

// split KP pixels over (w, h, b), w fastest
void pick_pixel_box(int KP, int Wo, int Ho, int* wt, int* ht, int* bt) {
    *wt = Wo < KP ? Wo : KP;
    int rest = KP / *wt;
    *ht = Ho < rest ? Ho : rest;
    *bt = rest / *ht;
}

template <int BN, int NB, int KP, int STAGES>
int launch_wgrad(WgradParams& p, int Cout, int sm_count, cudaStream_t st, const char* name) {
    using L = WgSmem<BN, NB, KP, STAGES>;
    static bool configured_dev[64] = {};          // per device: the attribute belongs to the device's context
    bool& configured = cb200_device_flag(configured_dev);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BN, NB, KP, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             L::kTotal);
        if (e != cudaSuccess) {
            cb200_set_error("%s: cudaFuncSetAttribute(smem=%d): %s", name, L::kTotal, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    const int groups = (p.total_btiles + NB - 1) / NB;
    const int out_tiles = (Cout / kTileM) * groups;
    // split the pixel range so that the grid is (a little under) a whole number of waves of one CTA per SM
    int waves = (out_tiles + sm_count - 1) / sm_count;
    int splits = (waves * sm_count) / out_tiles;
    int max_splits = (p.total_ptiles + 7) / 8;          // at least 8 pipeline stages of work per CTA
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.ptiles_per_split = (p.total_ptiles + splits - 1) / splits;
    splits = (p.total_ptiles + p.ptiles_per_split - 1) / p.ptiles_per_split;
    dim3 grid(splits, groups, Cout / kTileM);
    wgrad_kernel<BN, NB, KP, STAGES><<<grid, kThreads, L::kTotal, st>>>(p);
    CB200_COUNT_LAUNCH();
    CB200_CHECK_LAUNCH(name);
    return CB200_OK;
}

int sm_count_cached() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// dw_hat[Cout, ks*ks*Cin] (forward-pack layout, column (kh*ks+kw)*Cin+ci) = wgrad(x[B,H,W,Cin], dy[B,Ho,Wo,Cout]).
// (ks,stride) in {(3,1),(4,2)}, pad 1.  The buffer is zeroed here and filled with split-K atomics.
extern "C" int cb200_conv2d_nhwc_wgrad(const float* x, const float* dy, float* dw_hat, int B, int H, int W, int Cin,
                                       int Cout, int ks, int stride, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CB200_CHECK_ARG((ks == 3 && stride == 1) || (ks == 4 && stride == 2), "conv2d_nhwc_wgrad: unsupported %d/%d", ks, stride);
    CB200_CHECK_ARG(Cout % kTileM == 0, "conv2d_nhwc_wgrad: Cout=%d must be a multiple of 128", Cout);
    CB200_CHECK_ARG(Cin % 32 == 0, "conv2d_nhwc_wgrad: Cin=%d must be a multiple of 32", Cin);
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) |
                      reinterpret_cast<uintptr_t>(dw_hat)) & 15) == 0, "conv2d_nhwc_wgrad: pointers must be 16-byte aligned");
    const int Ho = H / stride, Wo = W / stride;
    CB200_CHECK_ARG((Wo & (Wo - 1)) == 0 && (Ho & (Ho - 1)) == 0, "conv2d_nhwc_wgrad: output size must be powers of two");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    int wt, ht, bt;
    constexpr int KP = 32;                     // pixels per pipeline stage (4 MMAs of K = 8 per B tile)
    pick_pixel_box(KP, Wo, Ho, &wt, &ht, &bt);
    const int ntaps = ks * ks;
    {
        uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B, 1};
        uint64_t str[5] = {4, (uint64_t)Cout * 4, (uint64_t)Wo * Cout * 4, (uint64_t)Ho * Wo * Cout * 4,
                           (uint64_t)B * Ho * Wo * Cout * 4};
        uint32_t box[5] = {32, (uint32_t)wt, (uint32_t)ht, (uint32_t)bt, 1};
        if (stride == 2) {   // same pixel order as the space-to-depth X map below: (w, parity, h, b)
            uint64_t d2[5] = {(uint64_t)Cout, (uint64_t)Wo, 1, (uint64_t)Ho, (uint64_t)B};
            uint64_t s2[5] = {4, (uint64_t)Cout * 4, (uint64_t)Wo * Cout * 4, (uint64_t)Wo * Cout * 4,
                              (uint64_t)Ho * Wo * Cout * 4};
            uint32_t b2[5] = {32, (uint32_t)wt, 1, (uint32_t)ht, (uint32_t)bt};
            if (int e = encode5(&p.tmap_dy, dy, d2, s2, b2, "dy")) return e;
        } else {
            if (int e = encode5(&p.tmap_dy, dy, dims, str, box, "dy")) return e;
        }
    }
    if (stride == 1) {
        uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B, 1};
        uint64_t str[5] = {4, (uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4,
                           (uint64_t)B * H * W * Cin * 4};
        uint32_t box[5] = {32, (uint32_t)wt, (uint32_t)ht, (uint32_t)bt, 1};
        if (int e = encode5(&p.tmap_x, x, dims, str, box, "x")) return e;
        p.box[0] = wt; p.box[1] = ht; p.box[2] = bt; p.box[3] = 1;
        p.tiles[0] = (Wo + wt - 1) / wt; p.tiles[1] = (Ho + ht - 1) / ht; p.tiles[2] = (B + bt - 1) / bt; p.tiles[3] = 1;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                int* tp = p.tap[kh * 3 + kw];
                tp[0] = 0; tp[1] = kw - 1; tp[2] = kh - 1; tp[3] = 0; tp[4] = 0;
            }
    } else {
        uint64_t dims[5] = {(uint64_t)2 * Cin, (uint64_t)Wo, 2, (uint64_t)Ho, (uint64_t)B};
        uint64_t str[5] = {4, (uint64_t)2 * Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)2 * W * Cin * 4,
                           (uint64_t)H * W * Cin * 4};
        uint32_t box[5] = {32, (uint32_t)wt, 1, (uint32_t)ht, (uint32_t)bt};
        if (int e = encode5(&p.tmap_x, x, dims, str, box, "x")) return e;
        p.box[0] = wt; p.box[1] = 1; p.box[2] = ht; p.box[3] = bt;
        p.tiles[0] = (Wo + wt - 1) / wt; p.tiles[1] = 1; p.tiles[2] = (Ho + ht - 1) / ht; p.tiles[3] = (B + bt - 1) / bt;
        static const int dpar[4][2] = {{-1, 1}, {0, 0}, {0, 1}, {1, 0}};
        for (int kh = 0; kh < 4; ++kh)
            for (int kw = 0; kw < 4; ++kw) {
                int* tp = p.tap[kh * 4 + kw];
                tp[0] = dpar[kw][1] * Cin; tp[1] = dpar[kw][0]; tp[2] = dpar[kh][1]; tp[3] = dpar[kh][0]; tp[4] = 0;
            }
    }
    p.ntaps = ntaps;
    p.total_ptiles = p.tiles[0] * p.tiles[1] * p.tiles[2] * p.tiles[3];
    p.ldw = ntaps * Cin;
    p.dw = dw_hat;
    cudaError_t e = cudaMemsetAsync(dw_hat, 0, sizeof(float) * (size_t)Cout * ntaps * Cin, st);
    if (e != cudaSuccess) { cb200_set_error("conv2d_nhwc_wgrad: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const int sms = sm_count_cached();
    const char* name = "conv2d_nhwc_wgrad";
    if (Cin % 128 == 0) {
        p.cin_tiles = Cin / 128; p.total_btiles = ntaps * p.cin_tiles;
        if (ks == 3) return launch_wgrad<128, 3, KP, 3>(p, Cout, sms, st, name);     // 64 KB / stage: dY shared by 3 taps
        return launch_wgrad<128, 2, KP, 4>(p, Cout, sms, st, name);                   // 48 KB / stage
    }
    if (Cin % 64 == 0) {
        p.cin_tiles = Cin / 64; p.total_btiles = ntaps * p.cin_tiles;
        return launch_wgrad<64, 4, KP, 4>(p, Cout, sms, st, name);                    // 48 KB / stage
    }
    p.cin_tiles = Cin / 32; p.total_btiles = ntaps * p.cin_tiles;
    return launch_wgrad<32, 4, KP, 4>(p, Cout, sms, st, name);
}

// dw[N, K] (row stride ldw) = dy[M, N]^T (row stride ldy) * x[M, K] (row stride ldx)   -- linear-layer weight gradient.
extern "C" int cb200_gemm_tn_wgrad(const float* dy, long long ldy, const float* x, long long ldx, float* dw,
                                   long long ldw, int M, int N, int K, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CB200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tn_wgrad: empty problem");
    CB200_CHECK_ARG(N % kTileM == 0, "gemm_tn_wgrad: N=%d must be a multiple of 128", N);
    CB200_CHECK_ARG(K % 32 == 0, "gemm_tn_wgrad: K=%d must be a multiple of 32", K);
    CB200_CHECK_ARG(ldy % 4 == 0 && ldx % 4 == 0, "gemm_tn_wgrad: ldy/ldx must be multiples of 4 floats");
    CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) |
                      reinterpret_cast<uintptr_t>(dw)) & 15) == 0, "gemm_tn_wgrad: pointers must be 16-byte aligned");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    constexpr int kKP = 32;
    {
        uint64_t dims[5] = {(uint64_t)N, (uint64_t)M, 1, 1, 1};
        uint64_t str[5] = {4, (uint64_t)ldy * 4, (uint64_t)ldy * 4 * M, (uint64_t)ldy * 4 * M, (uint64_t)ldy * 4 * M};
        uint32_t box[5] = {32, (uint32_t)kKP, 1, 1, 1};
        if (int e = encode5(&p.tmap_dy, dy, dims, str, box, "dy")) return e;
    }
    {
        uint64_t dims[5] = {(uint64_t)K, (uint64_t)M, 1, 1, 1};
        uint64_t str[5] = {4, (uint64_t)ldx * 4, (uint64_t)ldx * 4 * M, (uint64_t)ldx * 4 * M, (uint64_t)ldx * 4 * M};
        uint32_t box[5] = {32, (uint32_t)kKP, 1, 1, 1};
        if (int e = encode5(&p.tmap_x, x, dims, str, box, "x")) return e;
    }
    p.box[0] = kKP; p.box[1] = 1; p.box[2] = 1; p.box[3] = 1;
    p.tiles[0] = (M + kKP - 1) / kKP; p.tiles[1] = 1; p.tiles[2] = 1; p.tiles[3] = 1;
    p.ntaps = 1;
    p.total_ptiles = p.tiles[0];
    p.ldw = (int)ldw;
    p.dw = dw;
    cudaError_t e = cudaMemset2DAsync(dw, sizeof(float) * ldw, 0, sizeof(float) * K, N, st);
    if (e != cudaSuccess) { cb200_set_error("gemm_tn_wgrad: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const int sms = sm_count_cached();
    const char* name = "gemm_tn_wgrad";
    if (K % 128 == 0) {
        p.cin_tiles = K / 128; p.total_btiles = p.cin_tiles;
        return launch_wgrad<128, 3, kKP, 3>(p, N, sms, st, name);
    }
    if (K % 64 == 0) {
        p.cin_tiles = K / 64; p.total_btiles = p.cin_tiles;
        return launch_wgrad<64, 4, kKP, 4>(p, N, sms, st, name);
    }
    p.cin_tiles = K / 32; p.total_btiles = p.cin_tiles;
    return launch_wgrad<32, 4, kKP, 4>(p, N, sms, st, name);
}
