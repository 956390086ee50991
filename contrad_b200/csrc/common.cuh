// Shared helpers for the contrad_b200 sm_100a kernels (error plumbing, warp/block reductions,
// vector I/O).  Everything here is device-generic; tcgen05/TMA wrappers live in tc_ptx.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#if !defined(__CUDACC__)
#include <sched.h>
#include <string.h>
#endif

#define CB200_OK 0
#define CB200_ERR_ARG 1000        // bad argument (shape / alignment / unsupported size)
#define CB200_ERR_TMAP 1001       // cuTensorMapEncode* failed or driver entry point missing

void cb200_set_error(const char* fmt, ...);

#define CB200_CHECK_ARG(cond, ...)                    \
    do {                                              \
        if (!(cond)) {                                \
            cb200_set_error(__VA_ARGS__);             \
            return CB200_ERR_ARG;                     \
        }                                             \
    } while (0)

#define CB200_CHECK_LAUNCH(name)                                                       \
    do {                                                                               \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) {                                                       \
            cb200_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));    \
            return (int)_e;                                                            \
        }                                                                              \
    } while (0)

// Every kernel launch of this library goes through this counter so that bench.py can report
// `gpu_launches` from the library itself (not a Python-side guess).
extern unsigned long long g_cb200_launches;
#define CB200_COUNT_LAUNCH() ((void)__atomic_fetch_add(&g_cb200_launches, 1ULL, __ATOMIC_RELAXED))   // one host thread per GPU under nn.DataParallel

// Once-per-device flags for cudaFuncSetAttribute (nn.DataParallel drives several devices from one process).
inline bool& cb200_device_flag(bool (&flags)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    return flags[dev & 63];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of up to NV values per thread; result broadcast to every thread.
// `scratch` must hold NV * 32 floats.  Contains two __syncthreads().
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float t = (lane < nwarps) ? scratch[i * 32 + lane] : 0.f;
        v[i] = warp_sum(t);
    }
    __syncthreads();
}

// Column reductions over a row-major [M, C] matrix with few columns: the 256 threads of a CTA are arranged as
// (256 / cc) row lanes x cc columns (cc = min(C, 256), a power of two), so all lanes stay busy when C < 256.
// After the per-thread loop over rows, fold the row lanes: result valid in the threads with row lane 0.
template <int NV>
__device__ __forceinline__ void fold_row_lanes(float (&v)[NV], float* scratch /* NV * 256 floats */, int cc) {
    const int lanes = blockDim.x / cc;
    if (lanes <= 1) return;
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 256 + threadIdx.x] = v[i];
    __syncthreads();
    if ((int)threadIdx.x < cc) {
        for (int r = 1; r < lanes; ++r) {
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] += scratch[i * 256 + r * cc + threadIdx.x];
        }
    }
}

// The three PTX helpers below have plain-C equivalents so that the asm-free kernel files can also be compiled by a host
// compiler against a CUDA emulation header (tests/emu: CPU-side checks of the SIMT kernels; never part of the product).
#if defined(__CUDACC__)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Round an fp32 value to the nearest TF32 (10-bit mantissa), ties away from zero, so that the
// tensor core's operand truncation is exact and the rounding error is unbiased.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// a / b through ONE MUFU.RCP + one multiply.  `__fdividef` is the same product in the normal range but brackets it
// with a denormal guard (FSETP + two predicated FMUL by 2^24 per quotient); callers of this helper guarantee
// |b| >= 1e-30, where the two are bit-identical.
__device__ __forceinline__ float fast_div(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
}
#else
inline float fast_div(float a, float b) { return __fdividef(a, b); }
inline float4 ldg_stream4(const float* p) { return *reinterpret_cast<const float4*>(p); }
inline void stg_stream4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
inline float round_tf32(float x) {          // cvt.rna.tf32.f32: add half an ulp of the 10-bit mantissa, truncate
    uint32_t u = __float_as_uint(x);
    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
    return __uint_as_float(u);
}
#endif

// ---- async staging: cp.async.bulk (TMA bulk copy) of a contiguous block into shared memory, completion on an mbarrier ----
// Usage: one thread does bar_init(bar, 1) + fence_barrier_init() once, then per phase bar_expect_tx(bar, total_bytes)
// followed by one or more bulk_load(...) whose sizes add up to total_bytes; every consumer thread calls
// bar_wait(bar, phase_parity).  (The host versions in the #else branch give the same protocol a synchronous copy and a
// {phase count, bytes in flight} word, for the CPU-side kernel checks of tests/emu.)
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 26) && !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
    if (!ok) __trap();
}
#else
inline void bar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_RELEASE); }
inline void fence_barrier_init() {}
inline void bar_expect_tx(uint64_t* bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t)bytes, __ATOMIC_RELAXED); }
inline void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    memcpy(dst, src, bytes);
    const uint64_t v = __atomic_load_n(bar, __ATOMIC_RELAXED);
    const uint64_t left = (v & 0xffffffffull) - bytes;                      // bytes of this phase still in flight
    __atomic_store_n(bar, left ? ((v & ~0xffffffffull) | left) : (((v >> 32) + 1) << 32), __ATOMIC_RELEASE);
}
inline void bar_wait(uint64_t* bar, uint32_t parity) {
    while (((__atomic_load_n(bar, __ATOMIC_ACQUIRE) >> 32) & 1u) == parity) sched_yield();
}
#endif
