// Shared helpers for the contrad_b200 sm_100a kernels (error plumbing, warp/block reductions,
// vector I/O).  Everything here is device-generic; tcgen05/TMA wrappers live in tc_ptx.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
#define CB200_COUNT_LAUNCH() (++g_cb200_launches)

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
