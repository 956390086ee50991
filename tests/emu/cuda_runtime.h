// TEST INFRASTRUCTURE - a small CUDA execution-model emulator for the host.
//
// The asm-free kernel files of contrad_b200/csrc (SIMT kernels: BatchNorm, spectral norm, losses, Adam, StyleGAN2
// elementwise / FIR ops, the light augmentations) are compiled by g++ against THIS header, which stands in for
// <cuda_runtime.h> (tests/emu is put on the include path; the real CUDA headers are not), so that the very kernel source
// that nvcc compiles for sm_100a can be executed and checked on a machine without a GPU:
//   * a launch `k<<<grid, block, smem, stream>>>(args)` is rewritten (tests/emu/build_emu.py) to emu::launch(...), which runs
//     the blocks one after another and the threads of a block as real host threads, so __syncthreads(), warp shuffles,
//     shared memory, atomics and early-exiting threads behave as on the device;
//   * __shared__ variables become function-local statics (one block runs at a time), dynamic shared memory is a per-block
//     128-byte-aligned buffer;
//   * the runtime calls the launchers make (cudaMemsetAsync, cudaFuncSetAttribute, cudaGetLastError ...) are trivial.
// It is slow (one OS thread per CUDA thread) and exists only for tests at small shapes.  The product never sees it.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

// ---------------------------------------------------------------- vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

inline thread_local uint3 threadIdx = {0, 0, 0};
inline thread_local uint3 blockIdx = {0, 0, 0};
inline thread_local dim3 blockDim, gridDim;
constexpr int warpSize = 32;

// ---------------------------------------------------------------- runtime API used by the launchers
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }   // a 4-SM "GPU"
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 4; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---------------------------------------------------------------- block / warp context
namespace emu {

struct Warp {
    std::barrier<> bar;
    uint64_t slot[32];
    int lanes;
    explicit Warp(int n) : bar(n), lanes(n) {}
};

struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    char* dyn;
    Block(int nthreads, size_t smem) : bar(nthreads), dyn(nullptr) {
        for (int t = 0; t < nthreads; t += 32) warps.emplace_back(new Warp(std::min(32, nthreads - t)));
        if (smem) dyn = static_cast<char*>(aligned_alloc(128, (smem + 127) / 128 * 128));
    }
    ~Block() { free(dyn); }
};

inline thread_local Block* blk = nullptr;
inline thread_local int tid = 0;           // linear thread index in the block
inline void* dyn_smem() { return blk->dyn; }

struct Cfg {
    dim3 grid, block;
    size_t smem;
};
inline Cfg cfg(dim3 g, dim3 b, size_t smem = 0, cudaStream_t = nullptr) { return Cfg{g, b, smem}; }

// The threads of a block are host threads created ONCE per launch; they walk the grid block by block in lockstep (a
// fresh Block context - barriers, dynamic shared memory - is installed between two blocks by the completion step of the
// `next` barrier, while every thread is parked in it).
template <class F>
void launch(const Cfg& c, F&& body) {
    const int nthreads = (int)(c.block.x * c.block.y * c.block.z);
    const long long nblocks = (long long)c.grid.x * c.grid.y * c.grid.z;
    if (nthreads <= 0 || nblocks <= 0) return;
    std::unique_ptr<Block> cur(new Block(nthreads, c.smem));
    auto install_next = [&]() noexcept { cur.reset(new Block(nthreads, c.smem)); };
    std::barrier<decltype(install_next)> next(nthreads, install_next);
    std::vector<std::thread> threads;
    threads.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        threads.emplace_back([&, t]() {
            tid = t;
            threadIdx = uint3{t % c.block.x, (t / c.block.x) % c.block.y, t / (c.block.x * c.block.y)};
            blockDim = c.block;
            gridDim = c.grid;
            for (long long i = 0; i < nblocks; ++i) {
                Block* b = cur.get();
                blk = b;
                blockIdx = uint3{(unsigned)(i % c.grid.x), (unsigned)((i / c.grid.x) % c.grid.y),
                                 (unsigned)(i / ((long long)c.grid.x * c.grid.y))};
                body();
                b->warps[t / 32]->bar.arrive_and_drop();         // an exited thread no longer takes part in barriers
                b->bar.arrive_and_drop();
                next.arrive_and_wait();
            }
        });
    }
    for (auto& th : threads) th.join();
}

template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// exchange through the warp's slots: write own, barrier, read `src` lane, barrier
template <class T> inline T shuffle(T v, int src_lane) {
    Warp& w = *blk->warps[tid / 32];
    const int lane = tid % 32;
    w.slot[lane] = to_bits(v);
    w.bar.arrive_and_wait();
    const T r = (src_lane >= 0 && src_lane < w.lanes) ? from_bits<T>(w.slot[src_lane]) : v;
    w.bar.arrive_and_wait();
    return r;
}

}  // namespace emu

inline void __syncthreads() { emu::blk->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::blk->warps[emu::tid / 32]->bar.arrive_and_wait(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu::shuffle(v, (emu::tid % 32) ^ lane_mask); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) { return emu::shuffle(v, (emu::tid % 32) + delta); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shuffle(v, src); }

// ---------------------------------------------------------------- memory / atomics
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }

inline float atomicAdd(float* p, float v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
    float f;
    do {
        memcpy(&f, &old, 4);
        f += v;
        memcpy(&want, &f, 4);
    } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    memcpy(&f, &old, 4);
    return f;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------- math intrinsics
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.f / a; }
inline float __saturatef(float x) { return x >= 0.f ? (x <= 1.f ? x : 1.f) : 0.f; }     // NaN -> 0 like the device
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(uint32_t i) { float f; memcpy(&f, &i, 4); return f; }
using std::max;
using std::min;
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long min(int a, long long b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }
inline long long max(int a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
