// TEST INFRASTRUCTURE: just enough of the CUDA programming model to run simple SIMT kernels of contrad_b200/csrc
// single-threaded on the host (one thread per block; other lanes of a warp read as zero), so that their index arithmetic
// and per-element math can be checked without a GPU.  Used by tests/test_host_logic.py on sections of the .cu files
// delimited by "[host-testable: ...]" markers.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <algorithm>
using std::max;
using std::min;
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
struct dim3s { int x, y, z; };
static dim3s threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static inline float __ldg(const float* p) { return *p; }
static inline float __ldcs(const float* p) { return *p; }
static inline void __stcs(float* p, float v) { *p = v; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline void __syncthreads() {}
template <int NV> static inline void block_sum(float (&v)[NV], float*) {}      // one thread: the value is the block sum
constexpr int kT = 256;
