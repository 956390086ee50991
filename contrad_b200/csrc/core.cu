// Library-wide plumbing of the contrad_b200 C ABI: last-error string, version, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

unsigned long long g_cb200_launches = 0;

static thread_local char t_last_error[512] = "";

void cb200_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_last_error, sizeof(t_last_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* cb200_last_error(void) { return t_last_error; }

extern "C" int cb200_version(void) { return 100; }   // round 1, revision 00

extern "C" unsigned long long cb200_launch_count(void) { return __atomic_load_n(&g_cb200_launches, __ATOMIC_RELAXED); }

extern "C" void cb200_reset_launch_count(void) { __atomic_store_n(&g_cb200_launches, 0ULL, __ATOMIC_RELAXED); }

extern "C" void cb200_add_launch_count(long long n) {
    (void)__atomic_fetch_add(&g_cb200_launches, (unsigned long long)n, __ATOMIC_RELAXED);
}

extern "C" int cb200_device_arch(int device, int* major, int* minor) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        cb200_set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
        return (int)e;
    }
    *major = prop.major;
    *minor = prop.minor;
    return CB200_OK;
}
