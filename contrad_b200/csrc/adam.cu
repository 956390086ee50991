// Fused multi-tensor Adam for sm_100a (HBM-bound): one launch updates up to 48 parameter tensors.
// Reference: torch.optim.Adam as configured by train_gan.py:273-274 (betas from gin, eps 1e-8, no weight decay,
// no amsgrad), i.e. per element
//     m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// The reference's foreach implementation issues ~10 multi-tensor kernels per optimiser step; here the 16
// bytes/element of read-modify-write traffic happen exactly once.  Tensor tables travel as kernel parameters.
#include "common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kMaxTensors = 48;
constexpr int kElemsPerCta = kT * 4 * 4;      // 4 float4 per thread

struct AdamBatch {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* m[kMaxTensors];
    float* v[kMaxTensors];
    long long numel[kMaxTensors];
    int cta_begin[kMaxTensors + 1];
    int n;
    float lr, beta1, beta2, eps, bc1, bc2_sqrt;
    const float* hyper;     // optional device-resident {lr, 1-b1^t, sqrt(1-b2^t)} overriding lr/bc1/bc2_sqrt
};

__global__ void __launch_bounds__(kT) adam_kernel(const __grid_constant__ AdamBatch a) {
    int l = 0;
#pragma unroll 1
    while (l + 1 < a.n && (int)blockIdx.x >= a.cta_begin[l + 1]) ++l;
    const long long base = (long long)(blockIdx.x - a.cta_begin[l]) * kElemsPerCta;
    const long long n = a.numel[l];
    float* __restrict__ p = a.p[l];
    const float* __restrict__ g = a.g[l];
    float* __restrict__ m = a.m[l];
    float* __restrict__ v = a.v[l];
    const float lr = a.hyper ? __ldg(a.hyper + 0) : a.lr;
    const float bc1 = a.hyper ? __ldg(a.hyper + 1) : a.bc1;
    const float bc2_sqrt = a.hyper ? __ldg(a.hyper + 2) : a.bc2_sqrt;
    const float step = lr / bc1;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const long long i = base + ((long long)it * kT + threadIdx.x) * 4;
        if (i >= n) break;
        const bool vec = (i + 3 < n) && ((n & 3) == 0);
        float pv[4], gv[4], mv[4], vv[4];
        if (vec) {
            float4 t;
            t = *reinterpret_cast<const float4*>(p + i); pv[0] = t.x; pv[1] = t.y; pv[2] = t.z; pv[3] = t.w;
            t = *reinterpret_cast<const float4*>(g + i); gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
            t = *reinterpret_cast<const float4*>(m + i); mv[0] = t.x; mv[1] = t.y; mv[2] = t.z; mv[3] = t.w;
            t = *reinterpret_cast<const float4*>(v + i); vv[0] = t.x; vv[1] = t.y; vv[2] = t.z; vv[3] = t.w;
        } else {
            for (int e = 0; e < 4; ++e) {
                const bool ok = i + e < n;
                pv[e] = ok ? p[i + e] : 0.f; gv[e] = ok ? g[i + e] : 0.f; mv[e] = ok ? m[i + e] : 0.f; vv[e] = ok ? v[i + e] : 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            mv[e] = a.beta1 * mv[e] + (1.f - a.beta1) * gv[e];
            vv[e] = a.beta2 * vv[e] + (1.f - a.beta2) * gv[e] * gv[e];
            const float denom = sqrtf(vv[e]) / bc2_sqrt + a.eps;
            pv[e] -= step * (mv[e] / denom);
        }
        if (vec) {
            *reinterpret_cast<float4*>(p + i) = make_float4(pv[0], pv[1], pv[2], pv[3]);
            *reinterpret_cast<float4*>(m + i) = make_float4(mv[0], mv[1], mv[2], mv[3]);
            *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        } else {
            for (int e = 0; e < 4 && i + e < n; ++e) { p[i + e] = pv[e]; m[i + e] = mv[e]; v[i + e] = vv[e]; }
        }
    }
}

}  // namespace

extern "C" {
struct cb200_adam_tensor {
    float* p; const float* g; float* m; float* v; long long numel;
};
}

namespace {
int adam_launch(const cb200_adam_tensor* tensors, int n, float lr, float beta1, float beta2, float eps, float bc1,
                float bc2_sqrt, const float* hyper, cudaStream_t st) {
    for (int begin = 0; begin < n; begin += kMaxTensors) {
        AdamBatch a;
        a.n = (n - begin < kMaxTensors) ? (n - begin) : kMaxTensors;
        a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.bc1 = bc1; a.bc2_sqrt = bc2_sqrt;
        a.hyper = hyper;
        int total = 0;
        for (int l = 0; l < a.n; ++l) {
            const cb200_adam_tensor& t = tensors[begin + l];
            CB200_CHECK_ARG(t.numel > 0, "adam_step: empty tensor");
            CB200_CHECK_ARG(((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                              reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0,
                            "adam_step: tensors must be 16-byte aligned");
            a.p[l] = t.p; a.g[l] = t.g; a.m[l] = t.m; a.v[l] = t.v; a.numel[l] = t.numel;
            a.cta_begin[l] = total;
            total += (int)((t.numel + kElemsPerCta - 1) / kElemsPerCta);
        }
        a.cta_begin[a.n] = total;
        adam_kernel<<<total, kT, 0, st>>>(a);
        CB200_COUNT_LAUNCH();
    }
    CB200_CHECK_LAUNCH("adam_step");
    return CB200_OK;
}
}  // namespace

// One Adam step (step count t >= 1 shared by all tensors) on n tensors; processed in chunks of 48 per launch.
extern "C" int cb200_adam_step(const cb200_adam_tensor* tensors, int n, float lr, float beta1, float beta2, float eps,
                               int step, void* stream) {
    CB200_CHECK_ARG(n > 0 && step >= 1, "adam_step: bad arguments");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    return adam_launch(tensors, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, nullptr, static_cast<cudaStream_t>(stream));
}

// Same update with the step-dependent scalars {lr, 1 - beta1^t, sqrt(1 - beta2^t)} read from DEVICE memory
// (`hyper`, 3 floats): nothing step-dependent is baked into the launch, so the call can live in a CUDA graph.
extern "C" int cb200_adam_step_dev(const cb200_adam_tensor* tensors, int n, const float* hyper, float beta1,
                                   float beta2, float eps, void* stream) {
    CB200_CHECK_ARG(n > 0 && hyper != nullptr, "adam_step_dev: bad arguments");
    return adam_launch(tensors, n, 0.f, beta1, beta2, eps, 1.f, 1.f, hyper, static_cast<cudaStream_t>(stream));
}
