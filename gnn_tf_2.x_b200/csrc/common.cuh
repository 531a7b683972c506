// common.cuh -- error handling, launch accounting, activations and the counter-based dropout generator
// shared by every translation unit of libgnn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gnn_b200.h"

namespace gnn {

void set_error(const std::string& msg);
void count_launch(int n = 1);

#define GNN_FAIL(code, ...)                                   \
    do {                                                      \
        char _buf[512];                                       \
        snprintf(_buf, sizeof(_buf), __VA_ARGS__);            \
        ::gnn::set_error(_buf);                               \
        return (code);                                        \
    } while (0)

#define GNN_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            GNN_FAIL(GNN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define GNN_LAUNCH_CHECK()                                                                      \
    do {                                                                                        \
        ::gnn::count_launch();                                                                  \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess)                                                                  \
            GNN_FAIL(GNN_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define GNN_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != GNN_OK) return _rc; \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
static inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
#define SELU_ALPHA 1.6732632423543772f
#define SELU_SCALE 1.0507009873554805f

// Activations. exp / division use the hardware approximations (ex2.approx, rcp.approx: ~2 ulp): the ABSOLUTE error
// stays near 1e-7, far inside the 1e-4 parity tolerance, and the element-wise pass stays a handful of instructions.
// selu / elu follow TensorFlow's own formula exp(z) - 1.
// e^x as one MUFU: ex2.approx.ftz(x * log2(e)).  Results below 2^-126 flush to zero (irrelevant for every use below).
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

// Branch-free on purpose (select, not if/else): the element-wise epilogues keep many independent elements in flight.
__device__ __forceinline__ float act_apply(int act, float z) {
    switch (act) {
        case GNN_ACT_RELU: return fmaxf(z, 0.f);
        case GNN_ACT_TANH: {
            const float a = fminf(fabsf(z), 15.f);          // tanh(15) == 1 in fp32
            const float e = fast_exp(2.f * a);
            return copysignf(1.f - __fdividef(2.f, e + 1.f), z);
        }
        case GNN_ACT_SIGMOID: return __fdividef(1.f, 1.f + fast_exp(-z));
        case GNN_ACT_SELU: {
            const float neg = fmaf(SELU_SCALE * SELU_ALPHA, fast_exp(fminf(z, 0.f)), -(SELU_SCALE * SELU_ALPHA));
            const float pos = SELU_SCALE * z;
            return z > 0.f ? pos : neg;
        }
        case GNN_ACT_ELU: {
            const float neg = fast_exp(fminf(z, 0.f)) - 1.f;
            return z > 0.f ? z : neg;
        }
        case GNN_ACT_SOFTPLUS: return z > 15.f ? z : __logf(1.f + fast_exp(fminf(z, 15.f)));
        default: return z;  // linear (softmax is handled row-wise by its caller)
    }
}

// derivative of the activation expressed through its OUTPUT y = act(z) (no pre-activation needs storing)
__device__ __forceinline__ float act_grad_from_output(int act, float y) {
    switch (act) {
        case GNN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case GNN_ACT_TANH: return 1.f - y * y;
        case GNN_ACT_SIGMOID: return y * (1.f - y);
        case GNN_ACT_SELU: return y > 0.f ? SELU_SCALE : y + SELU_SCALE * SELU_ALPHA;
        case GNN_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
        case GNN_ACT_SOFTPLUS: return 1.f - expf(-y);
        default: return 1.f;
    }
}

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}

// key of one (call seed, dropout position, loop iteration); same formula as keras_compat.dropout_key
__host__ __device__ __forceinline__ uint32_t dropout_key(uint32_t seed, uint32_t stream, uint32_t step) {
    uint32_t a = fmix32(seed + 0x9E3779B9u * (stream + 1u));
    return fmix32(a ^ (step * 0x85EBCA6Bu + 0x27D4EB2Fu));
}

// dropout seed of a call: by value, or -- for calls captured in a CUDA graph, whose arguments are frozen -- from device memory
template <typename Params>
__device__ __forceinline__ uint32_t call_seed(const Params& p) { return p.seed_dev ? __ldg(p.seed_dev) : p.seed; }

// keep decision of element idx = row * width + col
__device__ __forceinline__ bool dropout_keep(uint32_t key, uint64_t idx, float rate) {
    uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
    uint32_t v = fmix32(lo ^ key);
    v = fmix32(v + hi * 0xC2B2AE35u + 0x165667B1u);
    float u = (float)(v >> 8) * (1.0f / 16777216.0f);
    return u >= rate;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// L2 eviction policies: the gathered state rows are re-read ~degree times per iteration -> keep (evict_last);
// everything that streams through once (new state, saved aggregates) -> evict_first
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg4_hint(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st4_hint(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float4 fma4(float w, float4 r, float4 acc) {
    acc.x = fmaf(w, r.x, acc.x);
    acc.y = fmaf(w, r.y, acc.y);
    acc.z = fmaf(w, r.z, acc.z);
    acc.w = fmaf(w, r.w, acc.w);
    return acc;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

}  // namespace gnn
