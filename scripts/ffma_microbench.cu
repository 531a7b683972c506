// ffma_microbench.cu -- FP32 FMA throughput of one SM sub-partition on sm_100a: scalar FFMA (3 register operands) vs the packed
// fma.rn.f32x2 (FFMA2), 16 independent accumulator chains per thread, 1..8 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ffma_mb scripts/ffma_microbench.cu && scripts/ffma_mb
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a0, float b0) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    float a = a0 + threadIdx.x * 1e-6f, b = b0;
    float a2 = a * 1.0001f, b2 = b * 0.9999f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], (i & 1) ? a : a2, (i & 2) ? b : b2);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2)
                asm volatile("{\n\t.reg .b64 d, x, y;\n\tmov.b64 d, {%0, %1};\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%4, %5};\n\t"
                             "fma.rn.f32x2 d, d, x, y;\n\tmov.b64 {%0, %1}, d;\n\t}"
                             : "+f"(acc[i]), "+f"(acc[i + 1]) : "f"(a), "f"(a2), "f"(b), "f"(b2));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads = 128; threads <= 1024; threads *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, threads>>>(out, iters, 1.0001f, 1e-7f);
                else k<1><<<148, threads>>>(out, iters, 1.0001f, 1e-7f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = 148.0 * threads * 32.0 * iters;
            printf("%s warps/SMSP %d: %.3f ms, %.2f TFMA/s, %.1f FMA/clk/SM at %d MHz nominal\n", mode ? "fma.rn.f32x2" : "fma.rn.f32  ", threads / 128, ms,
                   fma / ms / 1e9, fma / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
