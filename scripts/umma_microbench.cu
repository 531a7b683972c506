// umma_microbench.cu -- how long does the Dense layer of one 128-node tile take on the tcgen05 tensor core?
// One CTA per SM, one issuing thread: NMMA x tcgen05.mma.kind::tf32 (M = 128, N, K = 8, A from tensor memory, B from shared memory),
// tcgen05.commit -> mbarrier, wait; clock64 around it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mb umma_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(const void* ptr, int lbo, int sbo) {
    uint64_t d = (uint64_t)((s32(ptr) & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16; d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32; d |= 1ull << 46; return d;
}
template <int N, int NMMA, bool FROM_TMEM>
__global__ void k(long long* out, int reps) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint32_t tm; __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 0.001f * (i % 97);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tm)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tm;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int i = 0; i < NMMA; ++i) {
                const uint64_t b = desc(sm + (i % 9) * N * 8, (N / 8) * 128, 128);
                if (FROM_TMEM) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(base), "r"(base + 256 + 8 * (i % 9)), "l"(b), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
                else { const uint64_t a = desc(sm + 8192 + (i % 9) * 1024, 16 * 128, 128);
                       asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(base), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory"); }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
            unsigned done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&bar)), "r"(r & 1) : "memory");
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = (t1 - t0) / reps;
    }
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}
int main() {
    long long* d; CK(cudaMalloc(&d, 8)); long long h;
#define RUN(N, NM, T) { CK(cudaFuncSetAttribute(k<N, NM, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); k<N, NM, T><<<148, 128, 65536>>>(d, 200); CK(cudaDeviceSynchronize()); \
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost)); printf("N=%3d  %2d MMAs  A from %s: %6lld cycles per batch (%5.1f per MMA)\n", N, NM, T ? "TMEM" : "smem", h, (double)h / NM); }
    RUN(32, 27, true) RUN(32, 18, true) RUN(32, 9, true) RUN(32, 1, true) RUN(64, 18, true) RUN(64, 9, true) RUN(128, 9, true) RUN(256, 9, true)
    RUN(32, 27, false) RUN(64, 18, false) RUN(16, 27, true)
    return 0;
}
