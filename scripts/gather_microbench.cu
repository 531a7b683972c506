// gather_microbench.cu -- what can a B200 do on the state gather alone?  10M random 128-byte rows out of a 128 MB table
// (the C4 shape), three mechanisms, several depths / occupancies.  Build: nvcc -arch=sm_100a -O3 -o gather_mb gather_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// A: register gather. 8 lanes per row; every group sums `deg` rows per output node, U loads in flight.
template <int U>
__global__ void gather_ldg(const float* __restrict__ x, const int* __restrict__ col, int deg, long long n_nodes, float* __restrict__ out) {
    const int lig = threadIdx.x & 7;
    const long long groups = (long long)gridDim.x * blockDim.x / 8;
    for (long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 8; n < n_nodes; n += groups) {
        float4 acc = make_float4(0, 0, 0, 0);
        const int* c = col + n * deg;
        for (int e = 0; e < deg; e += U) {
            float4 r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { int s = c[min(e + u, deg - 1)]; r[u] = ldg4(x + (size_t)s * 32 + 4 * lig); }
#pragma unroll
            for (int u = 0; u < U; ++u) if (e + u < deg) { acc.x += r[u].x; acc.y += r[u].y; acc.z += r[u].z; acc.w += r[u].w; }
        }
        *reinterpret_cast<float4*>(out + n * 32 + 4 * lig) = acc;
    }
}

// B: cp.async (LDGSTS) 16 B per lane into shared memory, whole node (deg rows) at once, 2 nodes in flight per group
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    unsigned d = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g));
}
template <int DEG, int STAGES>
__global__ void gather_cpasync(const float* __restrict__ x, const int* __restrict__ col, long long n_nodes, float* __restrict__ out) {
    extern __shared__ float4 sm[];   // [groups][STAGES][DEG][8 lanes]
    const int lig = threadIdx.x & 7, grp = threadIdx.x >> 3;
    float4* mine = sm + (size_t)grp * STAGES * DEG * 8;
    const long long groups = (long long)gridDim.x * blockDim.x / 8;
    long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 8;
    long long pre = n;
    for (int s = 0; s < STAGES - 1; ++s) {
        if (pre < n_nodes) { const int* c = col + pre * DEG;
#pragma unroll
            for (int u = 0; u < DEG; ++u) cp_async16(mine + (s * DEG + u) * 8 + lig, x + (size_t)c[u] * 32 + 4 * lig); }
        asm volatile("cp.async.commit_group;\n" ::);
        pre += groups;
    }
    int stage = 0;
    for (; n < n_nodes; n += groups) {
        const int ps = (stage + STAGES - 1) % STAGES;
        if (pre < n_nodes) { const int* c = col + pre * DEG;
#pragma unroll
            for (int u = 0; u < DEG; ++u) cp_async16(mine + (ps * DEG + u) * 8 + lig, x + (size_t)c[u] * 32 + 4 * lig); }
        asm volatile("cp.async.commit_group;\n" ::);
        pre += groups;
        asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 1));
        __syncwarp();
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < DEG; ++u) { float4 r = mine[(stage * DEG + u) * 8 + lig]; acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w; }
        *reinterpret_cast<float4*>(out + n * 32 + 4 * lig) = acc;
        stage = (stage + 1) % STAGES;
    }
}

// C: bulk async copies (cp.async.bulk, 128 B per row, one issuing lane per row), mbarrier completion, per-warp ring
template <int DEG, int NODES_PER_STAGE, int STAGES>
__global__ void gather_bulk(const float* __restrict__ x, const int* __restrict__ col, long long n_nodes, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    constexpr int ROWS = DEG * NODES_PER_STAGE;
    float* buf = reinterpret_cast<float*>(smraw) + (size_t)warp * STAGES * ROWS * 32;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + (size_t)nwarps * STAGES * ROWS * 128) + warp * STAGES;
    if (lane == 0) for (int s = 0; s < STAGES; ++s) {
        unsigned b = (unsigned)__cvta_generic_to_shared(bars + s);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b));
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    __syncwarp();
    const long long wstride = (long long)gridDim.x * nwarps * NODES_PER_STAGE;
    long long base = ((long long)blockIdx.x * nwarps + warp) * NODES_PER_STAGE;
    auto issue = [&](long long nb, int s) {
        unsigned b = (unsigned)__cvta_generic_to_shared(bars + s);
        if (nb >= n_nodes) return;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(ROWS * 128));
        __syncwarp();
        for (int r = lane; r < ROWS; r += 32) {
            long long node = nb + r / DEG;
            int src = col[min(node, n_nodes - 1) * DEG + r % DEG];
            unsigned d = (unsigned)__cvta_generic_to_shared(buf + ((size_t)s * ROWS + r) * 32);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];\n"
                         ::"r"(d), "l"(x + (size_t)src * 32), "r"(b) : "memory");
        }
    };
    long long pre = base;
    for (int s = 0; s < STAGES - 1; ++s) { issue(pre, s); pre += wstride; }
    int stage = 0, phase = 0;
    for (long long nb = base; nb < n_nodes; nb += wstride) {
        issue(pre, (stage + STAGES - 1) % STAGES);
        pre += wstride;
        unsigned b = (unsigned)__cvta_generic_to_shared(bars + stage);
        unsigned done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n" : "=r"(done) : "r"(b), "r"(phase));
        // 32 lanes: 4 nodes at a time, 8 lanes per node
        const int lig = lane & 7;
        for (int nn = lane >> 3; nn < NODES_PER_STAGE; nn += 4) {
            float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < DEG; ++u) {
                float4 r = *reinterpret_cast<float4*>(buf + ((size_t)stage * ROWS + nn * DEG + u) * 32 + 4 * lig);
                acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
            }
            if (nb + nn < n_nodes) *reinterpret_cast<float4*>(out + (nb + nn) * 32 + 4 * lig) = acc;
        }
        __syncwarp();
        stage = (stage + 1) % STAGES;
        if (stage == 0) phase ^= 1;
    }
}

int main(int argc, char** argv) {
    const long long N = 1000000; const int DEG = 10; const long long E = N * DEG;
    const int local = argc > 1 ? atoi(argv[1]) : 0;   // 0: uniform sources, else +-local window
    float *x, *out; int* col;
    CK(cudaMalloc(&x, N * 128)); CK(cudaMalloc(&out, N * 128)); CK(cudaMalloc(&col, E * 4));
    std::vector<int> h(E);
    unsigned long long sd = 12345;
    for (long long i = 0; i < E; ++i) {
        sd = sd * 6364136223846793005ULL + 1442695040888963407ULL;
        unsigned r = (unsigned)(sd >> 33);
        h[i] = local ? (int)(((i / DEG) + (long long)(r % (2 * local + 1)) - local + N) % N) : (int)(r % N);
    }
    CK(cudaMemcpy(col, h.data(), E * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(x, 0, N * 128));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %7.2f GB/s gathered  %6.2f Grows/s\n", name, ms, E * 128.0 / ms / 1e6, E / ms / 1e6); };
#define TIME(name, launch) { for (int w = 0; w < 2; ++w) { launch; } CK(cudaDeviceSynchronize()); cudaEventRecord(a); for (int r = 0; r < 5; ++r) { launch; } cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, a, b); report(name, ms / 5); }
    for (int ctas : {4, 8, 16}) {
        char nm[96];
        snprintf(nm, 96, "A ldg128 U=4  %2d CTAs/SM x 128 thr", ctas); TIME(nm, (gather_ldg<4><<<sms * ctas, 128>>>(x, col, DEG, N, out)));
        snprintf(nm, 96, "A ldg128 U=10 %2d CTAs/SM x 128 thr", ctas); TIME(nm, (gather_ldg<10><<<sms * ctas, 128>>>(x, col, DEG, N, out)));
    }
    {
        const int thr = 128; size_t sm2 = (size_t)(thr / 8) * 2 * DEG * 128, sm3 = (size_t)(thr / 8) * 3 * DEG * 128;
        CK(cudaFuncSetAttribute(gather_cpasync<10, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        CK(cudaFuncSetAttribute(gather_cpasync<10, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        for (int ctas : {2, 4}) {
            char nm[96];
            snprintf(nm, 96, "B cp.async 2 stages %d CTAs/SM x 128 thr", ctas); TIME(nm, (gather_cpasync<10, 2><<<sms * ctas, thr, sm2>>>(x, col, N, out)));
            snprintf(nm, 96, "B cp.async 3 stages %d CTAs/SM x 128 thr", ctas); TIME(nm, (gather_cpasync<10, 3><<<sms * ctas, thr, sm3>>>(x, col, N, out)));
        }
    }
    {
        const int thr = 128, NPS = 8, ST = 3;
        size_t smb = (size_t)(thr / 32) * ST * DEG * NPS * 128 + (thr / 32) * ST * 8 + 128;
        CK(cudaFuncSetAttribute(gather_bulk<10, NPS, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
        for (int ctas : {1, 2}) {
            char nm[96];
            snprintf(nm, 96, "C cp.async.bulk 128B rows, 8 nodes x 3 stages/warp, %d CTAs/SM", ctas);
            TIME(nm, (gather_bulk<10, NPS, ST><<<sms * ctas, thr, smb>>>(x, col, N, out)));
        }
    }
    // streaming reference: read x once, write out once
    CK(cudaMemcpy(out, x, N * 128, cudaMemcpyDeviceToDevice));
    cudaEventRecord(a); for (int r = 0; r < 5; ++r) CK(cudaMemcpyAsync(out, x, N * 128, cudaMemcpyDeviceToDevice)); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); printf("memcpy 128 MB d2d: %.3f ms (%.1f GB/s r+w)\n", ms / 5, 2 * N * 128.0 / (ms / 5) / 1e6);
    return 0;
}
