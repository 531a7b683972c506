// gather_microbench.cu -- what can a B200 do on the state gather alone?  10M random 128-byte rows out of a 128 MB table
// (the C4 shape), three mechanisms, several depths / occupancies.  Build: nvcc -arch=sm_100a -O3 -o gather_mb gather_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>
#include <cmath>

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


// D / E: the ring pipeline of the fused kernel in isolation.  One persistent CTA per SM: a loader thread stages the arc sources of
// each 16-node sub-tile (160 rows) in shared memory with ONE bulk copy, NIW issue warps land the rows in a ring of SLOTS sub-tile
// slots, NCW consume warps sum them.  MODE 0: cp.async (LDGSTS) 16 B per lane, 4 rows per warp instruction, lean loop (index LDS,
// address IMAD, copy).  MODE 1: TMA tile::gather4 (UTMALDG.2D.GATHER4), 4 rows per single-thread instruction.
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, int bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, unsigned parity) {
    unsigned done = 0, a = s32(b);
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
template <int MODE, int NIW, int NCW, int SLOTS>
__global__ void __launch_bounds__(32 * (1 + NIW + NCW), 1)
gather_ring(const float* __restrict__ x, const __grid_constant__ CUtensorMap tm, const int* __restrict__ col, long long n_sub, float* __restrict__ out) {
    constexpr int DEG = 10, SUB = 16, ROWS = DEG * SUB, NI = 2 * SLOTS;
    extern __shared__ __align__(128) unsigned char smraw[];
    float* land = reinterpret_cast<float*>(smraw);                                   // [SLOTS][ROWS][32]
    int* sidx = reinterpret_cast<int*>(smraw + (size_t)SLOTS * ROWS * 128);         // [NI][ROWS]
    __shared__ __align__(8) uint64_t landed[SLOTS], freeb[SLOTS], idxb[NI], idxfree[NI];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < SLOTS; ++i) { mb_init(&landed[i], (MODE == 0 || (MODE == 2 && (i % NIW) % 2 == 0)) ? 32 : 1); mb_init(&freeb[i], 1); }
        for (int i = 0; i < NI; ++i) { mb_init(&idxb[i], 1); mb_init(&idxfree[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long s0 = n_sub * blockIdx.x / gridDim.x, s1 = n_sub * (blockIdx.x + 1) / gridDim.x;
    const int cnt = (int)(s1 - s0);
    if (warp == 0) {
        if (lane == 0)
            for (int j = 0; j < cnt; ++j) {
                const int q = j % NI;
                if (j >= NI) mb_wait(&idxfree[q], ((j / NI) - 1) & 1);
                mb_expect(&idxb[q], ROWS * 4);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(sidx + q * ROWS)), "l"(col + (s0 + j) * ROWS), "r"(ROWS * 4), "r"(s32(&idxb[q])) : "memory");
            }
    } else if (warp <= NIW) {
        const int w = warp - 1;
        for (int j = w; j < cnt; j += NIW) {
            const int q = j % NI, slot = j % SLOTS;
            mb_wait(&idxb[q], (j / NI) & 1);
            if (j >= SLOTS) mb_wait(&freeb[slot], ((j / SLOTS) - 1) & 1);
            const int* si = sidx + q * ROWS;
            float* lb = land + (size_t)slot * ROWS * 32;
            if (MODE == 0 || (MODE == 2 && (w & 1) == 0)) {
                const int g = lane >> 3, l8 = lane & 7;
                const float* xl = x + 4 * l8;
                float* db = lb + 4 * l8;
#pragma unroll 8
                for (int r = g; r < ROWS; r += 4) {
                    const int s = si[r];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(db + r * 32)), "l"(xl + (size_t)s * 32));
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(&landed[slot])) : "memory");
                __syncwarp();
                if (lane == 0) mb_arrive(&idxfree[q]);
            } else {
                if (lane == 0) {
                    mb_expect(&landed[slot], ROWS * 128);
                    const int4* s4 = reinterpret_cast<const int4*>(si);
#pragma unroll 4
                    for (int r = 0; r < ROWS / 4; ++r) {
                        const int4 s = s4[r];
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                                     ::"r"(s32(lb + r * 128)), "l"(&tm), "r"(s32(&landed[slot])), "r"(0), "r"(s.x), "r"(s.y), "r"(s.z), "r"(s.w) : "memory");
                    }
                    mb_arrive(&idxfree[q]);
                }
                __syncwarp();
            }
        }
        if (MODE != 1) asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        const int c = warp - 1 - NIW, g = lane >> 3, l8 = lane & 7;
        for (int j = c; j < cnt; j += NCW) {
            const int slot = j % SLOTS;
            mb_wait(&landed[slot], (j / SLOTS) & 1);
            const float* lb = land + (size_t)slot * ROWS * 32 + 4 * l8;
#pragma unroll
            for (int u = 0; u < SUB / 4; ++u) {
                const int node = g + 4 * u;
                float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
                for (int e = 0; e < DEG; ++e) {
                    const float4 r = *reinterpret_cast<const float4*>(lb + (node * DEG + e) * 32);
                    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
                }
                *reinterpret_cast<float4*>(out + ((s0 + j) * SUB + node) * 32 + 4 * l8) = acc;
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&freeb[slot]);
        }
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

    {
        // checksum of variant A as the reference for D / E
        std::vector<float> hx(N * 32);
        for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u) % 1000) * 1e-3f;
        CK(cudaMemcpy(x, hx.data(), N * 128, cudaMemcpyHostToDevice));
        std::vector<float> ref(N * 32), got(N * 32);
        gather_ldg<10><<<sms * 16, 128>>>(x, col, DEG, N, out);
        CK(cudaMemcpy(ref.data(), out, N * 128, cudaMemcpyDeviceToHost));
        CUtensorMap tm;
        cuuint64_t gdim[2] = {32, (cuuint64_t)N}, gstr[1] = {128};
        cuuint32_t box[2] = {32, 1}, estr[2] = {1, 1};
        CUresult cr = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
        const long long n_sub = N / 16;
        auto check = [&](const char* nm) {
            CK(cudaMemcpy(got.data(), out, N * 128, cudaMemcpyDeviceToHost));
            double md = 0; for (size_t i = 0; i < got.size(); ++i) md = fmax(md, fabs((double)got[i] - ref[i]));
            printf("    %s max |diff| vs A = %.3g\n", nm, md);
        };
#define RING(MODE, NIW, NCW, SLOTS) { \
            size_t smr = (size_t)SLOTS * 160 * 128 + 2 * SLOTS * 160 * 4; \
            CK(cudaFuncSetAttribute(gather_ring<MODE, NIW, NCW, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr)); \
            char nm[128]; snprintf(nm, 128, "%s ring %d slots, %d issue + %d consume warps", MODE == 2 ? "F hybrid cp.async / gather4" : MODE ? "D gather4" : "E cp.async", SLOTS, NIW, NCW); \
            CK(cudaMemset(out, 0, N * 128)); \
            TIME(nm, (gather_ring<MODE, NIW, NCW, SLOTS><<<sms, 32 * (1 + NIW + NCW), smr>>>(x, tm, col, n_sub, out))); check(nm); }
        RING(0, 4, 8, 8) RING(0, 2, 8, 8) RING(0, 8, 8, 8) RING(0, 2, 10, 10) RING(0, 4, 4, 8) RING(0, 4, 4, 4)
        RING(2, 8, 8, 8) RING(2, 4, 8, 8) RING(2, 8, 4, 8)
        RING(1, 4, 8, 8) RING(1, 2, 8, 8) RING(1, 8, 8, 8) RING(1, 1, 8, 8) RING(1, 2, 10, 10)
    }
    // streaming reference: read x once, write out once
    CK(cudaMemcpy(out, x, N * 128, cudaMemcpyDeviceToDevice));
    cudaEventRecord(a); for (int r = 0; r < 5; ++r) CK(cudaMemcpyAsync(out, x, N * 128, cudaMemcpyDeviceToDevice)); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); printf("memcpy 128 MB d2d: %.3f ms (%.1f GB/s r+w)\n", ms / 5, 2 * N * 128.0 / (ms / 5) / 1e6);
    return 0;
}
