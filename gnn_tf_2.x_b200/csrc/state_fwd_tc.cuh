// state_fwd_tc.cuh -- the fused forward iteration kernel with the Dense layer on the 5th-generation tensor cores
// (tcgen05.mma, operands in TENSOR MEMORY) -- the kernel of the headline shape.
//
// Why (measured, profiles/r2_*): the warp-level tensor-core path (mma.sync m16n8k8 TF32, SASS HMMA.1688.F32.TF32) that round 1
// used for the fp32-accurate 3xTF32 Dense layer runs at ~32 pipe cycles per instruction on sm_100a; the 108 of them per 16 nodes
// kept the legacy HMMA sub-pipe 85 % busy and bounded the iteration at ~0.13 ms BEFORE any gather (state_fwd_ws.cuh).  One
// tcgen05.mma of shape M=128, N=DP, K=8 does the work of 32 of those instructions on the real tensor pipe; the contraction of
// a 128-node tile (27 instructions, one issuing thread) disappears from the critical path.
//
// One persistent CTA per SM, 24 warps (22 working), six roles (register budgets re-split with setmaxnreg):
//   loader warp  (1): bulk async copies (cp.async.bulk) of row pointers / per-node scales / arc sources of each 64-node staging
//                     tile, three tiles ahead (a CTA owns a contiguous range of tiles: every copy is one contiguous block).
//   issue warps  (4): warp c lands sub-tile c (16 nodes) of every staging tile: per 4 source rows one index load from shared
//                     memory, one address multiply-add, one 16-byte cp.async per lane, into a ring of 8 / 16 slots; completion
//                     on the slot's mbarrier LANDED (cp.async.mbarrier.arrive.noinc).  They never wait for data.
//   sum warps    (8): sub-tile j -> warp j % 8: segment sums out of the ring the moment the rows have landed (8 lanes per row,
//                     stored order, packed add.f32x2, deterministic) -> 16 aggregate rows in the 16-deep aggregate FIFO; the 20 KB
//                     slot goes straight back to the issue warp (the ring only saturates the memory system when ALL of it is in flight)
//   compute warps (8 = 2 quads): quad g takes the 128-node tiles g, g+2, ...; warp q of the quad owns nodes 32q..32q+31 of the
//                     tile = TMEM lanes 32q..32q+31, ONE THREAD PER NODE from the staging of the operands on:
//                       1. own state row + constant row straight from global memory (256-bit loads, one 32-byte sector each),
//                       2. wait LANDED -> segment sums of its 2 sub-tiles out of the ring in stored order (8 lanes per row,
//                          packed add.f32x2, deterministic, no atomics) -> warp-private scratch -> slots released at once,
//                       3. transposition through the scratch (XOR-swizzled: conflict-free both ways): thread = node,
//                       4. hi / lo split (hi = tf32 part, lo = exact remainder) of [x | agg | cst] -> tcgen05.st into the quad's
//                          A operand in tensor memory (144 columns), mbarrier A_FULL,
//                       5. wait D_FULL -> tcgen05.ld of the node's DP accumulators -> bias / activation / affine -> 256-bit
//                          stores of the new state (node-range partition: the same stores into every peer whose bit is set
//                          in the node's peer mask, over NVLink) + convergence test against the own row still in
//                          registers (+ BatchNormalization batch statistics when training).
//   MMA warp     (1): one thread: wait A_FULL -> 27 x tcgen05.mma.kind::tf32 (A from tensor memory, B = weights in shared
//                     memory, canonical K-major layout, pre-split hi / lo): D = A_hi B_hi + A_lo B_hi + A_hi B_lo ->
//                     tcgen05.commit -> mbarrier D_FULL.
// Tensor memory: per quad [D (DP) | A_hi (2 DP + 8 CS) | A_lo (same)] = 176 -> 192 columns at DP = 32; 512 allocated.
//
// Used for single-Dense-layer state nets (what the reference builds by default) with padded state width 16 or 32, a constant
// row of at most 16 floats, uniform row weights ('sum' / 'average' / 'normalized' aggregation) and no active dropout; every
// other case runs state_fwd_ws.cuh or the symmetric kernel.
#pragma once
#include "state_fwd_ws.cuh"

namespace gnn {

constexpr int TC_TILE = 128;                 // nodes per tcgen05 tile (M)
constexpr int TC_QUADS = 2;
constexpr int TC_COMPUTE = 128 * TC_QUADS;   // warps 0-7
constexpr int TC_ISSUE = 128;                // warps 8-11
constexpr int TC_SUM = 256;                  // warps 12-19
constexpr int TC_LOADER = 32;                // warp 20
constexpr int TC_MMA = 32;                   // warp 21
constexpr int TC_IDLE = 64;                  // warps 22-23: give their registers back and wait at the final barrier
constexpr int TC_THREADS = TC_COMPUTE + TC_ISSUE + TC_SUM + TC_LOADER + TC_MMA + TC_IDLE;   // 768 = 24 warps = 6 per SM sub-partition, 80 registers at launch
constexpr int TC_REGS_LAUNCH = 80, TC_REGS_COMPUTE = 128, TC_REGS_SUM = 64, TC_REGS_ISSUE = 48, TC_REGS_SMALL = 40, TC_REGS_IDLE = 24;
// The register file is PHYSICALLY split over the 4 SM sub-partitions (warp w lives on sub-partition w % 4): setmaxnreg.inc can only
// draw on registers released on its own sub-partition.  Every sub-partition holds 2 compute + 1 issue + 2 sum + 1 loader / MMA / idle
// warp and was given 6 x 80 registers per lane at launch (an 18-warp launch left two sub-partitions with 4 x 96 = 384 < 400: the
// compute warps there never got their registers and the kernel hung -- measured the hard way)
static_assert(TC_THREADS / 32 % 4 == 0, "the same number of warps on every sub-partition");
static_assert(2 * TC_REGS_COMPUTE + 2 * TC_REGS_SUM + TC_REGS_ISSUE + TC_REGS_SMALL <= (TC_THREADS / 128) * TC_REGS_LAUNCH, "registers per lane of one SM sub-partition");
constexpr int TC_AGGQ = 16;                  // aggregate FIFO: sub-tiles (16 rows each) between the sum warps and the compute warps
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_ROWQ = 8, TC_COLQ = 3;      // staging buffers (64-node tiles), as in state_fwd_ws.cuh
constexpr int TC_SLOTS = 16;

#ifndef GNN_TC_SLEEP
#define GNN_TC_SLEEP 32
#endif

// ---- tcgen05 / tensor-memory PTX -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, int ncols) {     // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int ncols) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]; M = 128, K = 8 (tf32)
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// thread i of the warp <-> TMEM lane (lane field of taddr) + i; N consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ float4 lds4(unsigned addr) {      // volatile: stays where it is written (loads ahead of the adds that consume them)
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major operand without swizzle (UMMA "INTERLEAVE"): core matrices of 8 rows x 16 bytes
// (128 contiguous bytes); lbo = byte distance between the two 16-byte K pieces of one K = 8 step, sbo = byte distance between
// 8-row blocks.  Fields in units of 16 bytes; bits [46,48) = 1 (sm_100 descriptor version)
__device__ __forceinline__ uint64_t tc_smem_desc(const void* ptr, int lbo_bytes, int sbo_bytes) {
    uint64_t d = (uint64_t)((smem_u32(ptr) & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;
    return d;
}
// instruction descriptor: D fp32, A / B tf32, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t tc_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {      // one 32-byte sector
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, const float (&v)[8]) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// v[j] of lane l = value (row l, column j).  After the call v[0] of lane l holds the sum over all 32 rows of column l % W
// (butterfly: every step halves the values a lane keeps and exchanges the other half with the lane `w` away)
template <int W>
__device__ __forceinline__ void warp_column_sums(float (&v)[W], int lane) {
#pragma unroll
    for (int w = W / 2; w >= 1; w >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = upper ? v[i] : v[i + w];
            const float keep = upper ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
#pragma unroll
    for (int w = W; w < 32; w <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], w);     // W < 32: lanes l, l + W hold halves of a column
}

// k-steps (8 inputs each): own state (DP / 8), aggregated state (DP / 8), constant row (CP / 8, CP padded to 8)
static inline int tc_ksteps(const NetLayout& lay) { return 2 * (lay.DP / 8) + (lay.CP + 7) / 8; }
// B operand: per k-step [2 K pieces][DP / 8 row blocks][8 rows][4 floats] = DP * 8 floats, hi and lo copies; + bias, affine a / c
static inline size_t tc_weight_floats(const NetLayout& lay) { return (size_t)tc_ksteps(lay) * lay.DP * 8 * 2 + 3 * (size_t)lay.DP; }
static inline size_t tc_smem_bytes(const NetLayout& lay, int ring, int capc, bool bn_train) {
    size_t fl = tc_weight_floats(lay) + (size_t)TC_AGGQ * WS_SUB * lay.DP + (size_t)ring * lay.DP + TC_ROWQ * 68 + TC_ROWQ * WS_TN +
                TC_COLQ * (size_t)(capc + WS_COLPAD) + (bn_train ? (size_t)(TC_COMPUTE / 32) * 2 * lay.DP * 2 : 0);
    return fl * 4 + 128;   // weights, aggregate FIFO (16 sub-tiles), ring, row pointers / scales, arc sources, BN sums (fp64); + alignment slack
}

// PEERS: node-range partition over several GPUs (rows other ranks gather from are also stored into their memory).  A template
// parameter, not a run-time test: the remote stores in the epilogue of the compute warps (capped at 128 registers) cost the
// single-GPU kernel 12 % when they were merely branched around.
template <int DP, bool PEERS>
__global__ void __launch_bounds__(TC_THREADS, 1) state_iter_tc_kernel(const IterParams p) {
    constexpr int TN = WS_TN, LPN = DP / 4;      // 64-node staging tiles, 16-node sub-tiles (as state_fwd_ws.cuh)
    constexpr int GPW = 32 / LPN;                // source rows per warp-wide cp.async / lane groups per warp
    constexpr int NPG = WS_SUB / GPW;            // nodes per lane group in the segment-sum phase
    constexpr int KX = DP / 8;                   // k-steps of a state row
    static_assert(DP == 16 || DP == 32, "padded state width");

    if (!iter_begin(p)) return;

    const NetLayout& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int CP = net.CP, capc = p.scol_cap, capb = capc + WS_COLPAD;
    const int CS = (CP + 7) / 8, KSTEPS = 2 * KX + CS, KA = 8 * KSTEPS;      // A columns per copy (hi or lo)
    const int nslot = p.ring_slots, slotcap = p.slot_rows;

    // The dynamic shared memory starts right behind the static variables: only 16-byte alignment is guaranteed (an alignment
    // attribute on the extern array does not move it).  With the base at 112 mod 128 every 128-byte state row of the landing ring
    // straddled two shared-memory lines and the whole kernel ran 15 % slower: align by hand (tc_smem_bytes reserves the slack).
    extern __shared__ __align__(16) float smem_raw[];
    float* smem = smem_raw + (((128u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u) >> 2);
    float* sBhi = smem;                                  // [KSTEPS][2][DP / 8][8][4]
    float* sBlo = sBhi + (size_t)KSTEPS * DP * 8;
    float* sBias = sBlo + (size_t)KSTEPS * DP * 8;       // [DP]
    float* sAff = sBias + DP;                            // a[DP], c[DP]
    float* agg0 = sAff + 2 * DP;                         // [TC_AGGQ][16][DP]: aggregated states, 16-byte pieces XOR-swizzled by the row
    float* land0 = agg0 + TC_AGGQ * WS_SUB * DP;         // [nslot][slotcap][DP]
    int* srow0 = reinterpret_cast<int*>(land0 + (size_t)nslot * slotcap * DP);   // [TC_ROWQ][68]
    float* sscale0 = reinterpret_cast<float*>(srow0 + TC_ROWQ * 68);             // [TC_ROWQ][TN]
    int* scol0 = reinterpret_cast<int*>(sscale0 + TC_ROWQ * TN);                 // [TC_COLQ][capb]
    double* bn_acc = reinterpret_cast<double*>(scol0 + TC_COLQ * capb);          // [8][2][DP] (training-mode BatchNormalization)
    __shared__ int s_flag;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t bar_landed[TC_SLOTS], bar_free[TC_SLOTS], bar_cols[TC_COLQ], bar_colfree[TC_COLQ], bar_rows[TC_ROWQ], bar_rowfree[TC_ROWQ],
        bar_aggfull[TC_AGGQ], bar_aggfree[TC_AGGQ], bar_afull[TC_QUADS], bar_xfull[TC_QUADS], bar_dfull[TC_QUADS];

    // B operand: element (n, k) of k-step ks at [ks][k / 4 (piece)][n / 8][n % 8][k % 4]; K order = [x | agg | cst]
    for (int i = tid; i < KSTEPS * DP * 8; i += TC_THREADS) {
        const int e = i & 3, r8 = (i >> 2) & 7, nb = (i >> 5) % (DP / 8), kc = ((i >> 5) / (DP / 8)) & 1, ks = (i >> 5) / (DP / 8) / 2;
        const int n = 8 * nb + r8, k = 8 * ks + 4 * kc + e;
        int row;    // input row of the packed Dense kernel [state(DP) | agg(DP) | cst(CP)]
        if (k < DP) row = k;
        else if (k < 2 * DP) row = k;
        else row = (k - 2 * DP) < CP ? k : -1;
        const float w = row >= 0 ? __ldg(p.wpack + net.w_off[0] + row * DP + n) : 0.f;
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        sBhi[i] = hi;
        sBlo[i] = w - hi;
    }
    for (int i = tid; i < DP; i += TC_THREADS) {
        sBias[i] = __ldg(p.wpack + net.b_off[0] + i);
        sAff[i] = __ldg(p.wpack + net.aff_off + i);
        sAff[DP + i] = __ldg(p.wpack + net.aff_off + DP + i);
    }
    if (p.bn_train)
        for (int i = tid; i < (TC_COMPUTE / 32) * 2 * DP; i += TC_THREADS) bn_acc[i] = 0.;
    if (tid == 0) {
        s_flag = 0;
        for (int i = 0; i < TC_SLOTS; ++i) { mbar_init(&bar_landed[i], 32); mbar_init(&bar_free[i], 1); }
        for (int i = 0; i < TC_COLQ; ++i) { mbar_init(&bar_cols[i], 1); mbar_init(&bar_colfree[i], WS_NSUB); }
        for (int i = 0; i < TC_ROWQ; ++i) { mbar_init(&bar_rows[i], 1); mbar_init(&bar_rowfree[i], 2 * WS_NSUB); }   // 4 issue + 4 sum warps
        for (int i = 0; i < TC_AGGQ; ++i) { mbar_init(&bar_aggfull[i], 1); mbar_init(&bar_aggfree[i], 1); }
        for (int i = 0; i < TC_QUADS; ++i) { mbar_init(&bar_afull[i], 4); mbar_init(&bar_xfull[i], 4); mbar_init(&bar_dfull[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&s_tmem, TC_TMEM_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the B operand written above is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&s_tmem);

    // this CTA's staging tiles (64 nodes): a contiguous range; compute tiles (128 nodes) = pairs of staging tiles
    const long long ntiles = (p.N + TN - 1) / TN;
    const long long t0 = ntiles * blockIdx.x / gridDim.x;
    const int ntl = (int)(ntiles * (blockIdx.x + 1) / gridDim.x - t0);
    const int ntl2 = (ntl + 1) / 2;                       // 128-node tiles (the last one may hold one staging tile only)
    const int stage_cols = (DP + 2 * KA + 63) & ~63;      // TMEM columns of one quad: D | A_hi | A_lo (D aligned to its own width)

    if (warp >= 22) {
        // ============================================== IDLE WARPS ==============================================
        // Present only so that every SM sub-partition holds the same number of warps (setmaxnreg pools are per sub-partition);
        // their registers go back to the pool.  Rows that other GPUs gather from are stored into the peers by the compute
        // threads themselves (see the epilogue): a separate copy role was latency-bound -- the two CTAs at the edges of a
        // locality-friendly partition forward 2048 rows each and finished 70 us after everybody else.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_IDLE));
    } else if (warp == 21) {
        // ============================================== MMA WARP ===============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_SMALL));
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_tf32(DP);
            const int kstep_bytes = DP * 8 * 4;           // one k-step of B: 2 pieces x DP / 8 blocks x 128 bytes
            for (int t = 0; t < ntl2; ++t) {
                const int quad = t & 1;
                const uint32_t d = tmem_base + quad * stage_cols, a_hi = d + DP, a_lo = a_hi + KA;
                // the own-state / constant-row part of the contraction starts as soon as those columns are staged (before the
                // aggregates exist); the aggregated-state part follows -- only 3 KX instructions sit behind the segment sums
                auto mma3 = [&](int ks, bool first) {
                    const uint64_t bhi = tc_smem_desc(reinterpret_cast<const char*>(sBhi) + (size_t)ks * kstep_bytes, (DP / 8) * 128, 128);
                    const uint64_t blo = tc_smem_desc(reinterpret_cast<const char*>(sBlo) + (size_t)ks * kstep_bytes, (DP / 8) * 128, 128);
                    tc_mma_tf32_ts(d, a_hi + 8 * ks, bhi, idesc, first ? 0u : 1u);
                    tc_mma_tf32_ts(d, a_lo + 8 * ks, bhi, idesc, 1);
                    tc_mma_tf32_ts(d, a_hi + 8 * ks, blo, idesc, 1);
                };
                mbar_wait(&bar_xfull[quad], (t >> 1) & 1);
                tc_fence_after();
                for (int ks = 0; ks < KX; ++ks) mma3(ks, ks == 0);                  // own state: k-steps [0, KX)
                for (int ks = 2 * KX; ks < KSTEPS; ++ks) mma3(ks, false);           // constant row: k-steps [2 KX, KSTEPS)
                mbar_wait(&bar_afull[quad], (t >> 1) & 1);
                tc_fence_after();
                for (int ks = KX; ks < 2 * KX; ++ks) mma3(ks, false);               // aggregated state: k-steps [KX, 2 KX)
                tc_commit(&bar_dfull[quad]);              // arrives once every MMA above has completed (accumulator ready, A free)
            }
        }
    } else if (warp == 20) {
        // ============================================ LOADER WARP ==============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_SMALL));
        const long long E = p.E;
        auto fetch_bounds = [&](int base) {
            return base + lane <= ntl ? __ldg(p.rowptr + min((t0 + base + lane) * TN, p.N)) : 0;
        };
        int bcur = fetch_bounds(0), bnext = fetch_bounds(32);
        for (int s = 0; s < ntl; ++s) {
            const long long n0 = (t0 + s) * TN;
            const int q8 = s & (TC_ROWQ - 1), q3 = s % TC_COLQ;
            if ((s & 31) == 0 && s > 0) { bcur = bnext; bnext = fetch_bounds(s + 32); }
            const int e0 = __shfl_sync(0xffffffffu, bcur, s & 31);
            const int e1 = (s & 31) == 31 ? __shfl_sync(0xffffffffu, bnext, 0) : __shfl_sync(0xffffffffu, bcur, (s + 1) & 31);
            const int e0a = e0 & ~3;
            const int narc = min(e1 - e0a, capb - 4), narc4 = (narc + 3) & ~3;
            const bool arcs_safe = (long long)e0a + narc4 <= E;
            const bool rows_safe = n0 + 68 <= p.N + 1 && n0 + TN <= p.N;
            if (s >= TC_ROWQ) mbar_wait<64>(&bar_rowfree[q8], ((s / TC_ROWQ) - 1) & 1);
            int* srow = srow0 + q8 * 68;
            float* ssc = sscale0 + q8 * TN;
            if (rows_safe) {
                if (lane == 0) {
                    mbar_expect_tx(&bar_rows[q8], 68 * 4 + TN * 4);
                    bulk_copy_g2s(srow, p.rowptr + n0, 68 * 4, &bar_rows[q8]);
                    bulk_copy_g2s(ssc, p.row_scale + n0, TN * 4, &bar_rows[q8]);
                }
            } else {   // last tile of the graph: element-wise, nothing is read past the end of an array
                for (int i = lane; i <= TN; i += 32) srow[i] = __ldg(p.rowptr + min(n0 + i, p.N));
                for (int i = lane; i < TN; i += 32) ssc[i] = n0 + i < p.N ? __ldg(p.row_scale + n0 + i) : 0.f;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_rows[q8]);
            }
            if (s >= TC_COLQ) mbar_wait<64>(&bar_colfree[q3], ((s / TC_COLQ) - 1) & 1);
            int* sc = scol0 + (size_t)q3 * capb;
            if (arcs_safe) {
                if (lane == 0) {
                    mbar_expect_tx(&bar_cols[q3], narc4 * 4);
                    if (narc4 > 0) bulk_copy_g2s(sc, p.col + e0a, narc4 * 4, &bar_cols[q3]);
                }
            } else {
                for (int i = lane; i < narc; i += 32) sc[i] = __ldg(p.col + e0a + i);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_cols[q3]);
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ============================================ ISSUE WARPS ==============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_ISSUE));
        const int c = warp - 8;                                 // my sub-tile of every staging tile
        const int grp = lane / LPN, lig = lane % LPN;
        const unsigned landed_u32 = smem_u32(bar_landed), free_u32 = smem_u32(bar_free);
        const float* xl = p.x_in + 4 * lig;
        const int slot_shift = nslot == 8 ? 3 : 4;
        for (int s = 0; s < ntl; ++s) {
            const int q8 = s & (TC_ROWQ - 1), q3 = s % TC_COLQ;
            mbar_wait(&bar_rows[q8], (s / TC_ROWQ) & 1);
            mbar_wait(&bar_cols[q3], (s / TC_COLQ) & 1);
            const int* srow = srow0 + q8 * 68;
            const int eb = srow[0];
            const int a0 = srow[WS_SUB * c] - eb;
            const int cnt = max(0, min(min(srow[WS_SUB * c + WS_SUB] - eb, capc) - a0, slotcap));
            const int j = WS_NSUB * s + c, slot = j & (nslot - 1);
            if (j >= nslot) mbar_wait_u32(free_u32 + 8 * slot, ((j >> slot_shift) - 1) & 1);
            const int* si = scol0 + (size_t)q3 * capb + (eb & 3) + a0;
            const unsigned lb = smem_u32(land0 + (size_t)slot * slotcap * DP + 4 * lig);
            if (!(p.ws_debug & 1)) {
                int r = grp;
                for (; r + 7 * GPW < cnt; r += 8 * GPW) {
                    int src[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) src[k] = si[r + k * GPW];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lb + (unsigned)(r + k * GPW) * (DP * 4)), "l"(xl + (size_t)src[k] * DP));
                }
                for (; r < cnt; r += GPW) {
                    const int src = si[r];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lb + (unsigned)r * (DP * 4)), "l"(xl + (size_t)src * DP));
                }
            }
            cp_async_mbar_arrive_u32(landed_u32 + 8 * slot);
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bar_colfree[q3]); mbar_arrive(&bar_rowfree[q8]); }
        }
        cp_async_wait_group<0>();
    } else if (warp >= 12) {
        // ============================================= SUM WARPS ===============================================
        // warp w drains sub-tile w % 4 of the staging tiles of parity w / 4 the moment it has landed: segment sums in stored order (8 lanes per row,
        // packed add.f32x2, deterministic) -> 16 aggregate rows (2 KB) in the aggregate FIFO -> the 20 KB slot goes straight back
        // to the issue warp.  The FIFO is 16 sub-tiles deep, so a busy compute quad never holds up the ring.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_SUM));
        const int c = (warp - 12) & 3;
        const int grpw = lane / LPN, lig = lane % LPN;
        const int slot_shift = nslot == 8 ? 3 : 4;
        const uint64_t stream_pol = l2_policy_evict_first();
        constexpr int SWZ = LPN - 1;
        for (int s = (warp - 12) >> 2; s < ntl; s += 2) {
            const int q8 = s & (TC_ROWQ - 1);
            const int j = WS_NSUB * s + c, slot = j & (nslot - 1), phase = (j >> slot_shift) & 1;
            const int e = j & (TC_AGGQ - 1);
            const long long n0 = (t0 + s) * TN;
            const int* srow = srow0 + q8 * 68;
            mbar_wait<GNN_TC_SLEEP>(&bar_rows[q8], (s / TC_ROWQ) & 1);
            mbar_wait<GNN_TC_SLEEP>(&bar_landed[slot], phase);
            if (j >= TC_AGGQ) mbar_wait<GNN_TC_SLEEP>(&bar_aggfree[e], ((j / TC_AGGQ) - 1) & 1);   // the compute warp has read this FIFO entry's previous rows
            const int ebase = srow[0];
            const int a0 = srow[WS_SUB * c] - ebase;
            const int cnt = max(0, min(min(srow[WS_SUB * c + WS_SUB] - ebase, capc) - a0, slotcap));
            const float* lb = land0 + ((size_t)slot * slotcap - a0) * DP + 4 * lig;
            float* out = agg0 + (size_t)e * WS_SUB * DP;
            if (!(p.ws_debug & 2)) {
                const int i0 = WS_SUB * c + grpw * NPG;          // first node (of the staging tile) of my lane group
                const int rlim = a0 + cnt;                        // rows of the sub-tile that landed in the slot
                // one node finished: arcs that did not get ring rows (rare, very dense sub-tiles) by direct loads, weight, store
                auto finish = [&](int u, float4 acc, int rb1) {
                    const int il = grpw * NPG + u, i = i0 + u;
                    for (int r = max(srow[i] - ebase, rlim); r < rb1; ++r) {
                        const int sidx = __ldg(p.col + ebase + r);
                        acc = add4(acc, ldg4(p.x_in + (size_t)sidx * DP + 4 * lig));
                    }
                    const float sc = sscale0[q8 * TN + i];
                    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
                    if (p.agg_save && n0 + i < p.N) st4_hint(p.agg_save + (size_t)(n0 + i) * DP + 4 * lig, acc, stream_pol);
                    st4(out + il * DP + 4 * (lig ^ (il & SWZ)), acc);
                };
                // regular sub-tile (every one of its 16 nodes has the same number of rows, all of them in the slot -- k-NN graphs, the
                // benchmark graph): the NPG nodes of my lane group side by side, software-pipelined in registers (the loads of step
                // e + 1 are issued before the adds of step e: a shared-memory load takes ~200 cycles next to the landing copies)
                const int r_first = srow[i0] - ebase;
                const int d0 = srow[i0 + 1] - ebase - r_first;
                bool regular = srow[i0 + NPG] - ebase <= rlim;
#pragma unroll
                for (int u = 1; u < NPG; ++u) regular &= (srow[i0 + u + 1] - srow[i0 + u]) == d0;
                const int d_lane0 = __shfl_sync(0xffffffffu, d0, 0);      // (outside the && below: every lane must execute the shuffle)
                if (__all_sync(0xffffffffu, regular && d0 == d_lane0)) {
                    const unsigned a_first = smem_u32(lb) + (unsigned)r_first * (DP * 4);      // node u starts u * d0 rows further
                    const unsigned node_step = (unsigned)d0 * (DP * 4);
                    float4 acc[NPG], va[NPG], vb[NPG];
#pragma unroll
                    for (int u = 0; u < NPG; ++u) { acc[u] = make_float4(0.f, 0.f, 0.f, 0.f); va[u] = acc[u]; }
                    if (d0 > 0) {
#pragma unroll
                        for (int u = 0; u < NPG; ++u) va[u] = lds4(a_first + u * node_step);
                    }
#pragma unroll 1
                    for (int e = 0; e < d0; e += 2) {
                        if (e + 1 < d0) {
#pragma unroll
                            for (int u = 0; u < NPG; ++u) vb[u] = lds4(a_first + u * node_step + (e + 1) * (DP * 4));
                        }
#pragma unroll
                        for (int u = 0; u < NPG; ++u) acc[u] = add4_x2(acc[u], va[u]);
                        if (e + 2 < d0) {
#pragma unroll
                            for (int u = 0; u < NPG; ++u) va[u] = lds4(a_first + u * node_step + (e + 2) * (DP * 4));
                        }
                        if (e + 1 < d0) {
#pragma unroll
                            for (int u = 0; u < NPG; ++u) acc[u] = add4_x2(acc[u], vb[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NPG; ++u) finish(u, acc[u], 0);
                } else {
                    // ragged sub-tile: node after node, rows in stored order, 4 loads in flight
#pragma unroll 1
                    for (int u = 0; u < NPG; ++u) {
                        const int r0 = srow[i0 + u] - ebase, r1 = srow[i0 + u + 1] - ebase;
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        const int rl = min(r1, rlim);
#pragma unroll 4
                        for (int r = r0; r < rl; ++r) acc = add4_x2(acc, ld4(lb + (ptrdiff_t)r * DP));
                        finish(u, acc, r1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_free[slot]);
                mbar_arrive(&bar_rowfree[q8]);
                mbar_arrive(&bar_aggfull[e]);
            }
        }
    } else {
        // =========================================== COMPUTE WARPS =============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_COMPUTE));
        const int quad = warp >> 2, q = warp & 3;             // q = TMEM lane quarter = nodes 32 q .. 32 q + 31 of the tile
        const int act = net.act[0], D = net.D;
        const bool affine = !p.bn_train;
        bool any_moving = false;
        double* bn_mine = bn_acc + (size_t)warp * 2 * DP;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;  // TMEM address: lane in bits [31:16], column in [15:0]
        const uint32_t d_acc = tmem_base + quad * stage_cols + lane_base, a_hi = d_acc + DP, a_lo = a_hi + KA;
        constexpr int SWZ = LPN - 1;                          // pieces per row - 1

        // own state row + constant row of MY node of a tile: 32-byte sectors straight from global memory, requested ONE TILE AHEAD
        // (in flight during the staging / MMA / epilogue of the current tile -- nothing waits on their latency)
        float xn[DP], cn[16];
        auto prefetch_own = [&](int t) {
            const int s = 2 * t + (q >> 1);
            const long long node = (t0 + s) * TN + 32 * (q & 1) + lane;
            const bool valid = t < ntl2 && s < ntl && node < p.N;
#pragma unroll
            for (int i = 0; i < DP / 8; ++i) {
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) ldg8(p.x_in + (size_t)(p.row_offset + node) * DP + 8 * i, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) xn[8 * i + e] = v[e];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = (valid && 4 * i < CP) ? ldg4(p.cst + (size_t)node * CP + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                cn[4 * i] = v.x; cn[4 * i + 1] = v.y; cn[4 * i + 2] = v.z; cn[4 * i + 3] = v.w;
            }
        };
        prefetch_own(quad);

        for (int t = quad; t < ntl2; t += TC_QUADS) {
            const int s = 2 * t + (q >> 1);                   // my staging tile; my sub-tiles c0, c0 + 1
            const int c0 = 2 * (q & 1);
            const bool have = s < ntl;                        // the last 128-node tile may end after its first staging tile
            const long long n0 = (t0 + s) * TN;               // first node of my staging tile
            const long long node = n0 + 16 * c0 + lane;       // MY node (thread = node from the operand staging on)
            const bool valid = have && node < p.N;

            // 2. thread = node: hi / lo split of my own state row and constant row -> tensor memory (the quad's previous tile is
            //    complete: its D_FULL was awaited by this very warp before its epilogue); the MMA warp starts on them at once
            uint32_t hi[8], lo[8];
            auto split8 = [&](const float* v) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { hi[e] = __float_as_uint(v[e]) & 0xffffe000u; lo[e] = __float_as_uint(v[e] - __uint_as_float(hi[e])); }
            };
#pragma unroll
            for (int i = 0; i < KX; ++i) {                 // own state: columns [0, DP)
                split8(xn + 8 * i);
                tmem_st8(a_hi + 8 * i, hi); tmem_st8(a_lo + 8 * i, lo);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)                    // constant row: columns [2 DP, 2 DP + 8 CS)
                if (i < CS) {
                    split8(cn + 8 * i);
                    tmem_st8(a_hi + 2 * DP + 8 * i, hi); tmem_st8(a_lo + 2 * DP + 8 * i, lo);
                }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_xfull[quad]);
            prefetch_own(t + TC_QUADS);        // next tile's rows: in flight during the rest of this tile
            uint32_t need = 0u;                // peers that gather from MY node (direct peer stores)
            if (PEERS && p.n_peers > 1 && valid)
                need = (p.peer_mask ? __ldg(p.peer_mask + node) : 0xffffffffu) & ~(1u << p.rank) & ((1u << p.n_peers) - 1u);

            // 3. my 32 aggregate rows = FIFO entries of sub-tiles j0, j0 + 1 (written by sum warps c0, c0 + 1) -> tensor memory
            const int j0 = WS_NSUB * s + c0, e0 = j0 & (TC_AGGQ - 1);
            if (have) {
                mbar_wait<GNN_TC_SLEEP>(&bar_aggfull[e0], (j0 / TC_AGGQ) & 1);
                mbar_wait<GNN_TC_SLEEP>(&bar_aggfull[e0 + 1], (j0 / TC_AGGQ) & 1);
            }
#pragma unroll
            for (int i = 0; i < KX; ++i) {                 // aggregated state: columns [DP, 2 DP)
                float v[8];
                const float* arow = agg0 + ((size_t)e0 * WS_SUB + lane) * DP;      // entries e0, e0 + 1 are adjacent: row `lane` of the pair
                const float4 g0 = ld4(arow + 4 * ((2 * i) ^ (lane & SWZ))), g1 = ld4(arow + 4 * ((2 * i + 1) ^ (lane & SWZ)));
                v[0] = g0.x; v[1] = g0.y; v[2] = g0.z; v[3] = g0.w; v[4] = g1.x; v[5] = g1.y; v[6] = g1.z; v[7] = g1.w;
                if (!have) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = 0.f;
                }
                split8(v);
                tmem_st8(a_hi + DP + 8 * i, hi); tmem_st8(a_lo + DP + 8 * i, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (have) { mbar_arrive(&bar_aggfree[e0]); mbar_arrive(&bar_aggfree[e0 + 1]); }     // FIFO entries back to the sum warps
                mbar_arrive(&bar_afull[quad]);
            }

            // 5. epilogue: my node's DP accumulators -> bias / activation / affine -> store + convergence test
            mbar_wait<GNN_TC_SLEEP>(&bar_dfull[quad], (t >> 1) & 1);
            tc_fence_after();
            float y[DP];
#pragma unroll
            for (int i = 0; i < DP / 8; ++i) {
                uint32_t v[8];
                tmem_ld8(d_acc + 8 * i, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[8 * i + e] = __uint_as_float(v[e]);
            }
            tmem_wait_ld();
            tc_fence_before();       // orders these loads before the next tile's tcgen05.st / mma (through A_FULL)
            auto epilogue = [&](auto act_c) {
                constexpr int ACT = decltype(act_c)::value;
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) y[jj] = act_apply(ACT, y[jj] + sBias[jj]);
                if (affine) {
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj) y[jj] = fmaf(sAff[jj], y[jj], sAff[DP + jj]);
                }
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) if (jj >= D) y[jj] = 0.f;       // padding columns stay zero
            };
            switch (act) {
                case GNN_ACT_RELU: epilogue(std::integral_constant<int, GNN_ACT_RELU>{}); break;
                case GNN_ACT_TANH: epilogue(std::integral_constant<int, GNN_ACT_TANH>{}); break;
                case GNN_ACT_SIGMOID: epilogue(std::integral_constant<int, GNN_ACT_SIGMOID>{}); break;
                case GNN_ACT_SELU: epilogue(std::integral_constant<int, GNN_ACT_SELU>{}); break;
                case GNN_ACT_ELU: epilogue(std::integral_constant<int, GNN_ACT_ELU>{}); break;
                case GNN_ACT_SOFTPLUS: epilogue(std::integral_constant<int, GNN_ACT_SOFTPLUS>{}); break;
                default: epilogue(std::integral_constant<int, GNN_ACT_LINEAR>{}); break;
            }
            if (valid) {
                float* orow = p.x_out + (size_t)(p.row_offset + node) * DP;
#pragma unroll
                for (int i = 0; i < DP / 8; ++i) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = y[8 * i + e];
                    stg8(orow + 8 * i, v);
                }
                while (PEERS && need) {         // my row into every peer that gathers from it: plain remote stores, nothing waits on them
                    const int r = __ffs(need) - 1;
                    need &= need - 1u;
                    float* prow = p.peer_out[r] + (size_t)(p.row_offset + node) * DP;
#pragma unroll
                    for (int i = 0; i < DP / 8; ++i) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = y[8 * i + e];
                        stg8(prow + 8 * i, v);
                    }
                }
            }
            if (p.bn_train) {
                // column sums over my warp's 32 nodes (thread = node): butterfly transpose-reduce, 31 + 31 shuffles; lane j ends
                // with the totals of column j, accumulated in fp64 in the warp's private row
                float s1[DP], s2[DP];
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) { s1[jj] = valid ? y[jj] : 0.f; s2[jj] = s1[jj] * s1[jj]; }
                warp_column_sums<DP>(s1, lane);
                warp_column_sums<DP>(s2, lane);
                if (lane < DP) {
                    bn_mine[lane] += (double)s1[0];
                    bn_mine[DP + lane] += (double)s2[0];
                }
            } else {
                // convergence test against the OLD state of my node: hi + lo = the exact fp32 row, still in the A operand (tensor memory)
                float d2 = 0.f, o2 = 0.f;
#pragma unroll
                for (int i = 0; i < DP / 8; ++i) {
                    uint32_t vh[8], vl[8];
                    tmem_ld8(a_hi + 8 * i, vh);
                    tmem_ld8(a_lo + 8 * i, vl);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float xo = __uint_as_float(vh[e]) + __uint_as_float(vl[e]);
                        const float dx = y[8 * i + e] - xo;
                        d2 = fmaf(dx, dx, d2);
                        o2 = fmaf(xo, xo, o2);
                    }
                }
                tc_fence_before();     // these loads precede the next tile's tcgen05.st into the same columns
                any_moving |= valid && (sqrtf(d2) > p.thr * sqrtf(o2));
            }
        }

        if (p.bn_train) {
            named_bar_sync(GNN_BAR_MLP_ALL, TC_COMPUTE);
            if (tid < 2 * DP) {
                double sum = 0.;
                for (int w = 0; w < TC_COMPUTE / 32; ++w) sum += bn_acc[(size_t)w * 2 * DP + tid];
                p.bn_partial[(size_t)blockIdx.x * 2 * DP + tid] = sum;
            }
        } else {
            if (p.go_next && __any_sync(0xffffffffu, any_moving) && lane == 0) s_flag = 1;
            if (PEERS && p.n_peers > 1) __threadfence_system();      // my peer stores before the arrival mark of iter_end
            named_bar_sync(GNN_BAR_MLP_ALL, TC_COMPUTE);
            if (tid == 0) iter_end(p, s_flag);
        }
        tc_fence_before();
    }
    // tensor memory goes back once every role is done with it
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TC_TMEM_COLS); }
}

}  // namespace gnn
