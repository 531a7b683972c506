// state_fwd_ws.cuh -- warp-specialised, software-pipelined variant of the fused forward iteration kernel.
//
// Why: the gather of a uniform-random graph is bound by how many 128-byte rows an SM keeps in flight.  A gather-only
// microbenchmark on B200 (scripts/gather_microbench.cu) needs >= 1000 rows in flight per SM to reach the L2/HBM limit;
// the symmetric kernel (state_fwd.cuh) holds its rows in registers and cannot have that many in flight next to the MLP.
// Here the rows land in shared memory through cp.async (no registers, deep queue) and FOUR roles run concurrently in
// one persistent CTA of 768 threads per SM (21 working warps) (register budgets re-split per role with setmaxnreg).  Work unit of the
// gather: a SUB-TILE of 16 nodes (= one m16 mma block); 4 sub-tiles = one 64-node tile.  The landing zone is a ring of
// 8 or 16 slots, one sub-tile each: a slot is in flight from the moment its copies are issued until its segment sums are
// done (a ring of two whole-tile stages spends half of its life waiting to be consumed).
//
//   loader warp   (1): stages what the other roles index with, two tiles ahead, with BULK ASYNC COPIES (cp.async.bulk,
//                      one elected thread, no per-element instructions): the arc sources of a tile (mbarrier COLS) and its
//                      row pointers / per-node scales / per-arc weights (mbarrier ROWS).  A CTA owns a CONTIGUOUS range of
//                      tiles, so every copy is one aligned contiguous block.
//   issue warps   (4): warp w lands sub-tile w of every tile; never waits for data.  Per sub-tile: wait for its slot
//                      (mbarrier FREE), then per 4 source rows ONE index load from shared memory, ONE address multiply-add
//                      and ONE 16-byte cp.async per lane; completion of all of them arrives on the slot's mbarrier LANDED
//                      (cp.async.mbarrier.arrive.noinc).  ~1 warp instruction per landed row (round 1: ~5).
//   consume warps (8): two groups of 4 (even / odd tiles); warp c owns sub-tile c.  wait LANDED -> segment sums out of
//                      shared memory in stored order (deterministic, no atomics, packed add.f32x2) into rows 16c..16c+15
//                      of aggregate tile g -> mbarrier FULL[g][c] for MLP warp c of group g, mbarrier FREE for the slot
//   MLP warps     (8): two groups of 4 (even / odd tiles), paired 1:1 with the consume warps; warp c owns rows
//                      16c..16c+15: own state / constant rows straight from global memory as mma A fragments, wait FULL
//                      -> Dense layer on the tensor cores (mma.sync m16n8k8, 3xTF32 = fp32-accurate) -> mbarrier EMPTY
//                      -> bias / activation / affine -> store of the new state + convergence test (+ BatchNormalization
//                      batch statistics when training).  No block-wide barrier in the loop.
// Every ring slot and every tile row has exactly one producer warp set and one consumer warp: nobody can observe an
// mbarrier two phases late (a parity wait cannot tell phase k from phase k+2).
//
// Used for single-Dense-layer state nets (what the reference builds by default) with padded state width 16..32 and no
// active dropout; every other case runs the symmetric kernel.
#pragma once
#include "state_fwd.cuh"

namespace gnn {

// named barriers (0 is __syncthreads): the MLP warps once at the end of the kernel
#define GNN_BAR_MLP_ALL 2

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16_if(bool pred, float* smem_dst, const float* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(gmem_src), "r"((int)pred));
}
// packed fp32 pairs (sm_100 add.f32x2 / fma.rn.f32x2): same IEEE result per element, half the instructions
__device__ __forceinline__ float4 add4_x2(float4 a, float4 b) {
    float4 r;
    asm("{\n\t.reg .b64 a0, a1, b0, b1, r0, r1;\n\t"
        "mov.b64 a0, {%4, %5};\n\tmov.b64 a1, {%6, %7};\n\tmov.b64 b0, {%8, %9};\n\tmov.b64 b1, {%10, %11};\n\t"
        "add.rn.f32x2 r0, a0, b0;\n\tadd.rn.f32x2 r1, a1, b1;\n\t"
        "mov.b64 {%0, %1}, r0;\n\tmov.b64 {%2, %3}, r1;\n\t}"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
    return r;
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col): the legacy warp-level tensor-core path of sm_100a
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo with hi exactly representable in tf32 (low 13 mantissa bits cleared): the 3-pass split that keeps the
// product at fp32 accuracy (hi*hi + hi*lo + lo*hi; the dropped lo*lo term is ~2^-20 relative)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

constexpr int WS_TN = 64;          // nodes per tile
constexpr int WS_SUB = 16;         // nodes per sub-tile (one m16 block, one consume warp, one MLP warp)
constexpr int WS_NSUB = WS_TN / WS_SUB;
constexpr int WS_MLP = 128;        // threads of ONE MLP group (4 warps)
constexpr int WS_MLP_GROUPS = 2;   // group g takes the tiles it = g, g+2, ... (tile buffer g)
constexpr int WS_MLP_ALL = WS_MLP * WS_MLP_GROUPS;   // warps 0-7
constexpr int WS_CONS = 256;       // consume threads (warps 8-15): group g (4 warps) takes the tiles it = g, g+2, ...; warp c sub-tile c
constexpr int WS_ISSUE = 128;      // issue threads (warps 16-19): warp w lands sub-tile w of every tile
constexpr int WS_LOADER = 32;      // loader warp (warp 20): bulk copies of arc sources / row pointers / scales
constexpr int WS_IDLE = 96;        // warps 21-23 give their registers back and exit: the register file is split over the 4 SM
                                   // sub-partitions (16K each), so 21 warps get no more registers per thread at launch than 24 do
constexpr int WS_THREADS = WS_MLP_ALL + WS_CONS + WS_ISSUE + WS_LOADER + WS_IDLE;   // 768
// registers per thread after the roles split (setmaxnreg): 768 x 80 at launch -> MLP 128, consume 64, issue 48, loader 40, idle 24.
// Per sub-partition: 2 MLP + 2 consume + 1 issue + 1 loader / idle warp = 2*128 + 2*64 + 48 + 40 = 472 <= 512 registers per lane.
constexpr int WS_REGS_LAUNCH = 80, WS_REGS_MLP = 128, WS_REGS_CONS = 64, WS_REGS_ISSUE = 48, WS_REGS_LOADER = 40, WS_REGS_IDLE = 24;
static_assert(WS_THREADS * WS_REGS_LAUNCH <= 65536, "registers of one SM at launch");
static_assert(WS_MLP_ALL * (WS_REGS_MLP - WS_REGS_LAUNCH) <= WS_ISSUE * (WS_REGS_LAUNCH - WS_REGS_ISSUE) + WS_CONS * (WS_REGS_LAUNCH - WS_REGS_CONS) +
                                                                  WS_LOADER * (WS_REGS_LAUNCH - WS_REGS_LOADER) + WS_IDLE * (WS_REGS_LAUNCH - WS_REGS_IDLE),
              "setmaxnreg.inc only draws on what this CTA's own warps released (spare registers of the SM do not count)");
constexpr int WS_SLOTS = 16;       // sub-tiles in flight at most (mbarrier slots of the ring)
constexpr int WS_ROWQ = 8;         // row-pointer / scale buffers (tiles): staged 2 tiles ahead of the issue, consumed <= 4 tiles behind it
constexpr int WS_COLQ = 4;         // arc-source buffers (tiles): staged up to 3 tiles ahead of the issue
constexpr int WS_COLPAD = 8;       // slack of each arc-source / weight buffer: the copy starts at the 16-byte boundary below the first arc

#ifndef GNN_WS_CONS_UNROLL
#define GNN_WS_CONS_UNROLL 4
#endif
#ifndef GNN_WS_SLEEP_MLP
#define GNN_WS_SLEEP_MLP 64
#endif
#ifndef GNN_WS_SLEEP_CONS
#define GNN_WS_SLEEP_CONS 32
#endif
// mbarrier (shared memory, CTA scope)
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrives (without touching the pending count) once every cp.async this thread has issued so far has completed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// the same operations on a precomputed shared-memory address (saves the generic -> shared conversion per call)
__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void cp_async_mbar_arrive_u32(unsigned addr) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(unsigned addr, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// SLEEP_NS > 0: back off between polls.  A spinning warp polls about once per 30 cycles and takes issue slots from the warps
// it is waiting for; the roles with slack (MLP, consume) sleep, the issue warps (the critical role) spin.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    }
}

// aggregate tile row stride (floats): rows g and g+1 of a quarter-warp's 128-bit fragment loads fall in different bank halves
static inline int ws_tile_stride(int DP) { return DP % 32 == 0 ? DP + 16 : DP; }
// k-steps (8 inputs each) of the Dense layer: own state, constant row (padded to 8), aggregated state
static inline int ws_ksteps(const NetLayout& lay) { return 2 * (lay.DP / 8) + (lay.CP + 7) / 8; }
// weights in shared memory: B fragments {hi(b0), hi(b1), lo(b0), lo(b1)} per (k-step, n-tile, lane), bias, affine a / c
static inline size_t ws_weight_floats(const NetLayout& lay) { return (size_t)ws_ksteps(lay) * (lay.DP / 8) * 128 + 3 * (size_t)lay.DP; }

// shared-memory footprint (bytes): `ring` landing rows (slots x rows per slot), arc-index capacity `capc` per tile
static inline size_t ws_smem_bytes(const NetLayout& lay, int ring, int capc, bool has_val) {
    size_t fl = ws_weight_floats(lay) + WS_MLP_GROUPS * (size_t)WS_TN * ws_tile_stride(lay.DP) + (size_t)ring * lay.DP + WS_ROWQ * 68 + WS_ROWQ * WS_TN +
                WS_COLQ * (size_t)(capc + WS_COLPAD) + (has_val ? WS_ROWQ * (size_t)(capc + WS_COLPAD) : 0);   // weights, tiles x2, ring, row pointers / scales, arc sources, arc weights
    return fl * 4 + 128;      // + alignment slack of the dynamic base
}

// bulk asynchronous copy global -> shared (UBLKCP): one thread, one instruction, any multiple of 16 bytes; both addresses
// 16-byte aligned.  Completion is counted in bytes on the mbarrier (expect_tx).
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, int bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// PEERS: node-range partition over several GPUs (rows other ranks gather from are also stored into their memory); a template
// parameter so that the single-GPU kernel carries none of it (its MLP warps sit at the 128-register cap)
template <int DP, bool HAS_VAL, bool PEERS>
__global__ void __launch_bounds__(WS_THREADS, 1) state_iter_ws_kernel(const IterParams p) {
    constexpr int TN = WS_TN, LPN = DP / 4;
    constexpr int GPW = 32 / LPN;            // lane groups per warp = source rows per warp-wide cp.async
    constexpr int NPG = WS_SUB / GPW;        // nodes per lane group of a consume warp
    static_assert(GPW * NPG == WS_SUB, "lane mapping");
    static_assert(DP % 8 == 0 && TN == 64 && WS_NSUB == 4, "4 MLP warps x 16 nodes, DP / 8 accumulator fragments each");
    static_assert(WS_ISSUE == 32 * WS_NSUB, "one issue warp per sub-tile of a tile");

    if (!iter_begin(p)) return;

    const NetLayout& net = p.net;
    const int tid = threadIdx.x;
    const int CP = net.CP, capc = p.scol_cap, capb = capc + WS_COLPAD;
    const int nslot = p.ring_slots, slotcap = p.slot_rows;   // landing ring: nslot slots of slotcap rows, sub-tile j -> slot j % nslot

    // dynamic shared memory is only 16-byte aligned behind the static variables: the 128-byte rows of the ring want whole lines
    extern __shared__ __align__(16) float smem_raw[];
    float* smem = smem_raw + (((128u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u) >> 2);
    constexpr int NT8 = DP / 8, NQ = DP / 16;           // n-tiles of 8 outputs; 16-column units of a state row
    constexpr int SAG = DP % 32 == 0 ? DP + 16 : DP;    // aggregate tile row stride (ws_tile_stride)
    const int CS = (CP + 7) / 8;                        // k-steps of the constant row
    const int KSTEPS = 4 * NQ + CS;
    float4* sW4 = reinterpret_cast<float4*>(smem);      // [KSTEPS][NT8][32 lanes]: B fragments {hi b0, hi b1, lo b0, lo b1}
    float* sBias = smem + (size_t)KSTEPS * NT8 * 128;   // [DP]
    float* sAff = sBias + DP;                           // a[DP], c[DP]
    float* tile0 = sAff + 2 * DP;                       // [WS_MLP_GROUPS][TN][SAG]: aggregated states
    float* land0 = tile0 + WS_MLP_GROUPS * TN * SAG;    // [nslot][slotcap][DP]
    int* srow0 = reinterpret_cast<int*>(land0 + (size_t)nslot * slotcap * DP);   // [WS_ROWQ][68]
    float* sscale0 = reinterpret_cast<float*>(srow0 + WS_ROWQ * 68);     // [WS_ROWQ][TN]
    int* scol0 = reinterpret_cast<int*>(sscale0 + WS_ROWQ * TN);         // [WS_COLQ][capb]
    float* sval0 = reinterpret_cast<float*>(scol0 + WS_COLQ * capb);     // [WS_ROWQ][capb] (HAS_VAL; read by the consume warps)
    __shared__ int s_flag;
    __shared__ __align__(8) uint64_t bar_landed[WS_SLOTS], bar_free[WS_SLOTS], bar_cols[WS_COLQ], bar_colfree[WS_COLQ], bar_rows[WS_ROWQ], bar_rowfree[WS_ROWQ],
        bar_full[WS_MLP_GROUPS][WS_NSUB], bar_empty[WS_MLP_GROUPS][WS_NSUB];
    __shared__ double bn_acc[4 * WS_MLP_GROUPS][2][DP];   // BatchNormalization batch statistics, one private row per MLP warp

    // The mma K and N indices are ours to permute as long as A, B and C agree.  Both are laid out so that a lane's fragment
    // elements are CONTIGUOUS in memory (128-bit loads / stores, whole 32-byte sectors per 4 lanes):
    //   K: k-steps 2q, 2q+1 of a 16-column unit q of a state row: lane ft holds columns 16q + 4ft + {0,1} (step 2q: k = ft,
    //      ft+4) and 16q + 4ft + {2,3} (step 2q+1); a k-step s of the constant row: columns 8s + 2ft + {0,1}.
    //      Order of the k-steps: own state (2 NQ), constant row (CS), aggregated state (2 NQ).
    //   N: n-tile nt = 2q + h, accumulator element j of lane ft <-> output column 16q + 4ft + 2h + j,
    //      i.e. a lane's 4 NT8 / 2 accumulators of a row are the same columns as its state-row fragments.
    // B fragment of (k-step s, n-tile nt, lane (g, ft)): b0 = W[krow(s, ft)][ncol(nt, g)], b1 = W[krow(s, ft + 4)][ncol(nt, g)],
    // stored pre-split for 3xTF32: hi = tf32 part, lo = exact remainder.
    for (int i = tid; i < KSTEPS * NT8 * 32; i += WS_THREADS) {
        const int lane = i & 31, nt = (i >> 5) % NT8, st = (i >> 5) / NT8;
        const int ft = lane & 3, g = lane >> 2;
        const int n = 16 * (nt >> 1) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);
        float w[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int row;   // input row of the packed kernel [state(DP) | agg(DP) | cst(CP)]
            if (st < 2 * NQ) row = 16 * (st >> 1) + 4 * ft + 2 * (st & 1) + h;
            else if (st < 2 * NQ + CS) { const int c = 8 * (st - 2 * NQ) + 2 * ft + h; row = c < CP ? 2 * DP + c : -1; }
            else { const int sl = st - 2 * NQ - CS; row = DP + 16 * (sl >> 1) + 4 * ft + 2 * (sl & 1) + h; }
            w[h] = row >= 0 ? __ldg(p.wpack + net.w_off[0] + row * DP + n) : 0.f;
        }
        const float h0 = __uint_as_float(__float_as_uint(w[0]) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(w[1]) & 0xffffe000u);
        sW4[i] = make_float4(h0, h1, w[0] - h0, w[1] - h1);
    }
    for (int i = tid; i < DP; i += WS_THREADS) {
        sBias[i] = __ldg(p.wpack + net.b_off[0] + i);
        sAff[i] = __ldg(p.wpack + net.aff_off + i);
        sAff[DP + i] = __ldg(p.wpack + net.aff_off + DP + i);
    }
    for (int i = tid; i < 4 * WS_MLP_GROUPS * 2 * DP; i += WS_THREADS) (&bn_acc[0][0][0])[i] = 0.;
    if (tid == 0) {
        s_flag = 0;
        for (int i = 0; i < WS_SLOTS; ++i) { mbar_init(&bar_landed[i], 32); mbar_init(&bar_free[i], 1); }     // LANDED: every lane of the slot's issue warp; FREE: its consume warp
        for (int i = 0; i < WS_COLQ; ++i) { mbar_init(&bar_cols[i], 1); mbar_init(&bar_colfree[i], WS_NSUB); }   // COLS: the loader (+ bytes); COLFREE: the 4 issue warps
        for (int i = 0; i < WS_ROWQ; ++i) { mbar_init(&bar_rows[i], 1); mbar_init(&bar_rowfree[i], 2 * WS_NSUB); }   // ROWFREE: 4 issue + 4 consume warps
        for (int i = 0; i < WS_MLP_GROUPS * WS_NSUB; ++i) { mbar_init(&bar_full[0][0] + i, 1); mbar_init(&bar_empty[0][0] + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // visible to the async proxy (bulk copies)
    }
    __syncthreads();

    // this CTA's tiles: a CONTIGUOUS range (its row pointers, arc sources and weights are contiguous in memory)
    const long long ntiles = (p.N + TN - 1) / TN;
    const long long t0 = ntiles * blockIdx.x / gridDim.x;
    const int ntl = (int)(ntiles * (blockIdx.x + 1) / gridDim.x - t0);

    // sub-tile c of a tile lands the arcs [a0, a0 + cnt) (tile-relative) in its ring slot; arcs beyond the slot (or beyond
    // the staged arc indices) are read directly by the consume warp
    auto sub_range = [&](const int* srow, int c, int& a0, int& cnt) {
        const int ebase = srow[0];
        a0 = srow[WS_SUB * c] - ebase;
        const int a1 = min(srow[WS_SUB * c + WS_SUB] - ebase, capc);
        cnt = max(0, min(a1 - a0, slotcap));
    };

    if (tid >= WS_MLP_ALL + WS_CONS + WS_ISSUE + WS_LOADER) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REGS_IDLE));   // idle warps: registers back to the pool, done
    } else if (tid >= WS_MLP_ALL + WS_CONS + WS_ISSUE) {
        // ============================================ LOADER WARP ==============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REGS_LOADER));
        const int lane = tid & 31;
        const long long E = p.E;
        // arc range of every tile = rowptr at the tile boundaries: 32 boundaries per load (one per lane), fetched one batch ahead,
        // so that no global-load latency sits between two tiles of the staging loop
        auto fetch_bounds = [&](int base) {
            return base + lane <= ntl ? __ldg(p.rowptr + min((t0 + base + lane) * TN, p.N)) : 0;
        };
        int bcur = fetch_bounds(0), bnext = fetch_bounds(32);
        for (int s = 0; s < ntl; ++s) {
            const long long n0 = (t0 + s) * TN;
            const int q8 = s & (WS_ROWQ - 1), q3 = s % WS_COLQ;
            if ((s & 31) == 0 && s > 0) { bcur = bnext; bnext = fetch_bounds(s + 32); }
            const int e0 = __shfl_sync(0xffffffffu, bcur, s & 31);
            const int e1 = (s & 31) == 31 ? __shfl_sync(0xffffffffu, bnext, 0) : __shfl_sync(0xffffffffu, bcur, (s + 1) & 31);   // (warp-uniform)
            const int e0a = e0 & ~3;                                            // 16-byte boundary at or below the first arc
            const int narc = min(e1 - e0a, capb - 4), narc4 = (narc + 3) & ~3;  // staged entries (a multiple of 16 bytes)
            const bool arcs_safe = (long long)e0a + narc4 <= E;                 // the rounded-up copy stays inside the array
            const bool rows_safe = n0 + 68 <= p.N + 1 && n0 + TN <= p.N;
            // ---- row pointers, per-node scales (or per-arc weights): buffer q8, free once tile s - WS_ROWQ has been consumed
            if (s >= WS_ROWQ) mbar_wait<64>(&bar_rowfree[q8], ((s / WS_ROWQ) - 1) & 1);
            int* srow = srow0 + q8 * 68;
            float* ssc = sscale0 + q8 * TN;
            float* sv = sval0 + (size_t)q8 * capb;
            if (rows_safe && (!HAS_VAL || arcs_safe)) {
                if (lane == 0) {
                    mbar_expect_tx(&bar_rows[q8], 68 * 4 + (HAS_VAL ? narc4 * 4 : TN * 4));
                    bulk_copy_g2s(srow, p.rowptr + n0, 68 * 4, &bar_rows[q8]);
                    if (HAS_VAL) { if (narc4 > 0) bulk_copy_g2s(sv, p.val + e0a, narc4 * 4, &bar_rows[q8]); }
                    else bulk_copy_g2s(ssc, p.row_scale + n0, TN * 4, &bar_rows[q8]);
                }
            } else {   // last tile of the graph: element-wise, nothing is read past the end of an array
                for (int i = lane; i <= TN; i += 32) srow[i] = __ldg(p.rowptr + min(n0 + i, p.N));
                if (HAS_VAL) { for (int i = lane; i < narc; i += 32) sv[i] = __ldg(p.val + e0a + i); }
                else for (int i = lane; i < TN; i += 32) ssc[i] = n0 + i < p.N ? __ldg(p.row_scale + n0 + i) : 0.f;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_rows[q8]);
            }
            // ---- arc sources: buffer q3, free once the issue warps are done with tile s - WS_COLQ
            if (s >= WS_COLQ) mbar_wait<64>(&bar_colfree[q3], ((s / WS_COLQ) - 1) & 1);
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
    } else if (tid >= WS_MLP_ALL + WS_CONS) {
        // ============================================ ISSUE WARPS ==============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REGS_ISSUE));
        const int c = (tid - WS_MLP_ALL - WS_CONS) >> 5;        // my sub-tile of every tile
        const int lane = tid & 31, grp = lane / LPN, lig = lane % LPN;
        const unsigned landed_u32 = smem_u32(bar_landed), free_u32 = smem_u32(bar_free);
        const float* xl = p.x_in + 4 * lig;
        const int slot_shift = nslot == 8 ? 3 : 4;
        for (int s = 0; s < ntl; ++s) {
            const int q8 = s & (WS_ROWQ - 1), q3 = s % WS_COLQ;
            mbar_wait(&bar_rows[q8], (s / WS_ROWQ) & 1);           // row pointers of tile s
            mbar_wait(&bar_cols[q3], (s / WS_COLQ) & 1);           // arc sources of tile s
            const int* srow = srow0 + q8 * 68;
            const int eb = srow[0];
            const int a0 = srow[WS_SUB * c] - eb;
            const int cnt = max(0, min(min(srow[WS_SUB * c + WS_SUB] - eb, capc) - a0, slotcap));
            const int j = WS_NSUB * s + c, slot = j & (nslot - 1);
            // a slot is only ever used by ONE issue warp and ONE consume warp (nslot is a multiple of 4): nobody waits on
            // its barriers more than one phase behind
            if (j >= nslot) mbar_wait_u32(free_u32 + 8 * slot, ((j >> slot_shift) - 1) & 1);   // the slot's previous sub-tile has been consumed
            // every source row -> ring (asynchronous, no registers): one row per lane group and step
            const int* si = scol0 + (size_t)q3 * capb + (eb & 3) + a0;
            const unsigned lb = smem_u32(land0 + (size_t)slot * slotcap * DP + 4 * lig);
            if (!(p.ws_debug & 1)) {
                // batches of 8 steps: 8 independent index loads, 8 address multiply-adds, 8 copies (no dependent pair back to back)
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
    } else if (tid >= WS_MLP_ALL) {
        // =========================================== CONSUME WARPS =============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REGS_CONS));
        const int cwarp = (tid - WS_MLP_ALL) >> 5;
        const int cw = cwarp & 3, cg = cwarp >> 2;          // sub-tile cw of the tiles it = cg, cg+2, ...: the warp pairs with MLP warp
                                                            // cw of group cg (tile buffer cg) and, nslot being 8 or 16, is the only
                                                            // consumer of its ring slots
        const int lane = tid & 31, grpw = lane / LPN, lig = lane % LPN;
        const uint64_t stream_pol = l2_policy_evict_first();
        const int b = cg;
        const int slot_shift = nslot == 8 ? 3 : 4;
        int round = 0;                                      // it / 2
        for (int it = cg; it < ntl; it += 2, ++round) {
            const long long tile = t0 + it;
            const int j = WS_NSUB * it + cw, slot = j & (nslot - 1), phase = (j >> slot_shift) & 1;
            const int q8 = it & (WS_ROWQ - 1);
            const long long n0 = tile * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            const int* srow = srow0 + q8 * 68;
            mbar_wait<GNN_WS_SLEEP_CONS>(&bar_rows[q8], (it / WS_ROWQ) & 1);   // row pointers / scales (/ weights) of the tile (bulk copies)
            mbar_wait<GNN_WS_SLEEP_CONS>(&bar_landed[slot], phase);            // rows of the sub-tile
            int a0, cnt;
            sub_range(srow, cw, a0, cnt);
            const int start = slot * slotcap;
            if (round > 0) mbar_wait<GNN_WS_SLEEP_CONS>(&bar_empty[b][cw], (round - 1) & 1);   // MLP warp cw of group b is done with tile it - WS_MLP_GROUPS
            const int ebase = srow[0];
            const float* lb = land0 + ((size_t)start - a0) * DP + 4 * lig;   // row of tile-relative arc r: lb + r * DP
            const float* sv = sval0 + (size_t)q8 * capb + (ebase & 3);       // weight of tile-relative arc r (staged from the 16-byte boundary below)
            float* tb = tile0 + (size_t)b * TN * SAG + 4 * lig;
            // segment sums of this lane group's NPG nodes out of the ring, stored order
#pragma unroll 1
            for (int u = 0; u < ((p.ws_debug & 2) ? 0 : NPG); ++u) {
                const int i = WS_SUB * cw + grpw * NPG + u;
                const int r0 = srow[i] - ebase, r1 = srow[i + 1] - ebase;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int r = r0;
                const int rl = min(r1, a0 + cnt);
                constexpr int CONS_UNROLL = GNN_WS_CONS_UNROLL;
#pragma unroll CONS_UNROLL
                for (; r < rl; ++r) {
                    const float4 v = ld4(lb + (ptrdiff_t)r * DP);
                    if (HAS_VAL) acc = fma4(sv[r], v, acc);
                    else acc = add4_x2(acc, v);
                }
                for (; r < r1; ++r) {   // arcs that did not get ring rows: direct loads (rare, very dense sub-tiles)
                    const int s = __ldg(p.col + ebase + r);
                    const float4 v = ldg4(p.x_in + (size_t)s * DP + 4 * lig);
                    if (HAS_VAL) acc = fma4(__ldg(p.val + ebase + r), v, acc);
                    else acc = add4(acc, v);
                }
                if (!HAS_VAL) {
                    const float sc = sscale0[q8 * TN + i];
                    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
                }
                if (p.agg_save && i < nvalid) st4_hint(p.agg_save + (size_t)(n0 + i) * DP + 4 * lig, acc, stream_pol);
                st4(tb + i * SAG, acc);
            }
            __syncwarp();                                  // every lane's reads of the slot / row buffers and writes of the tile rows are done
            if (lane == 0) {
                mbar_arrive(&bar_free[slot]);
                mbar_arrive(&bar_rowfree[q8]);
                mbar_arrive(&bar_full[b][cw]);
            }
        }
    } else {
        // ============================================= MLP WARPS ==============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REGS_MLP));
        // two MLP groups of 4 warps: group g takes the tiles with sequence number it = g, g+2, ... (= tile buffer g).
        // Inside a group warp w owns nodes [16 w, 16 w + 16) of the tile and all DP outputs: NT8 accumulator fragments.
        // The warps are independent of each other: each one pairs with consume warp w through FULL / EMPTY[g][w].
        // Own state rows and constant rows never touch shared memory: a lane reads its A fragments straight from global
        // memory (128-bit, in flight while it waits for the aggregates) and keeps them for the convergence test.
        const int mgroup = tid / WS_MLP, mt = tid % WS_MLP;
        const int mwarp = mt >> 5, lane = mt & 31, fg = lane >> 2, ft = lane & 3;   // fragment coordinates (groupID, thread-in-group)
        const int act = net.act[0], D = net.D;
        const bool affine = !p.bn_train;
        const uint64_t stream_pol = l2_policy_evict_first();
        bool any_moving = false;
        const float* tb = tile0 + (size_t)mgroup * TN * SAG + (size_t)(16 * mwarp + fg) * SAG + 4 * ft;   // my row fg of the tile, + 8 SAG: row fg + 8
        double* bn_mine = &bn_acc[tid >> 5][0][0];   // [2][DP], private to this warp

        // own state rows fg, fg + 8 (columns 16 q + 4 ft .. + 3) and constant rows (columns 8 s + 2 ft, + 1; CS <= 2) of a tile,
        // loaded ONE TILE OF THIS GROUP AHEAD (a whole tile period in flight)
        float4 xnext[2][NQ];
        float2 cnext[2][2];
        uint32_t need_next = 0u;     // partition: peers that gather from my rows fg (bits 0-7) and fg + 8 (bits 8-15), fetched with the rows
        const uint32_t others = PEERS ? (~(1u << p.rank) & ((1u << p.n_peers) - 1u) & 0xffu) : 0u;
        auto load_own = [&](long long tile) {
            const long long n0 = tile * TN;
            if (PEERS) need_next = 0u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long n = n0 + 16 * mwarp + fg + 8 * h;
                const bool valid = n < p.N;
                if (PEERS && valid) need_next |= ((p.peer_mask ? __ldg(p.peer_mask + n) : 0xffu) & others) << (8 * h);
                const float* xr = p.x_in + (size_t)(p.row_offset + n) * DP + 4 * ft;
                const float* cr = p.cst + (size_t)n * CP + 2 * ft;
#pragma unroll
                for (int q = 0; q < NQ; ++q) xnext[h][q] = valid ? ldg4(xr + 16 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2)
                    cnext[h][s2] = (valid && 8 * s2 + 2 * ft < CP) ? __ldg(reinterpret_cast<const float2*>(cr + 8 * s2)) : make_float2(0.f, 0.f);
            }
        };
        if (mgroup < ntl) load_own(t0 + mgroup);

        int round = 0;
        for (int it = mgroup; it < ntl; it += WS_MLP_GROUPS, ++round) {
            const long long tile = t0 + it;
            const long long n0 = tile * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            float4 xcur[2][NQ];
            float2 ccur[2][2];
            const uint32_t need_cur = need_next;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) xcur[h][q] = xnext[h][q];
                ccur[h][0] = cnext[h][0]; ccur[h][1] = cnext[h][1];
            }
            if (it + WS_MLP_GROUPS < ntl) load_own(tile + WS_MLP_GROUPS);

            // Dense layer on the tensor cores: 16 x DP outputs per warp, 3 x TF32 (fp32-accurate)
            float acc[NT8][4];
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            auto mma_step = [&](int st, float a0, float a1, float a2, float a3) {   // rows fg, fg+8 x k = ft, ft+4
                if (p.ws_debug & 4) return;
                uint32_t ahi[4], alo[4];
                split_tf32(a0, ahi[0], alo[0]); split_tf32(a1, ahi[1], alo[1]);
                split_tf32(a2, ahi[2], alo[2]); split_tf32(a3, ahi[3], alo[3]);
                const float4* wf = sW4 + (size_t)st * NT8 * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt) {
                    const float4 w = wf[nt * 32];
                    const uint32_t bhi[2] = {__float_as_uint(w.x), __float_as_uint(w.y)}, blo[2] = {__float_as_uint(w.z), __float_as_uint(w.w)};
                    mma_tf32_16x8x8(acc[nt], alo, bhi);
                    mma_tf32_16x8x8(acc[nt], ahi, blo);
                    mma_tf32_16x8x8(acc[nt], ahi, bhi);
                }
            };
            mbar_wait<GNN_WS_SLEEP_MLP>(&bar_full[mgroup][mwarp], round & 1);        // aggregates of my 16 nodes are in the buffer
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 g0 = ld4(tb + 16 * q), g1 = ld4(tb + 8 * SAG + 16 * q);
                mma_step(2 * NQ + CS + 2 * q, g0.x, g1.x, g0.y, g1.y);
                mma_step(2 * NQ + CS + 2 * q + 1, g0.z, g1.z, g0.w, g1.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[mgroup][mwarp]);   // my rows of the buffer are free again (every lane's loads are complete)
            // own state and constant row
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                mma_step(2 * q, xcur[0][q].x, xcur[1][q].x, xcur[0][q].y, xcur[1][q].y);
                mma_step(2 * q + 1, xcur[0][q].z, xcur[1][q].z, xcur[0][q].w, xcur[1][q].w);
            }
            mma_step(2 * NQ, ccur[0][0].x, ccur[1][0].x, ccur[0][0].y, ccur[1][0].y);
            if (CS > 1) mma_step(2 * NQ + 1, ccur[0][1].x, ccur[1][1].x, ccur[0][1].y, ccur[1][1].y);

            // epilogue straight from the accumulator fragments: bias + activation + affine, one 128-bit store per 16-column
            // unit and row, convergence test against the old state held in xcur, reduced over the 4 lanes of a row.  The
            // activation is a compile-time constant inside (one uniform switch per tile instead of one per element)
            auto epilogue = [&](auto act_c) {
                constexpr int ACT = decltype(act_c)::value;
                // 1. new state values in place (element-wise, no branches inside: 4 * NT8 independent chains)
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const float4 b4 = ld4(sBias + 16 * q + 4 * ft);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            acc[2 * q + (e >> 1)][2 * h + (e & 1)] = act_apply(ACT, acc[2 * q + (e >> 1)][2 * h + (e & 1)] + bb[e]);
                }
                if (affine) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float4 a4 = ld4(sAff + 16 * q + 4 * ft), c4 = ld4(sAff + DP + 16 * q + 4 * ft);
                        const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                acc[2 * q + (e >> 1)][2 * h + (e & 1)] = fmaf(aa[e], acc[2 * q + (e >> 1)][2 * h + (e & 1)], cc[e]);
                    }
                }
                if (D < DP) {   // padding columns stay zero
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (16 * q + 4 * ft + e >= D) acc[2 * q + (e >> 1)][e & 1] = acc[2 * q + (e >> 1)][2 + (e & 1)] = 0.f;
                }
                // 2. stores: 16 bytes per lane, 4 lanes = two 32-byte sectors
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = 16 * mwarp + fg + 8 * h;
                    if (row < nvalid) {
                        float* orow = p.x_out + (size_t)(p.row_offset + n0 + row) * DP + 4 * ft;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const float4 y = make_float4(acc[2 * q][2 * h], acc[2 * q][2 * h + 1], acc[2 * q + 1][2 * h], acc[2 * q + 1][2 * h + 1]);
                            st4_hint(orow + 16 * q, y, stream_pol);
                            if (PEERS) {       // the same 16 bytes into every peer that gathers from this row (mask fetched a tile ago)
                                uint32_t nd = (need_cur >> (8 * h)) & 0xffu;
                                while (nd) {
                                    const int r = __ffs(nd) - 1;
                                    nd &= nd - 1u;
                                    st4(p.peer_out[r] + (size_t)(p.row_offset + n0 + row) * DP + 4 * ft + 16 * q, y);
                                }
                            }
                        }
                    }
                }
                // 3. BatchNormalization batch statistics (training) or the convergence test against the old state
                if (p.bn_train) {
                    // column sums over this warp's 16 rows: fp32 over the 2 rows of a lane and the 8 lanes that share its
                    // columns (xor 4, 8, 16), then fp64 accumulation in the warp's private shared-memory row
                    const bool v0 = 16 * mwarp + fg < nvalid, v1 = 16 * mwarp + fg + 8 < nvalid;
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float y0 = v0 ? acc[2 * q + (e >> 1)][e & 1] : 0.f, y1 = v1 ? acc[2 * q + (e >> 1)][2 + (e & 1)] : 0.f;
                            float s1 = y0 + y1, s2 = fmaf(y0, y0, y1 * y1);
#pragma unroll
                            for (int off = 4; off < 32; off <<= 1) {
                                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                                s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                            }
                            if (fg == 0) {
                                bn_mine[16 * q + 4 * ft + e] += (double)s1;
                                bn_mine[DP + 16 * q + 4 * ft + e] += (double)s2;
                            }
                        }
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float d2 = 0.f, o2 = 0.f;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const float xo[4] = {xcur[h][q].x, xcur[h][q].y, xcur[h][q].z, xcur[h][q].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float dx = acc[2 * q + (e >> 1)][2 * h + (e & 1)] - xo[e];
                                d2 = fmaf(dx, dx, d2);
                                o2 = fmaf(xo[e], xo[e], o2);
                            }
                        }
                        d2 += __shfl_xor_sync(0xffffffffu, d2, 1); o2 += __shfl_xor_sync(0xffffffffu, o2, 1);
                        d2 += __shfl_xor_sync(0xffffffffu, d2, 2); o2 += __shfl_xor_sync(0xffffffffu, o2, 2);
                        any_moving |= (16 * mwarp + fg + 8 * h < nvalid) && (sqrtf(d2) > p.thr * sqrtf(o2));
                    }
                }
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
        }

        if (p.bn_train) {
            // per-CTA partial sums: the MLP warps' private rows added in warp order (deterministic)
            named_bar_sync(GNN_BAR_MLP_ALL, WS_MLP_ALL);
            if (tid < 2 * DP) {
                const int which = tid / DP, j = tid % DP;
                double sum = 0.;
                for (int w = 0; w < 4 * WS_MLP_GROUPS; ++w) sum += bn_acc[w][which][j];
                p.bn_partial[(size_t)blockIdx.x * 2 * DP + which * DP + j] = sum;
            }
        } else {
            if (p.go_next && __any_sync(0xffffffffu, any_moving) && (tid & 31) == 0) s_flag = 1;
            if (PEERS && p.n_peers > 1) __threadfence_system();   // peer stores performed before the kernel is reported complete
            named_bar_sync(GNN_BAR_MLP_ALL, WS_MLP_ALL);
            if (tid == 0) iter_end(p, s_flag);
        }
    }
}

}  // namespace gnn
