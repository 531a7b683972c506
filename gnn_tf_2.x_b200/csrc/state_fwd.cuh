// state_fwd.cuh -- the fused forward iteration kernel of the state-convergence loop (sm_100a).
//
// One launch = one iteration t of the reference's while_loop body + the condition for t+1
// (GNN/GNN.py:223-242 and :202-220).  Per tile of TN nodes a CTA
//   1. stages the tile's row pointers and arc (source) indices in shared memory (coalesced),
//   2. GATHER: LPN = DP/4 adjacent lanes own one destination node; each lane pulls 16 bytes of every incoming
//      source row with 128-bit read-only loads, accumulating the segment sum in registers in stored (ascending
//      source) order -- no atomics, deterministic; own state row and the per-node constant row are loaded the
//      same way; everything lands in a shared-memory tile [TN][SA] = [state | agg | cst],
//   3. MLP: Dense layers as register-blocked (8 nodes x 4 units per thread) FMA loops over the tile, weights
//      in shared memory (staged once per CTA), bias + activation (+ dropout) fused,
//   4. EPILOGUE: final affine (BatchNormalization in inference mode), 128-bit coalesced store of the new state,
//      per-node test ||x_new - x|| > thr * ||x|| reduced by shuffles over the node's lanes and a ballot per
//      warp into one atomicOr on the device flag of iteration t+1.
// The kernel starts by reading the flag of iteration t and returns at once when the loop has already
// stopped, so the host can enqueue max_iter launches without ever synchronising.
#pragma once
#include "net_layout.cuh"

namespace gnn {

struct IterParams {
    // graph (destination-sorted CSR)
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;        // NULL => row scale in cst[:, C]
    const float* row_scale;  // [N] the same scale as a contiguous array (bulk-copied per tile by the warp-specialised kernel)
    long long N;
    long long E;             // entries of col / val
    long long row_offset;    // global id of local row 0 (node-range partition; 0 on a single GPU)
    int n_peers, rank;       // fused NVLink exchange: new rows are also stored into the peers' state buffers
    float* peer_out[GNN_MAX_PEERS];
    const uint32_t* peer_mask;
    // in-kernel cross-GPU signalling of the node-range partition (no host-side collective per iteration): every rank owns a signal
    // area in peer-mapped memory, arrive[sig_iters][8] | flag[sig_iters][8] (int32, slot [t][r] written by rank r): the last CTA of
    // iteration t stores its rank's convergence flag and then an arrival mark (the call's epoch, release, system scope) into every
    // peer's area; the CTAs of iteration t+1 wait for all marks, OR the flags and run or stop together
    int32_t* sig_local;
    int32_t* sig_peer[GNN_MAX_PEERS];
    uint32_t sig_epoch;
    int sig_iters;
    int* stopped;            // local: set once the loop has stopped, later launches return at once
    int* done_ctr;           // local: CTAs of this launch that are done (zeroed with the loop control words)
    // state
    const float* x_in;       // [N, DP]
    float* x_out;            // [N, DP]  (pre-BN output h_t when bn_train)
    float* agg_save;         // [N, DP] or NULL
    const float* cst;        // [N, CP]
    const float* wpack;      // packed net
    // loop control
    const int* go_cur;       // flag of this iteration
    int* go_next;            // flag of the next one (NULL: do not test)
    int* k_ptr;
    int t;
    float thr;
    double* bn_partial;      // [gridDim.x][2][DP] when bn_train
    int bn_train;
    uint32_t seed;
    const uint32_t* seed_dev;   // when set, the dropout seed of the call is read from device memory (CUDA-graph replays)
    int training;
    int scol_cap;
    int ring_slots, slot_rows;   // warp-specialised kernel: landing ring = ring_slots x slot_rows state rows
    int ws_debug;                // performance experiments only (GNN_B200_WS_DEBUG): 1 no row copies, 2 no segment sums, 4 no mma
    NetLayout net;
};

// ---------------------------------------------------------------------------------------------------------
// register-blocked dense layer over a shared-memory tile: out[n, j] = sum_k in[n, k] * W[k, j]
// thread micro-tile: 8 nodes (ng + NG*i) x 4 units (4cg .. 4cg+3)
// ---------------------------------------------------------------------------------------------------------
template <int TN>
__device__ __forceinline__ void dense_micro(const float* __restrict__ inb, int row_step, int K,
                                            const float* __restrict__ wb, int HP, float (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const float4 w0 = ld4(wb + (k + 0) * HP);
        const float4 w1 = ld4(wb + (k + 1) * HP);
        const float4 w2 = ld4(wb + (k + 2) * HP);
        const float4 w3 = ld4(wb + (k + 3) * HP);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 a = ld4(inb + i * row_step + k);
            acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
        }
    }
}

// all micro-tiles of one layer; epi(ng, cg, acc) consumes the accumulators (no barrier inside)
template <int TN, int NT, typename Epilogue>
__device__ __forceinline__ void dense_tile(const float* __restrict__ in, int in_stride, int K,
                                           const float* __restrict__ W, int HP, Epilogue epi) {
    constexpr int NG = TN / 8;
    const int CG = HP >> 2;
    for (int mt = threadIdx.x; mt < NG * CG; mt += NT) {
        const int cg = mt % CG, ng = mt / CG;
        float acc[8][4];
        dense_micro<TN>(in + ng * in_stride, NG * in_stride, K, W + 4 * cg, HP, acc);
        epi(ng, cg, acc);
    }
}

// segment sum of the incoming source rows of one node; STAGED: indices / weights come from shared memory
template <int DP, bool HAS_VAL, bool STAGED>
__device__ __forceinline__ float4 gather_rows(const float* __restrict__ x_in, int lane_off, int e0, int e1,
                                              const int32_t* __restrict__ cols, const float* __restrict__ vals) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
        int s0, s1, s2, s3;
        if (STAGED) { s0 = cols[e]; s1 = cols[e + 1]; s2 = cols[e + 2]; s3 = cols[e + 3]; }
        else { s0 = __ldg(cols + e); s1 = __ldg(cols + e + 1); s2 = __ldg(cols + e + 2); s3 = __ldg(cols + e + 3); }
        const float4 r0 = ldg4(x_in + (size_t)s0 * DP + lane_off);
        const float4 r1 = ldg4(x_in + (size_t)s1 * DP + lane_off);
        const float4 r2 = ldg4(x_in + (size_t)s2 * DP + lane_off);
        const float4 r3 = ldg4(x_in + (size_t)s3 * DP + lane_off);
        if (HAS_VAL) {
            float w0, w1, w2, w3;
            if (STAGED) { w0 = vals[e]; w1 = vals[e + 1]; w2 = vals[e + 2]; w3 = vals[e + 3]; }
            else { w0 = __ldg(vals + e); w1 = __ldg(vals + e + 1); w2 = __ldg(vals + e + 2); w3 = __ldg(vals + e + 3); }
            acc = fma4(w0, r0, acc); acc = fma4(w1, r1, acc); acc = fma4(w2, r2, acc); acc = fma4(w3, r3, acc);
        } else {
            acc = add4(acc, r0); acc = add4(acc, r1); acc = add4(acc, r2); acc = add4(acc, r3);
        }
    }
    for (; e < e1; ++e) {
        const int s = STAGED ? cols[e] : __ldg(cols + e);
        const float4 r = ldg4(x_in + (size_t)s * DP + lane_off);
        if (HAS_VAL) acc = fma4(STAGED ? vals[e] : __ldg(vals + e), r, acc);
        else acc = add4(acc, r);
    }
    return acc;
}

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const void* ptr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(void* ptr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

// Does iteration p.t run?  Uniform over the grid; contains a block-wide barrier (call it before any role split).
// Single GPU / host-side exchange: the device flag go[t].  In-kernel signalling: wait for every rank's arrival mark of this
// iteration (written by the last CTA of its previous launch), then OR the ranks' flags.
__device__ __forceinline__ bool iter_begin(const IterParams& p) {
    if (p.sig_local == nullptr) return *reinterpret_cast<const volatile int*>(p.go_cur) != 0;
    __shared__ int s_go;
    if (threadIdx.x == 0) {
        int go = 0;
        if (*reinterpret_cast<const volatile int*>(p.stopped)) go = 0;
        else if (p.t == 0) go = *reinterpret_cast<const volatile int*>(p.go_cur);     // first test: evaluated on all rows by every rank
        else
            for (int r = 0; r < p.n_peers; ++r) {
                while ((int32_t)(ld_acquire_sys_u32(p.sig_local + (size_t)p.t * 8 + r) - p.sig_epoch) < 0) __nanosleep(40);
                go |= reinterpret_cast<const volatile int32_t*>(p.sig_local)[(size_t)(p.sig_iters + p.t) * 8 + r];
            }
        s_go = go;
        if (!go && blockIdx.x == 0) *p.stopped = 1;
    }
    __syncthreads();
    return s_go != 0;
}

// End of a CTA's work on iteration p.t; ONE thread per CTA, after a block-level barrier behind which every thread has made its
// (peer) stores and executed __threadfence_system().  cta_flag: some node of this CTA still moves.
__device__ __forceinline__ void iter_end(const IterParams& p, int cta_flag) {
    if (p.go_next && cta_flag) atomicOr(p.go_next, 1);
    if (blockIdx.x == 0) *p.k_ptr = p.t + 1;
    if (p.sig_local == nullptr || p.go_next == nullptr) return;
    __threadfence_system();
    if (atomicAdd(p.done_ctr, 1) != (int)gridDim.x - 1) return;
    // last CTA of this rank: everything this launch stored (locally and into the peers) is ordered before the marks below
    __threadfence_system();
    const int flag = atomicOr(p.go_next, 0);
    for (int r = 0; r < p.n_peers; ++r) {
        int32_t* area = r == p.rank ? p.sig_local : p.sig_peer[r];
        reinterpret_cast<volatile int32_t*>(area)[(size_t)(p.sig_iters + p.t + 1) * 8 + p.rank] = flag;
    }
    __threadfence_system();
    for (int r = 0; r < p.n_peers; ++r) {
        int32_t* area = r == p.rank ? p.sig_local : p.sig_peer[r];
        st_release_sys_u32(area + (size_t)(p.t + 1) * 8 + p.rank, p.sig_epoch);
    }
}

#ifndef GNN_GATHER_BATCH
#define GNN_GATHER_BATCH 8
#endif
#ifndef GNN_FWD_MIN_CTAS
#define GNN_FWD_MIN_CTAS 1
#endif

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16_hint(float* smem_dst, const float* gmem_src, uint64_t policy) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "l"(policy));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// store one 16-byte piece of a new state row into the state buffers of the peers that gather from this row
__device__ __forceinline__ void store_to_peers(const IterParams& p, long long local_row, int col4, float4 v) {
    const uint32_t need = p.peer_mask ? __ldg(p.peer_mask + local_row) : 0xffffffffu;
    const size_t off = (size_t)(p.row_offset + local_row) * (size_t)p.net.DP + col4;
#pragma unroll
    for (int r = 0; r < GNN_MAX_PEERS; ++r)
        if (r < p.n_peers && r != p.rank && ((need >> r) & 1u)) *reinterpret_cast<float4*>(p.peer_out[r] + off) = v;
}

__device__ __forceinline__ void store_pair_to_peers(const IterParams& p, long long local_row, int col, float y0, float y1) {
    const uint32_t need = p.peer_mask ? __ldg(p.peer_mask + local_row) : 0xffffffffu;
    const size_t off = (size_t)(p.row_offset + local_row) * (size_t)p.net.DP + col;
#pragma unroll
    for (int r = 0; r < GNN_MAX_PEERS; ++r)
        if (r < p.n_peers && r != p.rank && ((need >> r) & 1u)) *reinterpret_cast<float2*>(p.peer_out[r] + off) = make_float2(y0, y1);
}

__device__ __forceinline__ float drop1(float v, bool active, uint32_t key, uint64_t idx, float rate, float scale) {
    if (!active) return v;
    return dropout_keep(key, idx, rate) ? v * scale : 0.f;
}

// Segment sums of NPG consecutive nodes (i0 .. i0+NPG-1 of the tile) by one lane group: the arcs of those nodes
// are one contiguous range [srow[i0], srow[i0+NPG]) walked in batches of GB.  Per batch: (1) all source indices
// (and weights), (2) all GB row loads back to back -- nothing with a scoreboard in between, so they overlap --
// (3) accumulation in stored order, flushing the sum of a node to the tile when the walk crosses its row pointer.
// cols / vals are indexed by the GLOBAL arc position (the caller pre-offsets the shared-memory copies).
template <int DP, bool HAS_VAL, bool STAGED, int GB, int NPG>
__device__ __forceinline__ void gather_group(const float* __restrict__ x_in, const int32_t* cols, const float* vals,
                                             const int* srow, const float* sscale, int i0, int lig, int nvalid,
                                             float* tile_agg, int SA, float* agg_save) {
    int i = i0;
    int e = srow[i0];
    const int eend = srow[i0 + NPG];
    int next = srow[i0 + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&]() {
        if (!HAS_VAL) {
            const float sc = sscale[i];
            acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
        }
        if (agg_save && i < nvalid) st4_hint(agg_save + (size_t)i * DP + 4 * lig, acc, l2_policy_evict_first());
        st4(tile_agg + i * SA + 4 * lig, acc);
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
        ++i;
        next = srow[i + 1 <= i0 + NPG ? i + 1 : i0 + NPG];
    };
    const float* xl = x_in + 4 * lig;
    const uint64_t keep = l2_policy_evict_last();
    while (e < eend) {
        int idx[GB];
        float w[GB];
        float4 r[GB];
#pragma unroll
        for (int b = 0; b < GB; ++b) {   // (1) indices: past the end -> repeat the last arc (its row is loaded, not summed)
            const int ee = min(e + b, eend - 1);
            idx[b] = STAGED ? cols[ee] : __ldg(cols + ee);
            if (HAS_VAL) w[b] = STAGED ? vals[ee] : __ldg(vals + ee);
        }
#pragma unroll
        for (int b = 0; b < GB; ++b) r[b] = ldg4_hint(xl + (size_t)idx[b] * DP, keep);   // (2) GB loads in flight
#pragma unroll
        for (int b = 0; b < GB; ++b) {   // (3) stored-order accumulation
            const int ee = e + b;
            if (ee < eend) {
                while (ee >= next) flush();
                if (HAS_VAL) acc = fma4(w[b], r[b], acc);
                else acc = add4(acc, r[b]);
            }
        }
        e += GB;
    }
    while (i < i0 + NPG) flush();
}

template <int DP, bool HAS_VAL, int TN, int NT>
__global__ void __launch_bounds__(NT, GNN_FWD_MIN_CTAS) state_iter_kernel(const IterParams p) {
    constexpr int GB = GNN_GATHER_BATCH;  // arcs in flight per lane
    static_assert(TN % 32 == 0 && NT % 32 == 0 && TN % 8 == 0, "tile shape");
    constexpr int LPN = DP / 4;          // lanes per node row
    constexpr int NGRP = NT / LPN;       // node rows gathered concurrently by a CTA
    static_assert(LPN >= 1 && LPN <= 32 && NT % LPN == 0, "lane mapping");

    if (!iter_begin(p)) return;  // loop already stopped (uniform)

    const NetLayout& net = p.net;
    const int tid = threadIdx.x;
    const int SA = net.SA, SB = net.SB, CP = net.CP, KP = net.KP, D = net.D;

    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* bufA = sW + net.fwd_floats;
    float* bufB = bufA + TN * SA;
    int* srow = reinterpret_cast<int*>(bufB + TN * SB);
    int* scol = srow + ((TN + 1 + 3) & ~3);
    float* sval = reinterpret_cast<float*>(scol + p.scol_cap);
    __shared__ int s_flag;
    __shared__ float sscale[TN];  // per-node arc weight of the tile (row-scale mode)

    for (int i = tid * 4; i < net.fwd_floats; i += NT * 4) st4(sW + i, ldg4(p.wpack + i));
    if (tid == 0) s_flag = 0;

    const int grp = tid / LPN, lig = tid % LPN;
    const bool drop_in = p.training && net.drop[0] > 0.f;
    const uint32_t key_in = dropout_key(call_seed(p), 0u, (uint32_t)p.t);
    const float scale_in = drop_in ? 1.f / (1.f - net.drop[0]) : 1.f;
    const int F_in = net.in_dim[0];

    // BatchNormalization batch statistics of this CTA (training mode): columns 4*lig .. 4*lig+3
    double bn_s1[4] = {0., 0., 0., 0.}, bn_s2[4] = {0., 0., 0., 0.};

    const long long ntiles = (p.N + TN - 1) / TN;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n0 = tile * TN;
        const int nvalid = (int)min((long long)TN, p.N - n0);
        __syncthreads();  // previous tile fully consumed (also orders the weight staging on the first pass)
        for (int i = tid; i <= TN; i += NT) srow[i] = __ldg(p.rowptr + min(n0 + i, p.N));
        if (!HAS_VAL)
            for (int i = tid; i < TN; i += NT) sscale[i] = i < nvalid ? __ldg(p.cst + (size_t)(n0 + i) * CP + net.C) : 0.f;
        __syncthreads();
        const int ebase = srow[0];
        const int ecount = srow[TN] - ebase;
        const bool staged = ecount <= p.scol_cap;
        if (staged) {
            for (int i = tid; i < ecount; i += NT) {
                scol[i] = __ldg(p.col + ebase + i);
                if (HAS_VAL) sval[i] = __ldg(p.val + ebase + i);
            }
        }
        __syncthreads();

        // ---- 2. gather ------------------------------------------------------------------------------
        // (a) own state rows and constant rows: asynchronous 16-byte copies straight into the tile (no registers)
        for (int item = tid; item < TN * LPN; item += NT) {
            const int i = item / LPN;
            float* dstp = bufA + i * SA + 4 * lig;
            if (i < nvalid) cp_async16(dstp, p.x_in + (size_t)(p.row_offset + n0 + i) * DP + 4 * lig);
            else st4(dstp, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        for (int item = tid; item < TN * (CP / 4); item += NT) {
            const int i = item / (CP / 4), c = item % (CP / 4);
            float* dstp = bufA + i * SA + 2 * DP + 4 * c;
            if (i < nvalid) cp_async16(dstp, p.cst + (size_t)(n0 + i) * CP + 4 * c);
            else st4(dstp, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        cp_async_commit();

        // (b) segment sums: a lane group owns NPG consecutive nodes and walks their arcs as ONE flat list in batches of
        //     GB arcs -- GB independent 128-bit loads in flight per lane whatever the degrees are; the running sum is
        //     flushed to the tile whenever the walk crosses a row pointer (stored order => deterministic)
        {
            constexpr int NPG = TN / NGRP;
            float* agg_dst = p.agg_save ? p.agg_save + (size_t)n0 * DP : nullptr;
            if (staged)
                gather_group<DP, HAS_VAL, true, GB, NPG>(p.x_in, scol - ebase, sval - ebase, srow, sscale, grp * NPG, lig, nvalid,
                                                         bufA + DP, SA, agg_dst);
            else
                gather_group<DP, HAS_VAL, false, GB, NPG>(p.x_in, p.col, p.val, srow, sscale, grp * NPG, lig, nvalid, bufA + DP,
                                                          SA, agg_dst);
        }
        cp_async_wait_all();
        __syncthreads();

        // (c) input dropout (training, Dropout in front of the first Dense): mask the assembled rows in place
        if (drop_in) {
            for (int item = tid; item < TN * (KP / 4); item += NT) {
                const int i = item / (KP / 4), c = item % (KP / 4);
                if (i >= nvalid) continue;
                float* q = bufA + i * SA + 4 * c;
                float4 v = ld4(q);
                float* vv = reinterpret_cast<float*>(&v);
                const uint64_t rbase = (uint64_t)(p.row_offset + n0 + i) * (uint64_t)F_in;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kc = keras_input_col(4 * c + u, D, DP, net.NL_self, net.NL_agg, net.AL);
                    if (kc >= 0) vv[u] = drop1(vv[u], true, key_in, rbase + kc, net.drop[0], scale_in);
                }
                st4(q, v);
            }
            __syncthreads();
        }

        // ---- 3. MLP ---------------------------------------------------------------------------------
        // hidden layers alternate bufB / bufA+DP; the last layer writes bufA + final_off.  When final_off == DP
        // and the layer reads bufA, source and destination overlap: the layout guarantees one micro-tile per
        // thread in that case, so the accumulators wait in registers across one barrier.
        constexpr int NG = TN / 8;
        for (int l = 0; l < net.L; ++l) {
            const bool last = (l == net.L - 1);
            const float* in = (l == 0) ? bufA : ((l & 1) ? bufB : bufA + DP);
            const int in_stride = (l == 0 || !(l & 1)) ? SA : SB;
            const float* W = sW + net.w_off[l];
            const float* bias = sW + net.b_off[l];
            const int HP = net.out_pad[l], act = net.act[l], odim = net.out_dim[l];
            const float rate = net.drop[l + 1];
            const bool drop_here = p.training && rate > 0.f;
            const uint32_t key = dropout_key(call_seed(p), (uint32_t)(l + 1), (uint32_t)p.t);
            const float dscale = drop_here ? 1.f / (1.f - rate) : 1.f;
            float* out;
            int out_stride;
            if (last) { out = bufA + net.final_off; out_stride = SA; }
            else if (l & 1) { out = bufA + DP; out_stride = SA; }
            else { out = bufB; out_stride = SB; }
            const bool hazard = last && !(l & 1) && net.final_off == DP;
            const float* aff_a = sW + net.aff_off;
            const float* aff_c = aff_a + HP;
            const bool affine = last && !p.bn_train;

            // the register-blocked product only stores raw pre-activations; bias, activation, dropout and the final
            // affine run as one element-wise pass over the tile (keeps the unrolled code small)
            auto epi = [&](int ng, int cg, float (&acc)[8][4]) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    st4(out + (ng + NG * i) * out_stride + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            };
            if (hazard) {
                const int CG = HP >> 2;
                const int cg = tid % CG, ng = tid / CG;
                const bool has_work = tid < NG * CG;  // NG*CG <= NT guaranteed by make_layout
                float acc[8][4];
                if (has_work) dense_micro<TN>(in + ng * in_stride, NG * in_stride, net.in_pad[l], W + 4 * cg, HP, acc);
                __syncthreads();
                if (has_work) epi(ng, cg, acc);
            } else {
                dense_tile<TN, NT>(in, in_stride, net.in_pad[l], W, HP, epi);
            }
            __syncthreads();
            {
                const int CG = HP >> 2;
                for (int item = tid; item < TN * CG; item += NT) {
                    const int row = item / CG, cg = item % CG;
                    float* q = out + row * out_stride + 4 * cg;
                    float4 v4 = ld4(q);
                    const float4 b4 = ld4(bias + 4 * cg);
                    float v[4] = {v4.x + b4.x, v4.y + b4.y, v4.z + b4.z, v4.w + b4.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = 4 * cg + c;
                        float y = act_apply(act, v[c]);
                        if (drop_here) y = drop1(y, j < odim, key, (uint64_t)(p.row_offset + n0 + row) * (uint64_t)odim + j, rate, dscale);
                        if (affine) y = fmaf(aff_a[j], y, aff_c[j]);
                        v[c] = (j < odim) ? y : 0.f;
                    }
                    st4(q, make_float4(v[0], v[1], v[2], v[3]));
                }
            }
            __syncthreads();
        }

        // ---- 4. epilogue: store + convergence test --------------------------------------------------
        bool any_moving = false;
        const uint64_t stream_pol = l2_policy_evict_first();
        for (int item = tid; item < TN * LPN; item += NT) {
            const int i = item / LPN;  // lig == item % LPN because NT % LPN == 0
            const long long n = n0 + i;
            const bool valid = i < nvalid;
            float4 xn = ld4(bufA + i * SA + net.final_off + 4 * lig);
            float d2 = 0.f, o2 = 0.f;
            if (valid) {
                st4_hint(p.x_out + (size_t)(p.row_offset + n) * DP + 4 * lig, xn, stream_pol);
                if (p.n_peers > 1) store_to_peers(p, n, 4 * lig, xn);
                if (p.bn_train) {
                    bn_s1[0] += xn.x; bn_s1[1] += xn.y; bn_s1[2] += xn.z; bn_s1[3] += xn.w;
                    bn_s2[0] += (double)xn.x * xn.x; bn_s2[1] += (double)xn.y * xn.y;
                    bn_s2[2] += (double)xn.z * xn.z; bn_s2[3] += (double)xn.w * xn.w;
                } else {
                    const float4 xo = drop_in ? ldg4(p.x_in + (size_t)(p.row_offset + n) * DP + 4 * lig) : ld4(bufA + i * SA + 4 * lig);
                    const float dx = xn.x - xo.x, dy = xn.y - xo.y, dz = xn.z - xo.z, dw = xn.w - xo.w;
                    d2 = dx * dx + dy * dy + dz * dz + dw * dw;
                    o2 = xo.x * xo.x + xo.y * xo.y + xo.z * xo.z + xo.w * xo.w;
                }
            }
            if (!p.bn_train) {
#pragma unroll
                for (int off = LPN / 2; off > 0; off >>= 1) {
                    d2 += __shfl_xor_sync(0xffffffffu, d2, off);
                    o2 += __shfl_xor_sync(0xffffffffu, o2, off);
                }
                any_moving |= valid && (sqrtf(d2) > p.thr * sqrtf(o2));
            }
        }
        if (!p.bn_train && p.go_next) {
            if (__any_sync(0xffffffffu, any_moving) && (tid & 31) == 0) s_flag = 1;
        }
    }

    __syncthreads();
    if (p.bn_train) {
        // reduce the per-thread column sums: lanes sharing lig inside a warp, then warps through shared memory
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            for (int off = LPN; off < 32; off <<= 1) {
                bn_s1[c] += __shfl_xor_sync(0xffffffffu, bn_s1[c], off);
                bn_s2[c] += __shfl_xor_sync(0xffffffffu, bn_s2[c], off);
            }
        }
        double* red = reinterpret_cast<double*>(bufA);  // [warps][LPN][8], fits (see net_layout.cuh)
        const int warp = tid >> 5, lane = tid & 31;
        if (lane < LPN) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                red[(warp * LPN + lane) * 8 + c] = bn_s1[c];
                red[(warp * LPN + lane) * 8 + 4 + c] = bn_s2[c];
            }
        }
        __syncthreads();
        if (tid < LPN) {
            double s1[4] = {0., 0., 0., 0.}, s2[4] = {0., 0., 0., 0.};
            for (int w = 0; w < NT / 32; ++w)
#pragma unroll
                for (int c = 0; c < 4; ++c) { s1[c] += red[(w * LPN + tid) * 8 + c]; s2[c] += red[(w * LPN + tid) * 8 + 4 + c]; }
            double* dst = p.bn_partial + (size_t)blockIdx.x * 2 * DP;
#pragma unroll
            for (int c = 0; c < 4; ++c) { dst[4 * tid + c] = s1[c]; dst[DP + 4 * tid + c] = s2[c]; }
        }
    } else {
        if (p.n_peers > 1) { __threadfence_system(); __syncthreads(); }   // peer stores performed before the kernel is reported complete
        if (tid == 0) iter_end(p, s_flag);
    }
}

}  // namespace gnn
