// state_fwd_ws.cuh -- warp-specialised, software-pipelined variant of the fused forward iteration kernel.
//
// Why: the gather of a uniform-random graph is latency bound.  A gather-only microbenchmark on B200
// (scripts/gather_microbench.cu) needs >= 1000 128-byte rows in flight per SM to reach the L2/HBM limit; the symmetric
// kernel (state_fwd.cuh) holds its rows in registers and cannot have that many in flight next to the MLP.  Here the
// rows land in shared memory through cp.async (no registers, deep queue) and two warp groups run concurrently:
//
//   gather warps (4): for tile j+1: stage row pointers / arc indices -> cp.async every source row into landing[(j+1)&1],
//                     own rows and constant rows straight into tile[(j+1)&1];  for tile j: wait for its rows, segment-sum
//                     them out of shared memory in stored order (deterministic, no atomics) into tile[j&1]  -> FULL[j&1]
//   MLP warps    (4): wait FULL[j&1] -> Dense layer as register-blocked FMA tiles (4 nodes x 4 units per thread, weights in
//                     shared memory) -> bias / activation / affine -> 128-bit coalesced store of the new state +
//                     convergence test (+ BatchNormalization batch statistics when training)          -> EMPTY[j&1]
//
// One persistent CTA of 256 threads per SM, 64-node tiles, double-buffered tiles and landing zones (2 x ~640 rows in
// flight per SM at the C4 shape).  Used for single-Dense-layer state nets (what the reference builds by default) with
// padded state width 16..32 and no active dropout; every other case runs the symmetric kernel.
#pragma once
#include "state_fwd.cuh"

namespace gnn {

#define GNN_BAR_GATHER 1
#define GNN_BAR_MLP0 2    // +group
#define GNN_BAR_FULL0 4   // +b
#define GNN_BAR_EMPTY0 6  // +b
#define GNN_BAR_MLP_ALL 8

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gmem_src));
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int WS_TN = 64;          // nodes per tile
constexpr int WS_MLP = 128;        // threads of ONE MLP group (4 warps); two groups: even / odd tiles
constexpr int WS_MLP_ALL = 256;    // both MLP groups (warps 0-7)
constexpr int WS_GATHER = 256;     // gather threads (warps 8-15)
constexpr int WS_THREADS = 512;
constexpr int WS_PAIR = WS_MLP + WS_GATHER;   // participants of a FULL / EMPTY barrier: one MLP group + the gather warps

// shared-memory footprint (bytes) for a landing capacity of `cap` rows per stage
static inline size_t ws_smem_bytes(const NetLayout& lay, int cap, bool has_val) {
    size_t fl = (size_t)lay.fwd_floats + 2 * (size_t)WS_TN * lay.SA + 2 * (size_t)cap * lay.DP + 4 * 68 + 4 * WS_TN +
                3 * (size_t)cap * (has_val ? 2 : 1);   // tiles x2, landing x2, row pointers / scales x4, arc indices x3
    return fl * 4;
}

template <int DP, bool HAS_VAL>
__global__ void __launch_bounds__(WS_THREADS, 1) state_iter_ws_kernel(const IterParams p) {
    constexpr int TN = WS_TN, LPN = DP / 4;
    constexpr int NGRP = WS_GATHER / LPN; // lane groups among the gather warps
    constexpr int NPG = TN / NGRP;        // consecutive nodes per lane group
    static_assert(NPG >= 1 && NGRP * NPG == TN, "lane mapping");
    constexpr int NG = TN / 4;            // MLP micro-tiles: 4 nodes (ng + NG*i) x 4 units
    static_assert(NG * (DP / 4) <= WS_MLP, "one micro-tile per MLP thread (in-place output)");

    if (*reinterpret_cast<const volatile int*>(p.go_cur) == 0) return;

    const NetLayout& net = p.net;
    const int tid = threadIdx.x;
    const int SA = net.SA, CP = net.CP, KP = net.KP, cap = p.scol_cap;

    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* tile0 = sW + net.fwd_floats;                 // [2][TN][SA]
    float* land0 = tile0 + 2 * TN * SA;                 // [2][cap][DP]
    int* srow0 = reinterpret_cast<int*>(land0 + 2 * (size_t)cap * DP);   // [4][68]
    float* sscale0 = reinterpret_cast<float*>(srow0 + 4 * 68);           // [4][TN]
    int* scol0 = reinterpret_cast<int*>(sscale0 + 4 * TN);               // [3][cap]
    float* sval0 = reinterpret_cast<float*>(scol0 + 3 * cap);            // [3][cap] (HAS_VAL)
    __shared__ int s_flag;

    for (int i = tid * 4; i < net.fwd_floats; i += WS_THREADS * 4) st4(sW + i, ldg4(p.wpack + i));
    if (tid == 0) s_flag = 0;
    __syncthreads();

    const long long ntiles = (p.N + TN - 1) / TN;
    const long long first = blockIdx.x, stride = gridDim.x;
    const bool is_gather = tid >= WS_MLP_ALL;

    if (is_gather) {
        // =========================================== GATHER WARPS ==============================================
        const int gt = tid - WS_MLP_ALL;
        const int grp = gt / LPN, lig = gt % LPN;
        const uint64_t keep = l2_policy_evict_last();

        // Buffers: tiles and landing zones x2 (b = it & 1), arc indices x3 (q3 = seq % 3), row pointers / scales x4 (q4).
        // row pointers (+ per-node weights) of a tile: asynchronous 4-byte copies, they ride in the row group
        auto stage_rowptr_async = [&](long long tile, int q4) {
            const long long n0 = tile * TN;
            for (int i = gt; i <= TN; i += WS_GATHER) cp_async4(srow0 + q4 * 68 + i, p.rowptr + min(n0 + i, p.N));
            if (!HAS_VAL)
                for (int i = gt; i < TN; i += WS_GATHER) {
                    if (n0 + i < p.N) cp_async4(sscale0 + q4 * TN + i, p.cst + (size_t)(n0 + i) * CP + net.C);
                    else sscale0[q4 * TN + i] = 0.f;
                }
        };
        // arc sources of a tile: loaded into registers first (so that independent work can overlap the latency) ...
        constexpr int CREG = 8;
        int creg[CREG];
        float vreg[CREG];
        auto load_cols = [&](int q4) {
            const int* srow = srow0 + q4 * 68;
            const int ebase = srow[0];
            const int ecount = min(srow[TN] - ebase, cap);
#pragma unroll
            for (int u = 0; u < CREG; ++u) {
                const int r = gt + u * WS_GATHER;
                creg[u] = 0; vreg[u] = 0.f;
                if (r < ecount) {
                    creg[u] = __ldg(p.col + ebase + r);
                    if (HAS_VAL) vreg[u] = __ldg(p.val + ebase + r);
                }
            }
        };
        // ... and stored to the index buffer afterwards
        auto store_cols = [&](int q4, int q3) {
            const int* srow = srow0 + q4 * 68;
            const int ebase = srow[0];
            const int ecount = min(srow[TN] - ebase, cap);
#pragma unroll
            for (int u = 0; u < CREG; ++u) {
                const int r = gt + u * WS_GATHER;
                if (r < ecount) {
                    scol0[(size_t)q3 * cap + r] = creg[u];
                    if (HAS_VAL) sval0[(size_t)q3 * cap + r] = vreg[u];
                }
            }
            for (int r = gt + CREG * WS_GATHER; r < ecount; r += WS_GATHER) {   // very dense tiles only
                scol0[(size_t)q3 * cap + r] = __ldg(p.col + ebase + r);
                if (HAS_VAL) sval0[(size_t)q3 * cap + r] = __ldg(p.val + ebase + r);
            }
        };
        // every source row of the tile -> landing zone b (asynchronous, no registers); a lane group takes chunks of 4
        // consecutive arcs so that their 4 indices are one 128-bit shared-memory load
        auto issue_rows = [&](int q4, int q3, int b) {
            const int* srow = srow0 + q4 * 68;
            const int ecount = min(srow[TN] - srow[0], cap);
            float* lb = land0 + (size_t)b * cap * DP + 4 * lig;
            const int* scol = scol0 + (size_t)q3 * cap;
            const float* xl = p.x_in + 4 * lig;
            for (int r = 4 * grp; r < ecount; r += 4 * NGRP) {
                const int4 s4 = *reinterpret_cast<const int4*>(scol + r);   // entries past ecount are never used
                cp_async16_hint(lb + (size_t)r * DP, xl + (size_t)s4.x * DP, keep);
                if (r + 1 < ecount) cp_async16_hint(lb + (size_t)(r + 1) * DP, xl + (size_t)s4.y * DP, keep);
                if (r + 2 < ecount) cp_async16_hint(lb + (size_t)(r + 2) * DP, xl + (size_t)s4.z * DP, keep);
                if (r + 3 < ecount) cp_async16_hint(lb + (size_t)(r + 3) * DP, xl + (size_t)s4.w * DP, keep);
            }
        };
        // segment sums of this lane group's NPG nodes out of the landing zone, stored order
        auto consume = [&](long long tile, int q4, int q3, int b) {
            const long long n0 = tile * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            const int* srow = srow0 + q4 * 68;
            const int ebase = srow[0];
            const float* lb = land0 + (size_t)b * cap * DP + 4 * lig;
            const float* sv = sval0 + (size_t)q3 * cap;
            float* tb = tile0 + (size_t)b * TN * SA + DP + 4 * lig;
#pragma unroll 1
            for (int u = 0; u < NPG; ++u) {
                const int i = grp * NPG + u;
                const int r0 = srow[i] - ebase, r1 = srow[i + 1] - ebase;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int r = r0;
                const int rl = min(r1, cap);
#pragma unroll 4
                for (; r < rl; ++r) {
                    const float4 v = ld4(lb + (size_t)r * DP);
                    if (HAS_VAL) acc = fma4(sv[r], v, acc);
                    else acc = add4(acc, v);
                }
                for (; r < r1; ++r) {   // arcs beyond the landing capacity: direct loads (rare, high-degree tiles)
                    const int s = __ldg(p.col + ebase + r);
                    const float4 v = ldg4(p.x_in + (size_t)s * DP + 4 * lig);
                    if (HAS_VAL) acc = fma4(__ldg(p.val + ebase + r), v, acc);
                    else acc = add4(acc, v);
                }
                if (!HAS_VAL) {
                    const float sc = sscale0[q4 * TN + i];
                    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
                }
                if (p.agg_save && i < nvalid) st4_hint(p.agg_save + (size_t)(n0 + i) * DP + 4 * lig, acc, l2_policy_evict_first());
                st4(tb + i * SA, acc);
            }
        };

        // Pipeline (the gather warps run one tile ahead of the MLP warps; nothing they wait for is on the critical path):
        //   iteration it : rows(it+1) -> landing[b^1] and row pointers(it+3) -> srow        [asynchronous, group X_it]
        //                  wait X_{it-1}: rows(it) and row pointers(it+2) have landed
        //                  arc sources(it+2) -> registers (latency overlaps consume)
        //                  wait EMPTY[b] (MLP group b finished tile it-2, long ago) ; consume(it) -> tile[b] ; FULL[b]
        //                  arc sources(it+2) -> scol
        // (own state rows and constant rows are fetched by the MLP group that owns the tile buffer)
        auto tile_at = [&](int seq) { return first + (long long)seq * stride; };
        for (int sq = 0; sq < 3; ++sq)
            if (tile_at(sq) < ntiles) stage_rowptr_async(tile_at(sq), sq);
        cp_async_commit();
        cp_async_wait_group<0>();
        named_bar_sync(GNN_BAR_GATHER, WS_GATHER);
        for (int sq = 0; sq < 2; ++sq)
            if (tile_at(sq) < ntiles) { load_cols(sq); store_cols(sq, sq); }
        named_bar_sync(GNN_BAR_GATHER, WS_GATHER);
        if (first < ntiles) issue_rows(0, 0, 0);
        cp_async_commit();   // plays the role of X_{-1}
        int it = 0;
        for (long long tile = first; tile < ntiles; tile += stride, ++it) {
            const int b = it & 1;
            const long long t1 = tile + stride, t2 = tile + 2 * stride, t3 = tile + 3 * stride;
            named_bar_sync(GNN_BAR_GATHER, WS_GATHER);     // consume(it-1) and store_cols(it+1) are done in every gather thread
            if (t1 < ntiles) issue_rows((it + 1) & 3, (it + 1) % 3, b ^ 1);
            if (t3 < ntiles) stage_rowptr_async(t3, (it + 3) & 3);
            cp_async_commit();                             // X_it
            cp_async_wait_group<1>();                      // everything older than X_it has landed
            named_bar_sync(GNN_BAR_GATHER, WS_GATHER);     // ... for every gather thread
            if (t2 < ntiles) load_cols((it + 2) & 3);
            if (it >= 2) named_bar_sync(GNN_BAR_EMPTY0 + b, WS_PAIR);   // MLP group b is done with tile it-2
            consume(tile, it & 3, it % 3, b);
            __threadfence_block();
            named_bar_arrive(GNN_BAR_FULL0 + b, WS_PAIR);
            if (t2 < ntiles) store_cols((it + 2) & 3, (it + 2) % 3);
        }
        cp_async_wait_group<0>();
    } else {
        // ============================================= MLP WARPS ==============================================
        // two MLP groups of 4 warps: group g takes the tiles with sequence number it = g, g+2, ... (= tile buffer g)
        constexpr int CG = DP / 4;
        const int mgroup = tid / WS_MLP, mt = tid % WS_MLP;
        const int cg = mt % CG, ng = mt / CG;
        const bool has_item = mt < NG * CG;
        const int lig = mt % LPN;
        const float* W = sW + net.w_off[0];
        const float* bias = sW + net.b_off[0];
        const float* aff_a = sW + net.aff_off;
        const float* aff_c = aff_a + DP;
        const int act = net.act[0], D = net.D;
        const bool affine = !p.bn_train;
        const uint64_t stream_pol = l2_policy_evict_first();
        double bn_s1[4] = {0., 0., 0., 0.}, bn_s2[4] = {0., 0., 0., 0.};
        bool any_moving = false;

        // own state rows and constant rows of a tile -> this group's tile buffer (asynchronous; the gather warps only
        // write the aggregate columns, so the two never touch the same bytes)
        float* tb = tile0 + (size_t)mgroup * TN * SA;
        auto issue_own = [&](long long tile) {
            const long long n0 = tile * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            for (int item = mt; item < TN * LPN; item += WS_MLP) {
                const int i = item / LPN;
                float* dstp = tb + i * SA + 4 * lig;
                if (i < nvalid) cp_async16(dstp, p.x_in + (size_t)(p.row_offset + n0 + i) * DP + 4 * lig);
                else st4(dstp, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            for (int item = mt; item < TN * (CP / 4); item += WS_MLP) {
                const int i = item / (CP / 4), c = item % (CP / 4);
                float* dstp = tb + i * SA + 2 * DP + 4 * c;
                if (i < nvalid) cp_async16(dstp, p.cst + (size_t)(n0 + i) * CP + 4 * c);
                else st4(dstp, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            cp_async_commit();
        };
        if (first + (long long)mgroup * stride < ntiles) issue_own(first + (long long)mgroup * stride);

        int it = mgroup;
        for (long long tile = first + (long long)mgroup * stride; tile < ntiles; tile += 2 * stride, it += 2) {
            const int b = mgroup;
            const long long n0 = tile * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            named_bar_sync(GNN_BAR_FULL0 + b, WS_PAIR);            // aggregates of this tile are in the buffer
            cp_async_wait_group<0>();                              // my own / constant rows too ...
            named_bar_sync(GNN_BAR_MLP0 + mgroup, WS_MLP);         // ... and those of the rest of the group

            // Dense layer: 4 nodes x 4 units per thread, k unrolled by 8
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            if (has_item) {
                const float* inb = tb + ng * SA;
                const float* wb = W + 4 * cg;
#pragma unroll 2
                for (int k = 0; k < KP; k += 4) {
                    const float4 w0 = ld4(wb + (k + 0) * DP), w1 = ld4(wb + (k + 1) * DP);
                    const float4 w2 = ld4(wb + (k + 2) * DP), w3 = ld4(wb + (k + 3) * DP);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = ld4(inb + i * NG * SA + k);
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
            // epilogue straight from the accumulators: bias + activation + affine, 128-bit store of the new state (8 lanes
            // = one 128-byte row), convergence test reduced over the CG lanes that share a node
            if (has_item) {
                const float4 b4 = ld4(bias + 4 * cg), a4 = ld4(aff_a + 4 * cg), c4 = ld4(aff_c + 4 * cg);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = ng + NG * i;
                    const long long n = n0 + row;
                    const bool valid = row < nvalid;
                    float v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float y = act_apply(act, acc[i][c] + bb[c]);
                        if (affine) y = fmaf(aa[c], y, cc[c]);
                        v[c] = (4 * cg + c < D) ? y : 0.f;
                    }
                    const float4 xn = make_float4(v[0], v[1], v[2], v[3]);
                    float d2 = 0.f, o2 = 0.f;
                    if (valid) {
                        st4_hint(p.x_out + (size_t)(p.row_offset + n) * DP + 4 * cg, xn, stream_pol);
                        if (p.n_peers > 1) store_to_peers(p, n, 4 * cg, xn);
                        if (p.bn_train) {
                            bn_s1[0] += xn.x; bn_s1[1] += xn.y; bn_s1[2] += xn.z; bn_s1[3] += xn.w;
                            bn_s2[0] += (double)xn.x * xn.x; bn_s2[1] += (double)xn.y * xn.y;
                            bn_s2[2] += (double)xn.z * xn.z; bn_s2[3] += (double)xn.w * xn.w;
                        } else {
                            const float4 xo = ld4(tb + row * SA + 4 * cg);
                            const float dx = xn.x - xo.x, dy = xn.y - xo.y, dz = xn.z - xo.z, dw = xn.w - xo.w;
                            d2 = dx * dx + dy * dy + dz * dz + dw * dw;
                            o2 = xo.x * xo.x + xo.y * xo.y + xo.z * xo.z + xo.w * xo.w;
                        }
                    }
                    if (!p.bn_train) {
#pragma unroll
                        for (int off = CG / 2; off > 0; off >>= 1) {
                            d2 += __shfl_xor_sync(0xffffffffu, d2, off);
                            o2 += __shfl_xor_sync(0xffffffffu, o2, off);
                        }
                        any_moving |= valid && (sqrtf(d2) > p.thr * sqrtf(o2));
                    }
                }
            }
            if (tile + 2 * stride < ntiles) {   // this buffer's next tile: fetch its own rows, then hand the buffer back
                named_bar_sync(GNN_BAR_MLP0 + mgroup, WS_MLP);   // the whole group has finished reading the tile
                issue_own(tile + 2 * stride);
                __threadfence_block();
                named_bar_arrive(GNN_BAR_EMPTY0 + b, WS_PAIR);
            }
        }

        if (p.bn_train) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                for (int off = LPN; off < 32; off <<= 1) {
                    bn_s1[c] += __shfl_xor_sync(0xffffffffu, bn_s1[c], off);
                    bn_s2[c] += __shfl_xor_sync(0xffffffffu, bn_s2[c], off);
                }
            // the landing zones are idle for the MLP warps' purposes only after the gather warps are done: use a private
            // static buffer instead (4 warps x LPN lanes x 8 doubles)
            __shared__ double red[8 * 8 * 8];
            const int warp = tid >> 5, lane = tid & 31;
            if (lane < LPN)
#pragma unroll
                for (int c = 0; c < 4; ++c) { red[(warp * LPN + lane) * 8 + c] = bn_s1[c]; red[(warp * LPN + lane) * 8 + 4 + c] = bn_s2[c]; }
            named_bar_sync(GNN_BAR_MLP_ALL, WS_MLP_ALL);
            if (tid < LPN) {
                double s1[4] = {0., 0., 0., 0.}, s2[4] = {0., 0., 0., 0.};
                for (int w = 0; w < 8; ++w)
#pragma unroll
                    for (int c = 0; c < 4; ++c) { s1[c] += red[(w * LPN + tid) * 8 + c]; s2[c] += red[(w * LPN + tid) * 8 + 4 + c]; }
                double* dst = p.bn_partial + (size_t)blockIdx.x * 2 * DP;
#pragma unroll
                for (int c = 0; c < 4; ++c) { dst[4 * tid + c] = s1[c]; dst[DP + 4 * tid + c] = s2[c]; }
            }
        } else {
            if (p.go_next && __any_sync(0xffffffffu, any_moving) && (tid & 31) == 0) s_flag = 1;
            if (p.n_peers > 1) __threadfence_system();   // peer stores performed before the kernel is reported complete
            named_bar_sync(GNN_BAR_MLP_ALL, WS_MLP_ALL);
            if (tid == 0) {
                if (p.go_next && s_flag) atomicOr(p.go_next, 1);
                if (blockIdx.x == 0) *p.k_ptr = p.t + 1;
            }
        }
    }
}

}  // namespace gnn
