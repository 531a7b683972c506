// state_bwd_l1.cuh -- BPTT node kernel for the state net the reference builds by default: ONE Dense layer (+ BatchNormalization),
// no active dropout, padded state width 16 or 32.  Same arithmetic contract as state_bwd_node_kernel (state_bwd.cuh), what
// tf.GradientTape replays for GNN/GNN.py:223-242 (reference GNN/GNN_BaseClass.py:233-237), restructured for throughput:
//
//   one persistent CTA per SM, 128-node tiles, three shared-memory stages, 20 warps in three roles joined by mbarriers (no
//   block-wide barrier inside the tile loop):
//     warps 16-19  producers: all five inputs of a tile (u_t = [x_t | A x_t | cst], dL/dx_{t+1}, y_t) land in shared memory by cp.async
//                  one tile ahead; delta = dL/dy * act'(y) (through the BatchNormalization backward when training) is then formed in
//                  place; db = column sums of delta stay in the producers' registers.  They run up to two tiles ahead of the
//                  arithmetic;
//     warps 0-7    dW += u^T delta : warp w owns nodes 16 w .. 16 w + 15 of every tile; lane (kb, jb) keeps an 8 x 8 block of the
//                  [2 DP x DP] gradient (+ its share of the constant rows) in REGISTERS for the whole kernel -- 5 shared-memory
//                  loads per 36 packed FMA (fma.rn.f32x2), nothing is reduced per tile;
//     warps 8-15   g_u = delta W^T : thread = (two nodes, one quarter of the 2 DP columns): 16 columns of g_self or g_agg of its
//                  two nodes in registers, W^T rows broadcast from shared memory, whole 32-byte sectors stored;
//   register budgets per role through setmaxnreg (128 / 80 / 64; the pools are per SM sub-partition: 2 + 2 + 1 warps each);
//   at the end the eight register copies of dW and the db pieces are added in a fixed order (deterministic) into the CTA's
//   partial slot.
//
// History (measured on the C4 graph, 1M nodes, per launch): phase-structured kernel of state_bwd.cuh 544 us; this arithmetic with
// block-wide barriers between load / delta / arithmetic phases 352 us (half of the time in the phases around the two FMA loops);
// producer warps + mbarriers 255 us (FMA pipe 61 % busy); inputs by cp.async one tile ahead: little more at DP = 32, 35 % off at
// DP = 16 (C5 batches), where a tile is too short to hide a DRAM latency.
#pragma once
#include "state_bwd.cuh"
#include "state_fwd_ws.cuh"
#include "state_fwd_tc.cuh"

namespace gnn {

constexpr int BL_TN = 128;       // nodes per tile
constexpr int BL_COMPUTE = 512;  // warps 0-7 weight gradient, warps 8-15 input gradient
constexpr int BL_PRODUCE = 128;  // warps 16-19
constexpr int BL_NT = BL_COMPUTE + BL_PRODUCE;
constexpr int BL_STAGES = 3;
constexpr int BL_CROWS = 16;     // constant rows held by the weight-gradient lanes (CP <= 16)
constexpr int BL_REGS_LAUNCH = 96, BL_REGS_DW = 128, BL_REGS_GU = 80, BL_REGS_PRODUCE = 64;
static_assert(BL_NT * BL_REGS_LAUNCH <= 65536, "register file at launch");
static_assert(2 * BL_REGS_DW + 2 * BL_REGS_GU + BL_REGS_PRODUCE <= 5 * BL_REGS_LAUNCH, "per sub-partition pool: 2 + 2 + 1 warps");

static inline int bwd_l1_su(const NetLayout& lay) { return odd_quad_stride(lay.KP); }      // row strides: 16-byte aligned, (stride / 4) odd
static inline int bwd_l1_sd(const NetLayout& lay) { return odd_quad_stride(lay.DP); }
static inline size_t bwd_l1_stage_floats(const NetLayout& lay) { return (size_t)BL_TN * (bwd_l1_su(lay) + bwd_l1_sd(lay)); }
static inline size_t bwd_l1_smem_bytes(const NetLayout& lay) {
    return ((size_t)lay.DP * lay.KP + BL_STAGES * bwd_l1_stage_floats(lay) + 2 * (size_t)BL_TN * lay.DP) * 4;   // W^T, stages, y_t of two tiles
}
static inline bool bwd_l1_applicable(const NetLayout& lay, bool training, bool y_saved) {
    if (lay.L != 1 || !y_saved || (lay.DP != 16 && lay.DP != 32) || lay.CP > BL_CROWS) return false;
    if (training && (lay.drop[0] > 0.f || lay.drop[1] > 0.f)) return false;
    // the final reduction re-uses the stages: 8 copies of [2 DP + 16][DP] and 4 db values per producer thread
    return 8 * (size_t)(2 * lay.DP + BL_CROWS) * lay.DP + 4 * (size_t)BL_PRODUCE <= BL_STAGES * bwd_l1_stage_floats(lay);
}

// d0 += a0 * b0, d1 += a1 * b1 as ONE instruction (fma.rn.f32x2 -> FFMA2): same IEEE result per element and the same FMA rate as
// two FFMA (127 vs 124 FMA / clk / SM measured, scripts/ffma_microbench.cu) for half the issue slots -- this kernel is issue-bound
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 d, a, b;\n\tmov.b64 d, {%0, %1};\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\t"
        "fma.rn.f32x2 d, a, b, d;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* gmem_src, bool ok) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(ok ? 16 : 0));
}

template <int DP>
static __global__ void __launch_bounds__(BL_NT, 1) state_bwd_node_l1_kernel(const BwdNodeParams p) {
    constexpr int TN = BL_TN, LPN = DP / 4, NS = BL_STAGES;
    constexpr int MK = DP / 4, MJ = DP / 4;      // block of the [2 DP x DP] gradient held by one lane: rows MK kb .., columns MJ jb ..
    constexpr int RROWS = 2 * DP + BL_CROWS;     // rows of one register copy of dW: main rows, constant rows
    static_assert(DP == 16 || DP == 32, "padded state width");
    static_assert(BL_PRODUCE % LPN == 0, "a producer thread forms delta for the same 4 columns in every pass");
    if (p.t >= *reinterpret_cast<const volatile int*>(p.k_ptr)) return;

    const NetLayout& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int CP = net.CP, KP = net.KP, D = net.D, SU = p.SU, SD = p.SD, CQ = CP / 4;

    extern __shared__ __align__(16) float smem[];
    float* sWt = smem;                          // [DP][KP] transposed weights
    float* stage0 = sWt + DP * KP;              // [NS] x { U [TN][SU] | delta [TN][SD] }
    const size_t STAGE = (size_t)TN * (SU + SD);
    float* ybuf = stage0 + NS * STAGE;          // [2][TN][DP]: y_t of the tile being formed and of the next one
    __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS];

    for (int i = tid * 4; i < DP * KP; i += BL_NT * 4) st4(sWt + i, ldg4(p.wpack + net.wt_off[0] + i));
    if (tid == 0) {
        // full: every producer thread arrives once (its copies have landed and its delta pieces are stored); empty: one lane per compute warp
        for (int i = 0; i < NS; ++i) { mbar_init(&bar_full[i], BL_PRODUCE); mbar_init(&bar_empty[i], BL_COMPUTE / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (p.N + TN - 1) / TN;
    const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    float* red = stage0;                              // after the tiles: [8][RROWS][DP] copies of dW, then [BL_PRODUCE][4] pieces of db

    if (warp >= 16) {
        // ================================================== PRODUCERS ==========================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BL_REGS_PRODUCE));
        const int pt = tid - BL_COMPUTE, lig = pt % LPN, act = net.act[0];
        const bool bn_train = p.bn_stats != nullptr;
        float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
        float bnc[5][4];        // training BN: mean, 1/std, a, mean(G), mean(G xhat) of my 4 columns; otherwise a of the final affine
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = 4 * lig + c;
            bnc[2][c] = bn_train ? p.bn_stats[2 * DP + j] : p.wpack[net.aff_off + j];
            bnc[0][c] = bn_train ? p.bn_stats[j] : 0.f;
            bnc[1][c] = bn_train ? rsqrtf(p.bn_stats[DP + j] + p.bn_eps) : 0.f;
            bnc[3][c] = bn_train ? p.bn_sums[j] : 0.f;
            bnc[4][c] = bn_train ? p.bn_sums[DP + j] : 0.f;
        }
        // All five inputs of a tile travel by cp.async, ONE TILE AHEAD of the delta formation: u_t = [x_t | A x_t | cst] into the stage,
        // dL/dx_{t+1} into the delta columns of the stage (delta is formed in place), y_t into a two-tile scratch.  A thread forms
        // delta exactly for the pieces it copied itself, so cp.async.wait_group on its own groups is all the ordering it needs; the
        // mbarrier arrive (release) then publishes copies and delta to the arithmetic warps.  (Loading dL/dx and y through
        // registers inside the tile exposed one DRAM latency per tile: 4.9 us per tile whatever the width.)
        auto issue = [&](int it) {
            const int s = it % NS;
            float* U = stage0 + (size_t)s * STAGE;
            float* Dl = U + TN * SU;
            float* Yb = ybuf + (size_t)(it & 1) * TN * DP;
            const long long n0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            for (int item = pt; item < TN * LPN; item += BL_PRODUCE) {
                const int i = item / LPN;
                const bool ok = i < nvalid;
                const size_t off = (size_t)(n0 + (ok ? i : 0)) * DP + 4 * lig;
                cp_async16_zfill(U + i * SU + 4 * lig, p.x_t + off, ok);
                cp_async16_zfill(U + i * SU + DP + 4 * lig, p.agg_t + off, ok);
                cp_async16_zfill(Dl + i * SD + 4 * lig, p.G + off, ok);
                cp_async16_zfill(Yb + i * DP + 4 * lig, p.y_t + off, ok);
            }
            for (int item = pt; item < TN * CQ; item += BL_PRODUCE) {
                const int i = item / CQ, c = item % CQ;
                const bool ok = i < nvalid;
                cp_async16_zfill(U + i * SU + 2 * DP + 4 * c, p.cst + (size_t)(n0 + (ok ? i : 0)) * CP + 4 * c, ok);
            }
            cp_async_commit();
        };
        auto form = [&](auto act_c) {
            constexpr int ACT = decltype(act_c)::value;
            if (my_tiles > 0) issue(0);
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % NS;
                if (it + 1 < my_tiles) {
                    if (it + 1 >= NS) mbar_wait<64>(&bar_empty[(it + 1) % NS], (((it + 1) / NS) - 1) & 1);
                    issue(it + 1);
                    cp_async_wait_group<1>();       // my copies of tile it have landed (those of tile it + 1 stay in flight)
                } else {
                    cp_async_wait_group<0>();
                }
                float* Dl = stage0 + (size_t)s * STAGE + TN * SU;
                const float* Yb = ybuf + (size_t)(it & 1) * TN * DP;
                const long long n0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TN;
                const int nvalid = (int)min((long long)TN, p.N - n0);
                for (int item = pt; item < TN * LPN; item += BL_PRODUCE) {
                    const int i = item / LPN;
                    const float4 g4 = ld4(Dl + i * SD + 4 * lig), y4 = ld4(Yb + i * DP + 4 * lig);
                    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, yy[4] = {y4.x, y4.y, y4.z, y4.w};
                    float r[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        // training BN: dL/dh = a (G - mean(G) - xhat mean(G xhat)),  xhat = (h - mean) / std;  otherwise the final affine a
                        const float xhat = (yy[c] - bnc[0][c]) * bnc[1][c];
                        const float gy = bnc[2][c] * (gg[c] - bnc[3][c] - xhat * bnc[4][c]);
                        r[c] = (i < nvalid && 4 * lig + c < D) ? gy * act_grad_from_output(ACT, yy[c]) : 0.f;
                        dbacc[c] += r[c];
                    }
                    st4(Dl + i * SD + 4 * lig, make_float4(r[0], r[1], r[2], r[3]));
                }
                mbar_arrive(&bar_full[s]);
            }
        };
        switch (act) {
            case GNN_ACT_RELU: form(std::integral_constant<int, GNN_ACT_RELU>{}); break;
            case GNN_ACT_TANH: form(std::integral_constant<int, GNN_ACT_TANH>{}); break;
            case GNN_ACT_SIGMOID: form(std::integral_constant<int, GNN_ACT_SIGMOID>{}); break;
            case GNN_ACT_SELU: form(std::integral_constant<int, GNN_ACT_SELU>{}); break;
            case GNN_ACT_ELU: form(std::integral_constant<int, GNN_ACT_ELU>{}); break;
            case GNN_ACT_SOFTPLUS: form(std::integral_constant<int, GNN_ACT_SOFTPLUS>{}); break;
            default: form(std::integral_constant<int, GNN_ACT_LINEAR>{}); break;
        }
        __syncthreads();          // every role is done with the stages
        st4(red + (size_t)8 * RROWS * DP + 4 * pt, make_float4(dbacc[0], dbacc[1], dbacc[2], dbacc[3]));
    } else if (warp < 8) {
        // ================================= dW += u^T delta over my 16 nodes of every tile ====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BL_REGS_DW));
        const int kb = lane >> 2, jb = lane & 3;
        float accW[MK][MJ], accC[2][MJ];
#pragma unroll
        for (int r = 0; r < MK; ++r)
#pragma unroll
            for (int c = 0; c < MJ; ++c) accW[r][c] = 0.f;
#pragma unroll
        for (int c = 0; c < MJ; ++c) accC[0][c] = accC[1][c] = 0.f;
        const bool wide_cst = CP > 8;
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % NS;
            mbar_wait<32>(&bar_full[s], (it / NS) & 1);
            const float* U = stage0 + (size_t)s * STAGE;
            const float* u = U + (size_t)(16 * warp) * SU;
            const float* d = U + TN * SU + (size_t)(16 * warp) * SD + MJ * jb;
#pragma unroll 2
            for (int n = 0; n < 16; ++n, u += SU, d += SD) {
                float a[MK], dv[MJ];
#pragma unroll
                for (int q = 0; q < MK / 4; ++q) {
                    const float4 v = ld4(u + MK * kb + 4 * q);
                    a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int q = 0; q < MJ / 4; ++q) {
                    const float4 v = ld4(d + 4 * q);
                    dv[4 * q] = v.x; dv[4 * q + 1] = v.y; dv[4 * q + 2] = v.z; dv[4 * q + 3] = v.w;
                }
                const float c0 = u[2 * DP + kb];          // constant rows (the columns past CP belong to the delta tile: rows ignored below)
#pragma unroll
                for (int r = 0; r < MK; ++r)
#pragma unroll
                    for (int c = 0; c < MJ; c += 2) fma2(accW[r][c], accW[r][c + 1], a[r], a[r], dv[c], dv[c + 1]);
#pragma unroll
                for (int c = 0; c < MJ; c += 2) fma2(accC[0][c], accC[0][c + 1], c0, c0, dv[c], dv[c + 1]);
                if (wide_cst) {
                    const float c1 = u[2 * DP + 8 + kb];
#pragma unroll
                    for (int c = 0; c < MJ; c += 2) fma2(accC[1][c], accC[1][c + 1], c1, c1, dv[c], dv[c + 1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[s]);
        }
        __syncthreads();
        float* mine = red + (size_t)warp * RROWS * DP;
#pragma unroll
        for (int r = 0; r < MK; ++r)
#pragma unroll
            for (int c = 0; c < MJ; ++c) mine[(MK * kb + r) * DP + MJ * jb + c] = accW[r][c];
#pragma unroll
        for (int c = 0; c < MJ; ++c) {
            mine[(2 * DP + kb) * DP + MJ * jb + c] = accC[0][c];
            mine[(2 * DP + 8 + kb) * DP + MJ * jb + c] = accC[1][c];
        }
    } else {
        // ============================ g_u = delta W^T: my two nodes, my quarter of the 2 DP columns ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BL_REGS_GU));
        constexpr int QW = DP / 2;                     // columns per quarter
        const int g = warp - 8, qtr = g >> 1, i0 = 32 * (g & 1) + lane, i1 = i0 + 64;
        float* dst = (qtr >= 2 ? p.GA : p.GS) + (qtr & 1) * QW;
        const float* wbase = sWt + qtr * QW;
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % NS;
            mbar_wait<32>(&bar_full[s], (it / NS) & 1);
            const float* U = stage0 + (size_t)s * STAGE;
            const float* Dl = U + TN * SU;
            const long long n0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TN;
            const int nvalid = (int)min((long long)TN, p.N - n0);
            float g0[QW], g1[QW];
#pragma unroll
            for (int k = 0; k < QW; ++k) g0[k] = g1[k] = 0.f;
            const float* d0 = Dl + (size_t)i0 * SD;
            const float* d1 = Dl + (size_t)i1 * SD;
#pragma unroll 2
            for (int j4 = 0; j4 < DP; j4 += 4) {
                const float4 e0 = ld4(d0 + j4), e1 = ld4(d1 + j4);
                const float dj0[4] = {e0.x, e0.y, e0.z, e0.w}, dj1[4] = {e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* wrow = wbase + (size_t)(j4 + c) * KP;
#pragma unroll
                    for (int kq = 0; kq < QW / 4; ++kq) {
                        const float4 w4 = ld4(wrow + 4 * kq);
                        fma2(g0[4 * kq], g0[4 * kq + 1], dj0[c], dj0[c], w4.x, w4.y);
                        fma2(g0[4 * kq + 2], g0[4 * kq + 3], dj0[c], dj0[c], w4.z, w4.w);
                        fma2(g1[4 * kq], g1[4 * kq + 1], dj1[c], dj1[c], w4.x, w4.y);
                        fma2(g1[4 * kq + 2], g1[4 * kq + 3], dj1[c], dj1[c], w4.z, w4.w);
                    }
                }
            }
            const bool scaled = qtr >= 2 && p.row_scale_mode;
            const float sc0 = scaled ? U[(size_t)i0 * SU + 2 * DP + net.C] : 1.f;
            const float sc1 = scaled ? U[(size_t)i1 * SU + 2 * DP + net.C] : 1.f;
            if (p.gcst && qtr == 0) {
                // gradient of the constant rows (only when the caller asks for label gradients): columns 2 DP .. 2 DP + CP - 1
                for (int which = 0; which < 2; ++which) {
                    const int i = which ? i1 : i0;
                    if (i >= nvalid) continue;
                    const float* dr = Dl + (size_t)i * SD;
                    for (int cq = 0; cq < CQ; ++cq) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int j = 0; j < DP; ++j) acc = fma4(dr[j], ld4(sWt + (size_t)j * KP + 2 * DP + 4 * cq), acc);
                        float* gd = p.gcst + (size_t)(n0 + i) * CP + 4 * cq;
                        st4(gd, add4(ld4(gd), acc));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[s]);       // the stage is free: the stores below come out of registers
#pragma unroll
            for (int q = 0; q < QW / 8; ++q) {
                float v[8];
                if (i0 < nvalid) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = g0[8 * q + e] * sc0;
                    stg8(dst + (size_t)(n0 + i0) * DP + 8 * q, v);
                }
                if (i1 < nvalid) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = g1[8 * q + e] * sc1;
                    stg8(dst + (size_t)(n0 + i1) * DP + 8 * q, v);
                }
            }
        }
        __syncthreads();
    }

    // ---- the eight register copies of dW and the db pieces, added in a fixed order into this CTA's partial slot ---------------
    __syncthreads();
    const float* redb = red + (size_t)8 * RROWS * DP;       // [BL_PRODUCE][4]
    float* slot = p.gpartial + (size_t)blockIdx.x * net.fwd_floats;
    for (int idx = tid; idx < KP * DP; idx += BL_NT) {      // rows of W in packed order: k * DP + j, k < KP
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[(size_t)w * RROWS * DP + idx];
        slot[net.w_off[0] + idx] += sum;
    }
    if (tid < DP) {                                         // column tid: the producer threads with pt % LPN == tid / 4 hold it, in thread order
        float sum = 0.f;
        for (int th = tid / 4; th < BL_PRODUCE; th += LPN) sum += redb[4 * th + (tid & 3)];
        slot[net.b_off[0] + tid] += sum;
    }
}

}  // namespace gnn
