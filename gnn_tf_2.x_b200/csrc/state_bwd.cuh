// state_bwd.cuh -- BPTT kernels of the state-convergence loop (what tf.GradientTape replays for
// GNN/GNN.py:223-242, reference GNN/GNN_BaseClass.py:233-237).
//
// Iteration t of the forward loop computed  u_t = [x_t | A x_t | cst],  y_t = net_state(u_t),
// x_{t+1} = y_t (or BatchNormalization(y_t)).  Given G = dL/dx_{t+1} the backward step is
//   node kernel    : per node, back through BatchNormalization and the Dense chain (hidden activations are
//                    recomputed from the saved u_t, the last output is read back), accumulating dW/db per CTA
//                    in shared memory (no atomics; per-CTA partials are reduced in a fixed order at the end)
//                    and producing g_self = dL/du[state part], g_agg = dL/du[agg part], g_cst;
//   scatter kernel : dL/dx_t[u] = g_self[u] + sum over arcs u->n of w(u->n) * g_agg[n], a gather over the
//                    source-sorted CSR^T with the same 128-bit lane mapping as the forward gather.
// Every kernel returns at once when t >= k (device-side iteration count), so the host enqueues max_iter steps.
#pragma once
#include "state_fwd.cuh"

namespace gnn {

// floats of one per-CTA partial gradient buffer: same layout as the packed forward net
// (dW, db per layer at w_off / b_off; dgamma, dbeta in the affine slot)
static inline size_t bwd_param_floats(const NetLayout& lay) { return (size_t)lay.fwd_floats; }

struct BwdNodeParams {
    long long N;
    const float* G;          // [N, DP] dL/dx_{t+1}
    const float* x_t;        // [N, DP]
    const float* agg_t;      // [N, DP]
    const float* cst;        // [N, CP]
    const float* y_t;        // [N, DP] output of the last Dense (+dropout) = x_{t+1} without BN, h_t with training BN;
                             // NULL => recompute the last layer too
    const float* wpack;
    float* GS;               // [N, DP]
    float* GA;               // [N, DP]
    float* gcst;             // [N, CP] accumulated, or NULL
    float* gpartial;         // [gridDim.x][fwd_floats] accumulated across launches
    const int* k_ptr;
    int t;
    uint32_t seed;
    const uint32_t* seed_dev;   // when set, the dropout seed of the call is read from device memory (CUDA-graph replays)
    int training;
    int row_scale_mode;      // fold cst[:, C] into GA
    // training-mode BatchNormalization
    const float* bn_stats;   // [4][DP] of iteration t (mean, var, a, c) or NULL
    const float* bn_sums;    // [2][DP]: mean_n(G), mean_n(G * xhat)
    float bn_eps;
    NetLayout net;
    int SU, SD;              // strides of the input tile and of the delta tiles
    int act_off[GNN_MAX_LAYERS];   // float offset of the saved activation tile of hidden layer l (l >= 1)
    int act_stride[GNN_MAX_LAYERS];
    int has_dB;              // second delta tile allocated (L >= 2 or y is recomputed)
};

// dW[k, j] += sum_n a[n, k] * d[n, j] over the tile; 4x4 micro-tiles, results added to the CTA accumulators
template <int TN, int NT>
__device__ __forceinline__ void outer_accumulate(const float* __restrict__ a, int a_stride, int KP,
                                                 const float* __restrict__ d, int d_stride, int HP,
                                                 float* __restrict__ sGw) {
    const int KQ = KP >> 2, JQ = HP >> 2;
    for (int item = threadIdx.x; item < KQ * JQ; item += NT) {
        const int jq = item % JQ, kq = item / JQ;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        const float* ap = a + 4 * kq;
        const float* dp = d + 4 * jq;
#pragma unroll 4
        for (int n = 0; n < TN; ++n) {
            const float4 av = ld4(ap + n * a_stride);
            const float4 dv = ld4(dp + n * d_stride);
            acc[0][0] = fmaf(av.x, dv.x, acc[0][0]); acc[0][1] = fmaf(av.x, dv.y, acc[0][1]);
            acc[0][2] = fmaf(av.x, dv.z, acc[0][2]); acc[0][3] = fmaf(av.x, dv.w, acc[0][3]);
            acc[1][0] = fmaf(av.y, dv.x, acc[1][0]); acc[1][1] = fmaf(av.y, dv.y, acc[1][1]);
            acc[1][2] = fmaf(av.y, dv.z, acc[1][2]); acc[1][3] = fmaf(av.y, dv.w, acc[1][3]);
            acc[2][0] = fmaf(av.z, dv.x, acc[2][0]); acc[2][1] = fmaf(av.z, dv.y, acc[2][1]);
            acc[2][2] = fmaf(av.z, dv.z, acc[2][2]); acc[2][3] = fmaf(av.z, dv.w, acc[2][3]);
            acc[3][0] = fmaf(av.w, dv.x, acc[3][0]); acc[3][1] = fmaf(av.w, dv.y, acc[3][1]);
            acc[3][2] = fmaf(av.w, dv.z, acc[3][2]); acc[3][3] = fmaf(av.w, dv.w, acc[3][3]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float* dst = sGw + (4 * kq + r) * HP + 4 * jq;
            float4 cur = ld4(dst);
            cur.x += acc[r][0]; cur.y += acc[r][1]; cur.z += acc[r][2]; cur.w += acc[r][3];
            st4(dst, cur);
        }
    }
}

// db[j] += sum_n d[n, j]
template <int TN, int NT>
__device__ __forceinline__ void column_accumulate(const float* __restrict__ d, int d_stride, int HP, float* __restrict__ sGb) {
    for (int j = threadIdx.x; j < HP; j += NT) {
        float s = 0.f;
        for (int n = 0; n < TN; ++n) s += d[n * d_stride + j];
        sGb[j] += s;
    }
}

template <int DP, int TN, int NT>
static __global__ void __launch_bounds__(NT) state_bwd_node_kernel(const BwdNodeParams p) {
    constexpr int LPN = DP / 4;
    constexpr int NG = TN / 8;
    if (p.t >= *reinterpret_cast<const volatile int*>(p.k_ptr)) return;

    const NetLayout& net = p.net;
    const int tid = threadIdx.x, L = net.L;
    const int CP = net.CP, KP = net.KP, D = net.D, SU = p.SU, SD = p.SD;

    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                                   // forward weights (+ affine)
    float* sWt = sW + net.fwd_floats;                   // transposed weights, packed from wt_off[0]
    float* sG = sWt + (net.total_floats - net.fwd_floats);  // CTA gradient accumulators (fwd layout)
    float* bufU = sG + net.fwd_floats;                  // [TN][SU] input tile, later g_u
    float* dA = bufU + TN * SU;                         // [TN][SD]
    float* dB = dA + TN * SD;                           // [TN][SD] (L >= 2)
    float* acts = dB + (p.has_dB ? TN * SD : 0);        // saved hidden activations

    for (int i = tid * 4; i < net.total_floats; i += NT * 4) st4(sW + i, ldg4(p.wpack + i));
    for (int i = tid; i < net.fwd_floats; i += NT) sG[i] = 0.f;

    const bool drop_in = p.training && net.drop[0] > 0.f;
    const uint32_t key_in = dropout_key(call_seed(p), 0u, (uint32_t)p.t);
    const float scale_in = drop_in ? 1.f / (1.f - net.drop[0]) : 1.f;
    const int F_in = net.in_dim[0];
    const float* aff_a = sW + net.aff_off;
    const bool bn_train = p.bn_stats != nullptr;

    const long long ntiles = (p.N + TN - 1) / TN;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n0 = tile * TN;
        const int nvalid = (int)min((long long)TN, p.N - n0);
        __syncthreads();

        // ---- load u_t (with the forward's input dropout) and dL/dy into shared memory --------------------
        for (int item = tid; item < TN * LPN; item += NT) {
            const int i = item / LPN, lig = item % LPN;
            float* rowU = bufU + i * SU;
            float4 own = make_float4(0.f, 0.f, 0.f, 0.f), agg = own, gy = own;
            if (i < nvalid) {
                const long long n = n0 + i;
                own = ldg4(p.x_t + (size_t)n * DP + 4 * lig);
                agg = ldg4(p.agg_t + (size_t)n * DP + 4 * lig);
                const float4 g = ldg4(p.G + (size_t)n * DP + 4 * lig);
                if (bn_train) {
                    // dL/dh = a * (G - mean(G) - xhat * mean(G*xhat)),  xhat = (h - mean) * rsqrt(var + eps)
                    const float4 h = ldg4(p.y_t + (size_t)n * DP + 4 * lig);
                    const float hh[4] = {h.x, h.y, h.z, h.w}, gg[4] = {g.x, g.y, g.z, g.w};
                    float r[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = 4 * lig + c;
                        const float mean = p.bn_stats[j], var = p.bn_stats[DP + j], a = p.bn_stats[2 * DP + j];
                        const float xhat = (hh[c] - mean) * rsqrtf(var + p.bn_eps);
                        r[c] = j < D ? a * (gg[c] - p.bn_sums[j] - xhat * p.bn_sums[DP + j]) : 0.f;
                    }
                    gy = make_float4(r[0], r[1], r[2], r[3]);
                } else {
                    const float4 a = ld4(aff_a + 4 * lig);  // inference BatchNormalization: x = a*y + c
                    gy = make_float4(a.x * g.x, a.y * g.y, a.z * g.z, a.w * g.w);
                }
                if (drop_in) {
                    const uint64_t rbase = (uint64_t)n * (uint64_t)F_in;
                    float* o = reinterpret_cast<float*>(&own);
                    float* a = reinterpret_cast<float*>(&agg);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = 4 * lig + c;
                        if (j < D) {
                            o[c] = drop1(o[c], true, key_in, rbase + j, net.drop[0], scale_in);
                            a[c] = drop1(a[c], true, key_in, rbase + D + net.NL_self + j, net.drop[0], scale_in);
                        }
                    }
                }
            }
            st4(rowU + 4 * lig, own);
            st4(rowU + DP + 4 * lig, agg);
            st4(dA + i * SD + 4 * lig, gy);
        }
        for (int item = tid; item < TN * (CP / 4); item += NT) {
            const int i = item / (CP / 4), c = item % (CP / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nvalid) {
                const long long n = n0 + i;
                v = ldg4(p.cst + (size_t)n * CP + 4 * c);
                if (drop_in) {
                    const uint64_t rbase = (uint64_t)n * (uint64_t)F_in;
                    float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int kc = keras_input_col(2 * DP + 4 * c + q, D, DP, net.NL_self, net.NL_agg, net.AL);
                        if (kc >= 0) vv[q] = drop1(vv[q], true, key_in, rbase + kc, net.drop[0], scale_in);
                    }
                }
            }
            st4(bufU + i * SU + 2 * DP + 4 * c, v);
        }
        __syncthreads();

        // ---- recompute the hidden activations a'_1 .. a'_{L-1} (and y when it was not saved) ---------------
        const bool need_y = (p.y_t == nullptr);
        float* ybuf = dB;  // only used when need_y (then L >= 1 and dB is allocated by the host for this case)
        for (int l = 0; l < L; ++l) {
            const bool last = (l == L - 1);
            if (last && !need_y) break;
            const float* in = (l == 0) ? bufU : acts + p.act_off[l];
            const int in_stride = (l == 0) ? SU : p.act_stride[l];
            float* out = last ? ybuf : acts + p.act_off[l + 1];
            const int out_stride = last ? SD : p.act_stride[l + 1];
            const float* bias = sW + net.b_off[l];
            const int act = net.act[l], odim = net.out_dim[l];
            const float rate = net.drop[l + 1];
            const bool drop_here = p.training && rate > 0.f;
            const uint32_t key = dropout_key(call_seed(p), (uint32_t)(l + 1), (uint32_t)p.t);
            const float dscale = drop_here ? 1.f / (1.f - rate) : 1.f;
            auto epi = [&](int ng, int cg, float (&acc)[8][4]) {
                const float4 b4 = ld4(bias + 4 * cg);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = ng + NG * i;
                    float v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = 4 * cg + c;
                        float y = act_apply(act, acc[i][c] + bb[c]);
                        if (drop_here) y = drop1(y, j < odim, key, (uint64_t)(n0 + row) * (uint64_t)odim + j, rate, dscale);
                        v[c] = (j < odim) ? y : 0.f;
                    }
                    st4(out + row * out_stride + 4 * cg, make_float4(v[0], v[1], v[2], v[3]));
                }
            };
            dense_tile<TN, NT>(in, in_stride, net.in_pad[l], sW + net.w_off[l], net.out_pad[l], epi);
            __syncthreads();
        }

        // ---- delta of the last layer: dA <- dL/dy * dropout_L * act'(y) -----------------------------------
        {
            const int l = L - 1;
            const int act = net.act[l], odim = net.out_dim[l];
            const float rate = net.drop[L];
            const bool drop_here = p.training && rate > 0.f;
            const uint32_t key = dropout_key(call_seed(p), (uint32_t)L, (uint32_t)p.t);
            const float dscale = drop_here ? 1.f / (1.f - rate) : 1.f;
            for (int item = tid; item < TN * LPN; item += NT) {
                const int i = item / LPN, lig = item % LPN;
                float4 g = ld4(dA + i * SD + 4 * lig);
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < nvalid) y = need_y ? ld4(ybuf + i * SD + 4 * lig) : ldg4(p.y_t + (size_t)(n0 + i) * DP + 4 * lig);
                float* gg = reinterpret_cast<float*>(&g);
                const float* yy = reinterpret_cast<const float*>(&y);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = 4 * lig + c;
                    float yv = yy[c], m = 1.f;
                    if (drop_here) {
                        const bool keep = dropout_keep(key, (uint64_t)(n0 + i) * (uint64_t)odim + j, rate);
                        m = keep ? dscale : 0.f;
                        yv = yv * (1.f - rate);  // value before the dropout where it was kept
                    }
                    gg[c] = (i < nvalid && j < odim) ? gg[c] * m * act_grad_from_output(act, yv) : 0.f;
                }
                st4(dA + i * SD + 4 * lig, g);
            }
        }
        __syncthreads();

        // ---- back through the Dense chain ----------------------------------------------------------------
        float* dcur = dA;
        float* dnext = dB;
        for (int l = L - 1; l >= 0; --l) {
            const float* a_in = (l == 0) ? bufU : acts + p.act_off[l];
            const int a_stride = (l == 0) ? SU : p.act_stride[l];
            outer_accumulate<TN, NT>(a_in, a_stride, net.in_pad[l], dcur, SD, net.out_pad[l], sG + net.w_off[l]);
            column_accumulate<TN, NT>(dcur, SD, net.out_pad[l], sG + net.b_off[l]);
            __syncthreads();  // bufU (l == 0) is about to be overwritten by g_u
            const float* Wt = sWt + (net.wt_off[l] - net.fwd_floats);
            if (l > 0) {
                // delta_{l-1} = (delta_l W_l^T) * dropout_l * act'_{l-1}(a_l)
                const int act = net.act[l - 1], idim = net.in_dim[l];
                const float rate = net.drop[l];
                const bool drop_here = p.training && rate > 0.f;
                const uint32_t key = dropout_key(call_seed(p), (uint32_t)l, (uint32_t)p.t);
                const float dscale = drop_here ? 1.f / (1.f - rate) : 1.f;
                auto epi = [&](int ng, int cg, float (&acc)[8][4]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = ng + NG * i;
                        const float4 av = ld4(a_in + row * a_stride + 4 * cg);
                        const float aa[4] = {av.x, av.y, av.z, av.w};
                        float v[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int j = 4 * cg + c;
                            float m = 1.f, yv = aa[c];
                            if (drop_here) {
                                const bool keep = dropout_keep(key, (uint64_t)(n0 + row) * (uint64_t)idim + j, rate);
                                m = keep ? dscale : 0.f;
                                yv = yv * (1.f - rate);
                            }
                            v[c] = (j < idim && row < nvalid) ? acc[i][c] * m * act_grad_from_output(act, yv) : 0.f;
                        }
                        st4(dnext + row * SD + 4 * cg, make_float4(v[0], v[1], v[2], v[3]));
                    }
                };
                dense_tile<TN, NT>(dcur, SD, net.out_pad[l], Wt, net.in_pad[l], epi);
                __syncthreads();
                float* tmp = dcur; dcur = dnext; dnext = tmp;
            } else {
                // g_u = (delta_0 W_0^T) * dropout_0  -> bufU
                auto epi = [&](int ng, int cg, float (&acc)[8][4]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = ng + NG * i;
                        float v[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float m = 1.f;
                            if (drop_in) {
                                const int kc = keras_input_col(4 * cg + c, D, DP, net.NL_self, net.NL_agg, net.AL);
                                m = (kc >= 0 && dropout_keep(key_in, (uint64_t)(n0 + row) * (uint64_t)F_in + kc, net.drop[0])) ? scale_in : 0.f;
                            }
                            v[c] = acc[i][c] * m;
                        }
                        st4(bufU + row * SU + 4 * cg, make_float4(v[0], v[1], v[2], v[3]));
                    }
                };
                dense_tile<TN, NT>(dcur, SD, net.out_pad[0], Wt, KP, epi);
                __syncthreads();
            }
        }

        // ---- write g_self, g_agg (scaled by the row weight), accumulate g_cst ---------------------------------
        for (int item = tid; item < TN * LPN; item += NT) {
            const int i = item / LPN, lig = item % LPN;
            if (i >= nvalid) continue;
            const long long n = n0 + i;
            const float4 gs = ld4(bufU + i * SU + 4 * lig);
            float4 ga = ld4(bufU + i * SU + DP + 4 * lig);
            if (p.row_scale_mode) {
                const float s = __ldg(p.cst + (size_t)n * CP + net.C);
                ga.x *= s; ga.y *= s; ga.z *= s; ga.w *= s;
            }
            st4(p.GS + (size_t)n * DP + 4 * lig, gs);
            st4(p.GA + (size_t)n * DP + 4 * lig, ga);
        }
        if (p.gcst) {
            for (int item = tid; item < TN * (CP / 4); item += NT) {
                const int i = item / (CP / 4), c = item % (CP / 4);
                if (i >= nvalid) continue;
                float* dst = p.gcst + (size_t)(n0 + i) * CP + 4 * c;
                const float4 g = ld4(bufU + i * SU + 2 * DP + 4 * c);
                float4 cur = ld4(dst);
                cur.x += g.x; cur.y += g.y; cur.z += g.z; cur.w += g.w;
                st4(dst, cur);
            }
        }
    }

    // ---- add the CTA accumulators to this CTA's partial slot (fixed CTA -> deterministic) -------------------
    __syncthreads();
    float* slot = p.gpartial + (size_t)blockIdx.x * net.fwd_floats;
    for (int i = tid * 4; i < net.fwd_floats; i += NT * 4) {
        float4 cur = ld4(slot + i);
        const float4 add = ld4(sG + i);
        cur.x += add.x; cur.y += add.y; cur.z += add.z; cur.w += add.w;
        st4(slot + i, cur);
    }
}

// dL/dx_t[u] = GS[u] + sum_{arcs u -> n} w * GA[n]   (source-sorted CSR^T)
template <int DP, bool HAS_VAL>
static __global__ void state_bwd_scatter_kernel(const int* __restrict__ k_ptr, int t, const int32_t* __restrict__ rowptr_T,
                                         const int32_t* __restrict__ col_T, const float* __restrict__ val_T, long long N,
                                         const float* __restrict__ GS, const float* __restrict__ GA, float* __restrict__ G) {
    if (t >= *reinterpret_cast<const volatile int*>(k_ptr)) return;
    constexpr int LPN = DP / 4;
    const long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long u = item / LPN;
    const int lig = (int)(item % LPN);
    if (u >= N) return;
    const int e0 = __ldg(rowptr_T + u), e1 = __ldg(rowptr_T + u + 1);
    // batches of 8 arcs: all indices, then all 8 row loads back to back (nothing with a scoreboard in between), then the sum in
    // stored order; past the end the last arc is repeated (its row is loaded, not summed) so that the batch stays branch-free
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; e += 8) {
        int idx[8];
        float w[8];
        float4 r[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int ee = min(e + b, e1 - 1);
            idx[b] = __ldg(col_T + ee);
            if (HAS_VAL) w[b] = __ldg(val_T + ee);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) r[b] = ldg4(GA + (size_t)idx[b] * DP + 4 * lig);
#pragma unroll
        for (int b = 0; b < 8; ++b)
            if (e + b < e1) acc = HAS_VAL ? fma4(w[b], r[b], acc) : add4(acc, r[b]);
    }
    const float4 gs = ldg4(GS + (size_t)u * DP + 4 * lig);
    st4(G + (size_t)u * DP + 4 * lig, add4(acc, gs));
}

// per-column partial sums of G and G * xhat over the nodes (training-mode BatchNormalization backward)
template <int DP>
static __global__ void bn_bwd_reduce_kernel(const int* __restrict__ k_ptr, int t, const float* __restrict__ G, const float* __restrict__ h,
                                     const float* __restrict__ stats, float eps, long long N, double* __restrict__ partial) {
    if (t >= *reinterpret_cast<const volatile int*>(k_ptr)) return;
    constexpr int LPN = DP / 4;
    const int lig = threadIdx.x % LPN;
    double s1[4] = {0., 0., 0., 0.}, s2[4] = {0., 0., 0., 0.};
    float mean[4], inv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { mean[c] = stats[4 * lig + c]; inv[c] = rsqrtf(stats[DP + 4 * lig + c] + eps); }
    const long long rows_per_pass = (long long)gridDim.x * (blockDim.x / LPN);
    for (long long n = blockIdx.x * (long long)(blockDim.x / LPN) + threadIdx.x / LPN; n < N; n += rows_per_pass) {
        const float4 g = ldg4(G + (size_t)n * DP + 4 * lig), hv = ldg4(h + (size_t)n * DP + 4 * lig);
        const float gg[4] = {g.x, g.y, g.z, g.w}, hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) { s1[c] += gg[c]; s2[c] += (double)gg[c] * (double)((hh[c] - mean[c]) * inv[c]); }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        for (int off = LPN; off < 32; off <<= 1) {
            s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], off);
            s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], off);
        }
    __shared__ double red[8 * 32 * 8];  // [warps <= 8][LPN <= 32][8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < LPN)
#pragma unroll
        for (int c = 0; c < 4; ++c) { red[(warp * LPN + lane) * 8 + c] = s1[c]; red[(warp * LPN + lane) * 8 + 4 + c] = s2[c]; }
    __syncthreads();
    if (threadIdx.x < LPN) {
        double a1[4] = {0., 0., 0., 0.}, a2[4] = {0., 0., 0., 0.};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
#pragma unroll
            for (int c = 0; c < 4; ++c) { a1[c] += red[(w * LPN + threadIdx.x) * 8 + c]; a2[c] += red[(w * LPN + threadIdx.x) * 8 + 4 + c]; }
        double* dst = partial + (size_t)blockIdx.x * 2 * DP;
#pragma unroll
        for (int c = 0; c < 4; ++c) { dst[4 * threadIdx.x + c] = a1[c]; dst[DP + 4 * threadIdx.x + c] = a2[c]; }
    }
}

// sums -> means used by the node kernel; dgamma += sum(G*xhat), dbeta += sum(G) (accumulated over iterations)
// one CTA of 64 threads per column: strided partial sums, then a fixed shuffle / shared-memory tree (deterministic)
static __global__ void bn_bwd_finalize_kernel(const int* __restrict__ k_ptr, int t, const double* __restrict__ partial, int nblocks, int DP,
                                       int D, long long N, float* __restrict__ sums, double* __restrict__ dgamma_dbeta) {
    if (t >= *reinterpret_cast<const volatile int*>(k_ptr)) return;
    const int j = blockIdx.x;
    double s1 = 0., s2 = 0.;
    if (j < D)
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) { s1 += partial[(size_t)b * 2 * DP + j]; s2 += partial[(size_t)b * 2 * DP + DP + j]; }
    for (int off = 16; off > 0; off >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, off); s2 += __shfl_down_sync(0xffffffffu, s2, off); }
    __shared__ double red[2][2];
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s1; red[threadIdx.x >> 5][1] = s2; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    s1 = red[0][0] + red[1][0];
    s2 = red[0][1] + red[1][1];
    sums[j] = (float)(s1 / (double)N);
    sums[DP + j] = (float)(s2 / (double)N);
    dgamma_dbeta[j] += s2;
    dgamma_dbeta[DP + j] += s1;
}

// G0 = pad(g_x) ; zero-fill helper
static __global__ void pad_rows_kernel(const float* __restrict__ src, long long N, int D, int DP, float* __restrict__ dst) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= N * DP) return;
    const long long n = idx / DP;
    const int j = (int)(idx % DP);
    dst[idx] = j < D ? src[n * D + j] : 0.f;
}

static __global__ void unpad_rows_kernel(const float* __restrict__ src, long long N, int D, int DP, float* __restrict__ dst) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= N * D) return;
    dst[idx] = src[(idx / D) * DP + (idx % D)];
}

// g_cst -> g_nodes / g_agg_nodes / g_agg_arcs
static __global__ void unpack_gcst_kernel(const float* __restrict__ gcst, long long N, int NL_self, int NL_agg, int AL, int CP,
                                   float* __restrict__ g_nodes, float* __restrict__ g_agg_nodes, float* __restrict__ g_agg_arcs) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int C = NL_self + NL_agg + AL;
    if (idx >= N * C) return;
    const long long n = idx / C;
    const int q = (int)(idx % C);
    const float v = gcst[n * CP + q];
    if (q < NL_self) { if (g_nodes) g_nodes[n * NL_self + q] = v; }
    else if (q < NL_self + NL_agg) { if (g_agg_nodes) g_agg_nodes[n * NL_agg + (q - NL_self)] = v; }
    else if (g_agg_arcs) g_agg_arcs[n * AL + (q - NL_self - NL_agg)] = v;
}

struct ReduceParams {
    gnn_mlp_grad grad;
    NetLayout lay;
    const float* gpartial;
    int nblocks;
    const double* dgamma_dbeta;  // [2][DP] or NULL
    int state_loop;
};

// sum the per-CTA partials in CTA order and scatter to the Keras layouts
static __global__ void reduce_params_kernel(const ReduceParams p) {
    const NetLayout& l = p.lay;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l.fwd_floats) return;
    float s = 0.f;
    if (idx < l.aff_off)
        for (int b = 0; b < p.nblocks; ++b) s += p.gpartial[(size_t)b * l.fwd_floats + idx];
    for (int i = 0; i < l.L; ++i) {
        const int ip = l.in_pad[i], op = l.out_pad[i];
        if (idx >= l.w_off[i] && idx < l.w_off[i] + ip * op) {
            const int r = (idx - l.w_off[i]) / op, c = (idx - l.w_off[i]) % op;
            int kr = r;
            if (i == 0 && p.state_loop) kr = keras_input_col(r, l.D, l.DP, l.NL_self, l.NL_agg, l.AL);
            else if (r >= l.in_dim[i]) kr = -1;
            if (kr >= 0 && c < l.out_dim[i] && p.grad.dW[i]) p.grad.dW[i][(size_t)kr * l.out_dim[i] + c] = s;
            return;
        }
        if (idx >= l.b_off[i] && idx < l.b_off[i] + op) {
            const int c = idx - l.b_off[i];
            if (c < l.out_dim[i] && p.grad.db[i]) p.grad.db[i][c] = s;
            return;
        }
    }
    // affine slot: dgamma | dbeta
    const int op = l.out_pad[l.L - 1], od = l.out_dim[l.L - 1];
    const int j = (idx - l.aff_off) % op;
    const bool is_gamma = (idx - l.aff_off) < op;
    if (j >= od) return;
    const float v = p.dgamma_dbeta ? (float)p.dgamma_dbeta[(is_gamma ? 0 : op) + j] : 0.f;
    if (is_gamma) { if (p.grad.dgamma) p.grad.dgamma[j] = v; }
    else if (p.grad.dbeta) p.grad.dbeta[j] = v;
}

}  // namespace gnn
