// state_aux.cuh -- small kernels around the fused iteration kernel: weight packing, constant-row packing,
// loop initialisation (condition of iteration 0), BatchNormalization batch statistics, final state copy.
#pragma once
#include "net_layout.cuh"

namespace gnn {

// loop control block at the start of the workspace (ints): go[0..max_iter] then k
struct LoopCtl {
    int* go;
    int* k;
};

struct PackParams {
    gnn_mlp net;       // raw Keras-layout pointers
    NetLayout lay;
    float* wpack;      // [lay.total_floats]
    int state_loop;
    int bn_inference;  // fold moving statistics into the final affine
};

// one thread per packed element (weights, biases, transposed weights, affine)
static __global__ void pack_net_kernel(const PackParams p) {
    const NetLayout& l = p.lay;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l.total_floats) return;
    float v = 0.f;
    bool found = false;
    for (int i = 0; i < l.L && !found; ++i) {
        const int ip = l.in_pad[i], op = l.out_pad[i];
        int r = -1, c = -1;
        if (idx >= l.w_off[i] && idx < l.w_off[i] + ip * op) { r = (idx - l.w_off[i]) / op; c = (idx - l.w_off[i]) % op; }
        else if (idx >= l.wt_off[i] && idx < l.wt_off[i] + ip * op) { c = (idx - l.wt_off[i]) / ip; r = (idx - l.wt_off[i]) % ip; }
        else if (idx >= l.b_off[i] && idx < l.b_off[i] + op) {
            c = idx - l.b_off[i];
            v = c < l.out_dim[i] ? p.net.b[i][c] : 0.f;
            found = true;
            break;
        }
        if (r >= 0) {
            int kr = r;  // Keras row
            if (i == 0 && p.state_loop) kr = keras_input_col(r, l.D, l.DP, l.NL_self, l.NL_agg, l.AL);
            else if (r >= l.in_dim[i]) kr = -1;
            v = (kr >= 0 && c < l.out_dim[i]) ? p.net.W[i][(size_t)kr * l.out_dim[i] + c] : 0.f;
            found = true;
        }
    }
    if (!found) {  // final affine a | c
        const int op = l.out_pad[l.L - 1], od = l.out_dim[l.L - 1];
        const int j = (idx - l.aff_off) % op;
        const bool is_a = (idx - l.aff_off) < op;
        float a = 1.f, c = 0.f;
        if (j >= od) { a = 0.f; c = 0.f; }
        else if (p.bn_inference) {
            a = p.net.bn_gamma[j] * rsqrtf(p.net.bn_moving_var[j] + p.net.bn_eps);
            c = p.net.bn_beta[j] - p.net.bn_moving_mean[j] * a;
        }
        v = is_a ? a : c;
    }
    p.wpack[idx] = v;
}

// cst[n] = [nodes | agg_nodes | agg_arcs | row_scale | 0...]
static __global__ void pack_cst_kernel(const float* __restrict__ nodes, const float* __restrict__ agg_nodes,
                                const float* __restrict__ agg_arcs, const float* __restrict__ row_scale, long long N,
                                int NL_self, int NL_agg, int AL, int CP, float* __restrict__ cst) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= N * CP) return;
    const long long n = idx / CP;
    const int q = (int)(idx % CP);
    float v = 0.f;
    if (q < NL_self) v = nodes[n * NL_self + q];
    else if (q < NL_self + NL_agg) v = agg_nodes[n * NL_agg + (q - NL_self)];
    else if (q < NL_self + NL_agg + AL) v = agg_arcs[n * AL + (q - NL_self - NL_agg)];
    else if (q == NL_self + NL_agg + AL) v = row_scale ? row_scale[n] : 1.f;
    cst[idx] = v;
}

// X0 = pad(x0); go[0] = any_n( ||x0_n - 1|| > thr * ||1|| ) && max_iter > 0   (GNN.py:266, :202-220)
// one thread per (node, 4 columns): coalesced reads of x0, 128-bit stores; the DP/4 lanes of a node are adjacent
static __global__ void init_state_kernel(const float* __restrict__ x0, long long N, int D, int DP, float thr, int max_iter,
                                  float* __restrict__ X0, int* __restrict__ go0) {
    const int LPN = DP >> 2;                       // power of two <= 32
    const long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = item / LPN;
    const int j0 = 4 * (int)(item % LPN);
    float d2 = 0.f;
    if (n < N) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            v[c] = j0 + c < D ? __ldg(x0 + n * D + j0 + c) : 0.f;
            if (j0 + c < D) d2 += (v[c] - 1.f) * (v[c] - 1.f);
        }
        st4(X0 + n * DP + j0, make_float4(v[0], v[1], v[2], v[3]));
    }
    for (int off = 1; off < LPN; off <<= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, off);
    const bool moving = n < N && sqrtf(d2) > thr * sqrtf((float)D);
    const int block_moving = __syncthreads_or(moving ? 1 : 0);
    if (max_iter > 0 && block_moving && threadIdx.x == 0 && *reinterpret_cast<volatile int*>(go0) == 0) atomicOr(go0, 1);
}

// mean / biased variance of h over the nodes from the per-CTA partial sums (fp64), the normalisation
// y = a*h + c, and the Keras moving-average update (once per iteration)
static __global__ void bn_stats_kernel(const int* __restrict__ go_cur, const double* __restrict__ partial, int nblocks, int DP, int D,
                                long long N, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float* __restrict__ moving_mean, float* __restrict__ moving_var, float eps, float momentum,
                                float* __restrict__ stats /* [4][DP]: mean, var, a, c */) {
    if (*reinterpret_cast<const volatile int*>(go_cur) == 0) return;
    // one CTA of 64 threads per column: strided partial sums, then a fixed shuffle / shared-memory tree (deterministic)
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
    float mean = 0.f, var = 0.f, a = 0.f, c = 0.f;
    if (j < D) {
        const double m = s1 / (double)N;
        double v = s2 / (double)N - m * m;
        if (v < 0.) v = 0.;
        mean = (float)m;
        var = (float)v;
        a = gamma[j] * rsqrtf(var + eps);
        c = beta[j] - mean * a;
        moving_mean[j] = moving_mean[j] * momentum + mean * (1.f - momentum);
        moving_var[j] = moving_var[j] * momentum + var * (1.f - momentum);
    }
    stats[j] = mean; stats[DP + j] = var; stats[2 * DP + j] = a; stats[3 * DP + j] = c;
}

// x_{t+1} = a*h + c, convergence test against x_t (training-mode BatchNormalization only).
// BN_APPLY_U 16-byte pieces per thread, all loads first: twice the bytes in flight per thread of the one-piece version.
constexpr int BN_APPLY_U = 2;
template <int DP>
static __global__ void bn_apply_kernel(const int* __restrict__ go_cur, int* __restrict__ go_next, int* __restrict__ k_ptr, int t,
                                const float* __restrict__ h, const float* __restrict__ x_old, const float* __restrict__ stats,
                                long long N, float thr, float* __restrict__ x_new) {
    if (*reinterpret_cast<const volatile int*>(go_cur) == 0) return;
    constexpr int LPN = DP / 4;
    const long long base = blockIdx.x * (long long)(blockDim.x * BN_APPLY_U) + threadIdx.x;
    const int lig = (int)(base % LPN);       // blockDim.x is a multiple of LPN: the same columns for every piece of this thread
    const float4 a = ldg4(stats + 2 * DP + 4 * lig), c = ldg4(stats + 3 * DP + 4 * lig);
    float4 hv[BN_APPLY_U], xo[BN_APPLY_U];
    long long node[BN_APPLY_U];
#pragma unroll
    for (int u = 0; u < BN_APPLY_U; ++u) {
        node[u] = (base + (long long)u * blockDim.x) / LPN;
        hv[u] = xo[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (node[u] < N) {
            hv[u] = ldg4(h + node[u] * DP + 4 * lig);
            xo[u] = ldg4(x_old + node[u] * DP + 4 * lig);
        }
    }
    bool moving = false;
#pragma unroll
    for (int u = 0; u < BN_APPLY_U; ++u) {
        const bool valid = node[u] < N;
        const float4 xn = make_float4(fmaf(a.x, hv[u].x, c.x), fmaf(a.y, hv[u].y, c.y), fmaf(a.z, hv[u].z, c.z), fmaf(a.w, hv[u].w, c.w));
        if (valid) st4(x_new + node[u] * DP + 4 * lig, xn);
        const float dx = xn.x - xo[u].x, dy = xn.y - xo[u].y, dz = xn.z - xo[u].z, dw = xn.w - xo[u].w;
        float d2 = valid ? dx * dx + dy * dy + dz * dz + dw * dw : 0.f;
        float o2 = valid ? xo[u].x * xo[u].x + xo[u].y * xo[u].y + xo[u].z * xo[u].z + xo[u].w * xo[u].w : 0.f;
#pragma unroll
        for (int off = LPN / 2; off > 0; off >>= 1) {
            d2 += __shfl_xor_sync(0xffffffffu, d2, off);
            o2 += __shfl_xor_sync(0xffffffffu, o2, off);
        }
        moving |= valid && (sqrtf(d2) > thr * sqrtf(o2));
    }
    // one atomic per CTA at most, and none once the flag is up (every CTA hitting one address serialises in L2)
    const int block_moving = __syncthreads_or(moving ? 1 : 0);
    if (go_next && block_moving && threadIdx.x == 0 && *reinterpret_cast<volatile int*>(go_next) == 0) atomicOr(go_next, 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) *k_ptr = t + 1;
}

// x_out[N, D] = iterate number k (un-padded); k_out = (float) k
constexpr int FINALIZE_U = 4;
static __global__ void finalize_kernel(const int* __restrict__ k_ptr, const float* __restrict__ base, long long slab_floats,
                                int ring /* 0: slab index = k, else k % ring */, long long N, int D, int DP,
                                float* __restrict__ x_out, float* __restrict__ k_out) {
    const int k = *k_ptr;
    const float* src = base + (size_t)(ring ? (k % ring) : k) * slab_floats;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx == 0 && k_out) *k_out = (float)k;
    if (D == DP) {                                  // no padding: 128-bit copies, FINALIZE_U pieces per thread (all loads first)
        const long long pieces = N * D / 4;
        const long long first = blockIdx.x * (long long)(blockDim.x * FINALIZE_U) + threadIdx.x;
        float4 v[FINALIZE_U];
#pragma unroll
        for (int u = 0; u < FINALIZE_U; ++u) {
            const long long q = first + (long long)u * blockDim.x;
            if (q < pieces) v[u] = ldg4(src + 4 * q);
        }
#pragma unroll
        for (int u = 0; u < FINALIZE_U; ++u) {
            const long long q = first + (long long)u * blockDim.x;
            if (q < pieces) st4(x_out + 4 * q, v[u]);
        }
        return;
    }
    if (idx >= N * D) return;
    const long long n = idx / D;
    const int j = (int)(idx % D);
    x_out[idx] = src[n * DP + j];
}

}  // namespace gnn
