// net_layout.cuh -- padded on-chip layout of an MLP (Dense chain + final affine) and of the state-loop tile.
//
// HBM / shared-memory layout decisions:
//   * state rows are padded to DP = next power of two >= max(D, 4) floats so that a row is read with
//     LPN = DP/4 128-bit loads by LPN adjacent lanes;
//   * the loop-invariant part of the net_state input is packed once per Loop into one row per node
//     cst[N, CP] = [nodes(NL_self) | agg_nodes(NL_agg) | agg_arcs(AL) | row_scale | 0-pad], CP multiple of 4;
//   * the first Dense layer is re-indexed to the internal input order [state(DP) | agg_state(DP) | cst(CP)]
//     (rows of padding are zero), every layer's width is padded to a multiple of 4;
//   * packed buffer: for each layer  W[in_pad x out_pad] (row-major), b[out_pad], then Wt[out_pad x in_pad]
//     (transposed copy, backward only), and at the end the final affine a[DP_out], c[DP_out]
//     (BatchNormalization in inference mode folded to y = a*h + c; identity otherwise).
#pragma once
#include "common.cuh"

namespace gnn {

struct NetLayout {
    int L;
    int in_dim[GNN_MAX_LAYERS], out_dim[GNN_MAX_LAYERS];
    int in_pad[GNN_MAX_LAYERS], out_pad[GNN_MAX_LAYERS];
    int act[GNN_MAX_LAYERS];
    int w_off[GNN_MAX_LAYERS], b_off[GNN_MAX_LAYERS], wt_off[GNN_MAX_LAYERS];
    int aff_off;      // a[last_pad], c[last_pad]
    int fwd_floats;   // weights + biases + affine, what the forward kernels stage in shared memory (W, b only + aff)
    int total_floats; // everything incl. transposed copies
    float drop[GNN_MAX_LAYERS + 1];
    int has_bn;
    // state-loop specifics (0 for a plain dense MLP)
    int D, DP, C, CP, KP, NL_self, NL_agg, AL;
    int SA, SB;       // shared-memory row strides (floats), SA/4 and SB/4 odd
    int final_off;    // column of bufA where the last layer writes the new state (DP, or past the inputs)
};

// stride with (s / 4) odd: 8 consecutive rows then start in 8 distinct 16-byte bank groups
static inline int odd_quad_stride(int width) {
    int s = round_up(width, 4);
    if (((s / 4) & 1) == 0) s += 4;
    return s;
}

// state_loop != 0: layer 0 input is [state | agg | cst] in the internal order
// tile_nodes / tile_threads: shape of the forward kernel variant that will consume the layout
static inline int make_layout(const gnn_mlp* net, int state_loop, int D, int NL_self, int NL_agg, int AL, int tile_nodes,
                              int tile_threads, NetLayout* out) {
    NetLayout& l = *out;
    memset(&l, 0, sizeof(l));
    if (net->n_layers < 1 || net->n_layers > GNN_MAX_LAYERS) GNN_FAIL(GNN_ERR_UNSUPPORTED, "MLP needs 1..%d Dense layers, got %d", GNN_MAX_LAYERS, net->n_layers);
    l.L = net->n_layers;
    l.has_bn = net->has_bn;
    for (int i = 0; i <= l.L; ++i) {
        l.drop[i] = net->drop_rate[i];
        if (l.drop[i] < 0.f || l.drop[i] >= 1.f) GNN_FAIL(GNN_ERR_INVALID, "dropout rate %f outside [0,1)", l.drop[i]);
    }
    if (state_loop) {
        if (D < 1) GNN_FAIL(GNN_ERR_INVALID, "state width must be >= 1");
        l.D = D; l.NL_self = NL_self; l.NL_agg = NL_agg; l.AL = AL;
        l.DP = next_pow2(D < 4 ? 4 : D);
        if (l.DP > 128) GNN_FAIL(GNN_ERR_UNSUPPORTED, "state width %d > 128 is not implemented", D);
        l.C = NL_self + NL_agg + AL;
        l.CP = round_up(l.C + 1, 4);
        l.KP = 2 * l.DP + l.CP;
        int expect = 2 * D + l.C;
        if (net->dims[0] != expect) GNN_FAIL(GNN_ERR_INVALID, "net_state input width %d != AL + 2*(NL + D) = %d", net->dims[0], expect);
        if (net->dims[l.L] != D) GNN_FAIL(GNN_ERR_INVALID, "net_state output width %d != state width %d", net->dims[l.L], D);
    }
    int off = 0;
    for (int i = 0; i < l.L; ++i) {
        l.in_dim[i] = net->dims[i];
        l.out_dim[i] = net->dims[i + 1];
        l.act[i] = net->act[i];
        if (l.in_dim[i] < 1 || l.out_dim[i] < 1) GNN_FAIL(GNN_ERR_INVALID, "layer %d has an empty dimension", i);
        if (state_loop && l.act[i] == GNN_ACT_SOFTMAX) GNN_FAIL(GNN_ERR_UNSUPPORTED, "softmax inside net_state is not implemented");
        l.in_pad[i] = (i == 0) ? (state_loop ? l.KP : round_up(l.in_dim[0], 4)) : l.out_pad[i - 1];
        l.out_pad[i] = (state_loop && i == l.L - 1) ? l.DP : round_up(l.out_dim[i], 4);
        l.w_off[i] = off; off += l.in_pad[i] * l.out_pad[i];
        l.b_off[i] = off; off += l.out_pad[i];
    }
    l.aff_off = off; off += 2 * l.out_pad[l.L - 1];
    l.fwd_floats = off;
    for (int i = 0; i < l.L; ++i) { l.wt_off[i] = off; off += l.in_pad[i] * l.out_pad[i]; }
    l.total_floats = off;
    if (state_loop) {
        // bufA: layer-0 input (KP) ; later [state_old(DP) | odd-layer output or final output]
        int wa = l.KP;
        if (2 * l.DP > wa) wa = 2 * l.DP;
        for (int i = 1; i < l.L - 1; i += 2) if (l.DP + l.out_pad[i] > wa) wa = l.DP + l.out_pad[i];
        int wb = 0;
        for (int i = 0; i < l.L - 1; i += 2) if (l.out_pad[i] > wb) wb = l.out_pad[i];
        // last layer output: in place at column DP when one micro-tile (8 nodes x 4 units) per thread covers it,
        // otherwise in a region of its own behind everything else
        l.final_off = l.DP;
        if ((tile_nodes / 8) * (l.DP / 4) > tile_threads) { l.final_off = round_up(wa, 4); wa = l.final_off + l.DP; }
        l.SA = odd_quad_stride(wa);
        l.SB = wb ? odd_quad_stride(wb) : 0;
    }
    return GNN_OK;
}

// Keras column of internal input column r of the state net (-1 for padding / the row-scale slot)
__host__ __device__ __forceinline__ int keras_input_col(int r, int D, int DP, int NL_self, int NL_agg, int AL) {
    if (r < DP) return r < D ? r : -1;
    if (r < 2 * DP) { int q = r - DP; return q < D ? D + NL_self + q : -1; }
    int q = r - 2 * DP;
    if (q < NL_self) return D + q;
    if (q < NL_self + NL_agg) return 2 * D + NL_self + (q - NL_self);
    if (q < NL_self + NL_agg + AL) return 2 * D + NL_self + NL_agg + (q - NL_self - NL_agg);
    return -1;
}

}  // namespace gnn
