// state_loop.cu -- host side of the state-convergence loop: workspace carving, kernel selection, and the
// enqueue of the whole while_loop (forward) / BPTT sweep (backward) without any host synchronisation.
// C ABI: gnn_state_loop_workspace_bytes / gnn_state_loop_forward / gnn_state_loop_backward.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "state_kernels.h"

namespace gnn {
namespace {

// optional event bracketing of the iteration launches (gnn_profile_iterations)
struct IterProfile {
    bool enabled = false;
    cudaEvent_t begin = nullptr, end = nullptr;
    int launches = 0;
    bool pending = false;
};
static IterProfile g_profile;
static char g_last_kernel[96] = "";
static char g_last_bwd_kernel[96] = "";

struct DeviceInfo {
    int sms = 0, smem_optin = 0;
};

int device_info(DeviceInfo* out) {
    static std::mutex mu;
    static std::map<int, DeviceInfo> cache;
    int dev = 0;
    GNN_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it == cache.end()) {
        DeviceInfo d;
        GNN_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        GNN_CUDA(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        it = cache.emplace(dev, d).first;
    }
    *out = it->second;
    return GNN_OK;
}

// resident CTAs per SM of a kernel at a dynamic shared-memory size (cached) + opt-in attribute
int kernel_occupancy(const void* fn, int threads, size_t smem, int* ctas_per_sm) {
    static std::mutex mu;
    static std::map<std::pair<const void*, size_t>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(fn, smem);
    auto it = cache.find(key);
    if (it == cache.end()) {
        GNN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 0;
        GNN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem));
        if (n < 1) GNN_FAIL(GNN_ERR_UNSUPPORTED, "kernel does not fit on an SM with %zu bytes of shared memory", smem);
        it = cache.emplace(key, n).first;
    } else {
        // the attribute is per function: make sure the largest size seen so far stays configured
        GNN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    *ctas_per_sm = it->second;
    return GNN_OK;
}

// ---- tile shapes --------------------------------------------------------------------------------------
struct TileShape {
    int tn, nt;
};

// big graphs: 128-node tiles; small graphs: 32-node tiles (one warp per CTA) so that more SMs get work
TileShape pick_tile(long long N, int sms) {
    const char* env = getenv("GNN_B200_TILE");
    if (env) {
        int v = atoi(env);
        if (v == 128) return {128, 128};
        if (v == 32) return {32, 32};
    }
    if ((N + 127) / 128 >= (long long)sms) return {128, 128};
    return {32, 32};
}

// ---- workspace -----------------------------------------------------------------------------------------
struct Workspace {
    int* ctl;          // go[0..max_iter], k, stopped, done[0..max_iter-1]
    float* wpack;
    float* cst;
    float* stats;      // [max_iter][4][DP] BatchNormalization batch statistics per iteration
    double* bn_partial;
    float* X;          // iterates: (max_iter + 1) slabs when saving, else 2 (ping-pong)
    float* AGG;        // aggregated states per iteration (saved for backward)
    float* H;          // pre-BatchNormalization outputs per iteration (training-mode BN)
    float* G;          // backward: gradient wrt the current iterate
    float* GS;         // backward: direct (own-state) part
    float* GA;         // backward: part that flows through the aggregation
    float* GH;         // backward: gradient wrt the pre-BN output (training-mode BN)
    float* gcst;       // backward: gradient wrt the per-node constant row
    float* gpartial;   // backward: per-CTA partial parameter gradients
    double* bn_bwd_partial;
    float* bn_bwd_sums;
    double* bn_dgdb;   // [2][DP] dgamma | dbeta accumulated over the iterations
    size_t slab;       // floats per [N, DP] slab
    int x_slabs;
    int max_ctas;
    size_t total;
};

int carve(const gnn_graph* g, const gnn_loop_args* a, const NetLayout& lay, void* base, Workspace* w) {
    DeviceInfo di;
    GNN_TRY(device_info(&di));
    const size_t N = (size_t)g->n_nodes;
    const size_t NG = a->n_global > 0 ? (size_t)a->n_global : N;   // rows of the (replicated) state buffers
    const bool bn_train = a->training && lay.has_bn;
    const bool save = a->save_for_backward != 0;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 4, 256); return o; };
    char* b = (char*)base;
    memset(w, 0, sizeof(*w));
    w->slab = NG * (size_t)lay.DP;
    w->max_ctas = di.sms * 32;
    w->x_slabs = save ? a->max_iter + 1 : 2;
    size_t o_ctl = take((size_t)(2 * a->max_iter + 3) * sizeof(int));
    size_t o_wp = take((size_t)lay.total_floats * 4);
    size_t o_cst = take(N * (size_t)lay.CP * 4);
    size_t o_stats = take(bn_train ? (size_t)a->max_iter * 4 * lay.DP * 4 : 0);
    size_t o_part = take(bn_train ? (size_t)w->max_ctas * 2 * lay.DP * 8 : 0);
    size_t o_X = take((size_t)w->x_slabs * w->slab * 4);
    size_t o_AGG = take(save ? (size_t)a->max_iter * w->slab * 4 : 0);
    size_t o_H = take(bn_train ? (size_t)(save ? a->max_iter : 1) * w->slab * 4 : 0);
    size_t o_G = take(save ? w->slab * 4 : 0);
    size_t o_GS = take(save ? w->slab * 4 : 0);
    size_t o_GA = take(save ? w->slab * 4 : 0);
    size_t o_GH = take(save && bn_train ? w->slab * 4 : 0);
    size_t o_gcst = take(save ? N * (size_t)lay.CP * 4 : 0);
    size_t o_gp = take(save ? (size_t)w->max_ctas * bwd_param_floats(lay) * 4 : 0);
    size_t o_bp = take(save && bn_train ? (size_t)w->max_ctas * 2 * lay.DP * 8 : 0);
    size_t o_bs = take(save && bn_train ? (size_t)2 * lay.DP * 4 : 0);
    size_t o_dg = take(save && bn_train ? (size_t)2 * lay.DP * 8 : 0);
    w->total = off;
    if (b) {
        w->ctl = (int*)(b + o_ctl); w->wpack = (float*)(b + o_wp); w->cst = (float*)(b + o_cst);
        w->stats = (float*)(b + o_stats); w->bn_partial = (double*)(b + o_part); w->X = (float*)(b + o_X);
        w->AGG = (float*)(b + o_AGG); w->H = (float*)(b + o_H); w->G = (float*)(b + o_G); w->GS = (float*)(b + o_GS);
        w->GA = (float*)(b + o_GA); w->GH = (float*)(b + o_GH); w->gcst = (float*)(b + o_gcst);
        w->gpartial = (float*)(b + o_gp); w->bn_bwd_partial = (double*)(b + o_bp); w->bn_bwd_sums = (float*)(b + o_bs);
        w->bn_dgdb = (double*)(b + o_dg);
    }
    return GNN_OK;
}

int check_args(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a) {
    if (!g || !net || !a) GNN_FAIL(GNN_ERR_INVALID, "NULL argument");
    if (g->n_nodes < 0 || g->n_arcs < 0) GNN_FAIL(GNN_ERR_INVALID, "negative graph size");
    if (a->max_iter < 0) GNN_FAIL(GNN_ERR_INVALID, "max_iter < 0");
    if (g->n_nodes > 0 && (!g->rowptr || (!g->col && g->n_arcs > 0))) GNN_FAIL(GNN_ERR_INVALID, "graph CSR missing");
    if (!g->val && !g->row_scale && g->n_arcs > 0) GNN_FAIL(GNN_ERR_INVALID, "graph needs val or row_scale");
    if (a->NL_self < 0 || a->NL_agg < 0 || a->AL < 0) GNN_FAIL(GNN_ERR_INVALID, "negative label width");
    if (a->n_global > 0) {
        if (a->row_offset < 0 || a->row_offset + g->n_nodes > a->n_global) GNN_FAIL(GNN_ERR_INVALID, "partition rows outside the graph");
        if (a->training || a->save_for_backward) GNN_FAIL(GNN_ERR_UNSUPPORTED, "partitioned calls are forward-only");
        if (a->n_peers < 0 || a->n_peers > GNN_MAX_PEERS) GNN_FAIL(GNN_ERR_INVALID, "n_peers outside 0..%d", GNN_MAX_PEERS);
        if (a->n_peers > 1 && (a->rank < 0 || a->rank >= a->n_peers)) GNN_FAIL(GNN_ERR_INVALID, "rank outside 0..n_peers-1");
        if (a->sig_local && a->n_peers > 1)
            for (int r = 0; r < a->n_peers; ++r)
                if (r != a->rank && !a->sig_peer[r]) GNN_FAIL(GNN_ERR_INVALID, "sig_peer[%d] missing", r);
    } else if (a->row_offset != 0 || a->exchange || a->n_peers > 1) {
        GNN_FAIL(GNN_ERR_INVALID, "row_offset / exchange / peers need n_global");
    }
    return GNN_OK;
}

struct Plan {
    bool ws;          // warp-specialised pipelined kernel (state_fwd_ws.cuh / state_fwd_tc.cuh)
    bool tc;          // ... its tcgen05 / tensor-memory variant
    NetLayout lay;
    TileShape ts;
    IterKernel kernel;
    size_t smem;
    int scol_cap;
    int ring_slots, slot_rows;
    int grid;
    bool has_val;
};

// the warp-specialised kernels cover: one Dense layer, padded state width 16..32, no active dropout, enough tiles
bool ws_applicable(const gnn_graph* g, const gnn_loop_args* a, const NetLayout& lay, int sms) {
    const char* env = getenv("GNN_B200_KERNEL");
    if (env && !strcmp(env, "sym")) return false;
    if (lay.L != 1 || lay.DP < 16 || lay.DP > 32 || lay.CP > 16) return false;   // (constant row: at most two k-steps)
    // the loader warp stages row pointers / arc sources / scales with bulk copies: 16-byte aligned arrays
    auto misaligned = [](const void* ptr) { return ptr && ((uintptr_t)ptr & 15) != 0; };
    if (misaligned(g->rowptr) || misaligned(g->col) || misaligned(g->val) || misaligned(g->row_scale)) return false;
    if (a->training)
        for (int i = 0; i <= lay.L; ++i) if (lay.drop[i] > 0.f) return false;
    if (env && (!strcmp(env, "ws") || !strcmp(env, "tc"))) return true;
    return (g->n_nodes + WS_TN - 1) / WS_TN >= 2LL * sms;
}

int make_plan(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, Plan* plan) {
    DeviceInfo di;
    GNN_TRY(device_info(&di));
    plan->ts = pick_tile(g->n_nodes, di.sms);
    GNN_TRY(make_layout(net, 1, a->D, a->NL_self, a->NL_agg, a->AL, plan->ts.tn, plan->ts.nt, &plan->lay));
    const NetLayout& lay = plan->lay;
    plan->has_val = g->val != nullptr && g->row_scale == nullptr;
    const KernelSet* ks = kernel_set(lay.DP);
    if (!ks) GNN_FAIL(GNN_ERR_UNSUPPORTED, "no kernel for padded state width %d", lay.DP);
    plan->ws = plan->tc = false;
    plan->ring_slots = plan->slot_rows = 0;

    if (ws_applicable(g, a, lay, di.sms) && ks->iter_ws[plan->has_val ? 1 : 0][0]) {
        // Two pipelines cover these graphs.  Measured on the same box with both shared-memory layouts aligned (C4, DP = 32): the
        // mma.sync pipeline 0.222 / 0.196 ms per iteration (uniform / local sources), the tcgen05 pipeline 0.237 / 0.219 ms; equal on
        // the C5 batches (DP = 16).  Node-range partition over 2 GPUs: 0.186 / 0.134 ms against 0.260 / 0.128 ms.  So the mma.sync
        // pipeline is the planner's choice and the tcgen05 pipeline runs on request (GNN_B200_KERNEL=tc).
        const char* env = getenv("GNN_B200_KERNEL");
        const bool tc = !plan->has_val && ks->iter_tc[0] && env && !strcmp(env, "tc");
        const bool bn_tr = a->training && lay.has_bn;
        // arc-index capacity per tile: 1.5x the average tile; the landing ring takes all the shared memory that is left
        // (at least 4 average sub-tiles, at most 4096 rows)
        const long long avg = g->n_nodes > 0 ? (g->n_arcs * WS_TN) / g->n_nodes : 0;
        long long capc_want = avg * 3 / 2;
        if (g->max_block16_arcs > 0) capc_want = std::min<long long>(capc_want, (long long)WS_NSUB * g->max_block16_arcs);   // no tile has more
        const int capc = (int)std::min<long long>(4096, std::max<long long>(128, (capc_want + 15) / 16 * 16));
        // static shared memory: ws = mbarriers + BN accumulators (4.6 KB), tc = mbarriers only
        const size_t budget = (size_t)di.smem_optin - (tc ? 1024 : 7168);
        const size_t fixed = tc ? tc_smem_bytes(lay, 0, capc, bn_tr) : ws_smem_bytes(lay, 0, capc, plan->has_val);
        const int ring = fixed < budget ? (int)std::min<size_t>(4096, (budget - fixed) / ((size_t)lay.DP * 4)) : 0;
        // slots of the ring: one sub-tile (16 nodes) each.  When the caller knows the densest block of 16 rows the slot
        // holds exactly that (bounded by 1.5x the average: hubs read their excess arcs directly); otherwise 1/8 above the
        // average sub-tile.  As many slots as fit, at most WS_SLOTS.
        const long long avg_sub = std::max<long long>(16, avg / WS_NSUB);
        long long want = g->max_block16_arcs > 0 ? std::min<long long>(g->max_block16_arcs, std::max<long long>(32, avg_sub * 3 / 2))
                                                 : avg_sub * 9 / 8;
        const int slot_rows = (int)((want + 3) / 4 * 4);
        // 8 or 16 slots (a slot then always serves the same issue warp); when not even 8 average sub-tiles fit, the slots shrink
        // and the compute warps read the excess arcs directly
        int slot_rows_fit = slot_rows;
        int slots = ring / slot_rows >= 16 ? 16 : 8;
        if (ring / slot_rows < 8) slot_rows_fit = (ring / 8) & ~3;
        if (slot_rows_fit < 16 || slot_rows_fit * 2 < avg_sub) slots = 0;
        if (slots >= 8) {
            plan->ws = true;
            plan->tc = tc;
            {
                const int peers = a->n_global > 0 && a->n_peers > 1 ? 1 : 0;
                plan->kernel = tc ? ks->iter_tc[peers] : ks->iter_ws[plan->has_val ? 1 : 0][peers];
            }
            plan->ts = TileShape{WS_TN, tc ? TC_THREADS : WS_THREADS};
            plan->scol_cap = capc;
            plan->ring_slots = slots;
            plan->slot_rows = slot_rows_fit;
            plan->smem = tc ? tc_smem_bytes(lay, slots * slot_rows_fit, capc, bn_tr) : ws_smem_bytes(lay, slots * slot_rows_fit, capc, plan->has_val);
            int occ = 0;
            GNN_TRY(kernel_occupancy((const void*)plan->kernel, plan->ts.nt, plan->smem, &occ));
            const long long ntiles = (g->n_nodes + WS_TN - 1) / WS_TN;
            plan->grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)occ * di.sms));
            return GNN_OK;
        }
    }

    plan->kernel = ks->iter[plan->ts.tn == 128 ? 0 : 1][plan->has_val ? 1 : 0];
    if (!plan->kernel) GNN_FAIL(GNN_ERR_UNSUPPORTED, "no kernel for padded state width %d", lay.DP);
    const int TN = plan->ts.tn;
    const size_t fixed = ((size_t)lay.fwd_floats + (size_t)TN * lay.SA + (size_t)TN * lay.SB + ((TN + 1 + 3) & ~3)) * 4;
    // arc indices of a tile are staged in shared memory when they fit. Two candidate capacities (1.6x and 1.2x the
    // average tile): the smaller one wins when it lets one more CTA live on an SM
    const long long avg = g->n_nodes > 0 ? (g->n_arcs * TN) / g->n_nodes : 0;
    const size_t per_arc = plan->has_val ? 8 : 4;
    if (fixed > (size_t)di.smem_optin)
        GNN_FAIL(GNN_ERR_UNSUPPORTED, "net_state too large for shared memory: %zu bytes needed, %d available", fixed, di.smem_optin);
    int best_cap = 0, best_occ = 0;
    const long long wanted[2] = {(avg * 8 / 5 + 255) / 256 * 256, (avg * 6 / 5 + 127) / 128 * 128};
    for (int c = 0; c < 2; ++c) {
        int cap = (int)std::min<long long>(8192, std::max<long long>(256, wanted[c]));
        while (cap > 0 && fixed + cap * per_arc > (size_t)di.smem_optin) cap /= 2;
        int occ = 0;
        GNN_TRY(kernel_occupancy((const void*)plan->kernel, plan->ts.nt, fixed + cap * per_arc, &occ));
        if (occ > best_occ) { best_occ = occ; best_cap = cap; }
    }
    plan->scol_cap = best_cap;
    plan->smem = fixed + best_cap * per_arc;
    int occ = 0;
    GNN_TRY(kernel_occupancy((const void*)plan->kernel, plan->ts.nt, plan->smem, &occ));
    long long ntiles = (g->n_nodes + TN - 1) / TN;
    plan->grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)occ * di.sms));
    if (plan->grid > di.sms * 32) plan->grid = di.sms * 32;
    return GNN_OK;
}

}  // namespace
}  // namespace gnn

using namespace gnn;

extern "C" int gnn_state_loop_workspace_bytes(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, size_t* bytes) {
    GNN_TRY(check_args(g, net, a));
    if (!bytes) GNN_FAIL(GNN_ERR_INVALID, "NULL bytes");
    NetLayout lay;
    GNN_TRY(make_layout(net, 1, a->D, a->NL_self, a->NL_agg, a->AL, 128, 128, &lay));
    Workspace w;
    GNN_TRY(carve(g, a, lay, nullptr, &w));
    *bytes = w.total;
    return GNN_OK;
}

extern "C" int gnn_state_loop_forward(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNN_TRY(check_args(g, net, a));
    if (!a->x0 || !a->x_out) GNN_FAIL(GNN_ERR_INVALID, "x0 / x_out missing");
    if (net->has_bn && (!net->bn_gamma || !net->bn_beta || !net->bn_moving_mean || !net->bn_moving_var))
        GNN_FAIL(GNN_ERR_INVALID, "BatchNormalization parameters missing");
    Plan plan;
    GNN_TRY(make_plan(g, net, a, &plan));
    const NetLayout& lay = plan.lay;
    Workspace w;
    GNN_TRY(carve(g, a, lay, workspace, &w));
    if (!workspace || workspace_bytes < w.total) GNN_FAIL(GNN_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.total);
    const long long N = g->n_nodes;
    const long long NGLOB = a->n_global > 0 ? a->n_global : N;
    const bool bn_train = a->training && lay.has_bn;
    const bool save = a->save_for_backward != 0;
    int* go = w.ctl;
    int* kptr = w.ctl + a->max_iter + 1;

    GNN_CUDA(cudaMemsetAsync(w.ctl, 0, (size_t)(2 * a->max_iter + 3) * sizeof(int), stream));
    {
        PackParams pp;
        pp.net = *net; pp.lay = lay; pp.wpack = w.wpack; pp.state_loop = 1;
        pp.bn_inference = lay.has_bn && !a->training;
        pack_net_kernel<<<(lay.total_floats + 255) / 256, 256, 0, stream>>>(pp);
        GNN_LAUNCH_CHECK();
    }
    if (NGLOB == 0) {
        GNN_CUDA(cudaMemsetAsync(a->k_out, 0, sizeof(float), stream));
        return GNN_OK;
    }
    if (N > 0) {
        pack_cst_kernel<<<(unsigned)ceil_div(N * lay.CP, 256), 256, 0, stream>>>(a->nodes, a->agg_nodes, a->agg_arcs,
                                                                                 plan.has_val ? nullptr : g->row_scale, N, lay.NL_self,
                                                                                 lay.NL_agg, lay.AL, lay.CP, w.cst);
        GNN_LAUNCH_CHECK();
    }
    // the initial state is replicated: every rank pads all rows and evaluates the first condition on all of them
    init_state_kernel<<<(unsigned)ceil_div(NGLOB * (lay.DP / 4), 256), 256, 0, stream>>>(a->x0, NGLOB, lay.D, lay.DP, a->threshold, a->max_iter, w.X, go);
    GNN_LAUNCH_CHECK();

    IterParams p;
    memset(&p, 0, sizeof(p));
    p.rowptr = g->rowptr; p.col = g->col; p.val = plan.has_val ? g->val : nullptr; p.row_scale = plan.has_val ? nullptr : g->row_scale;
    p.N = N; p.E = g->n_arcs; p.row_offset = a->row_offset;
    p.cst = w.cst; p.wpack = w.wpack; p.k_ptr = kptr; p.thr = a->threshold; p.bn_partial = w.bn_partial;
    p.bn_train = bn_train; p.seed = a->seed; p.seed_dev = a->seed_dev; p.training = a->training; p.scol_cap = plan.scol_cap; p.ring_slots = plan.ring_slots; p.slot_rows = plan.slot_rows;
    { const char* dbg = getenv("GNN_B200_WS_DEBUG"); p.ws_debug = dbg ? atoi(dbg) : 0; } p.net = lay;
    p.n_peers = a->n_global > 0 ? a->n_peers : 0; p.rank = a->rank; p.peer_mask = a->peer_mask;
    if (p.n_peers > 1 && a->sig_local) {
        p.sig_local = a->sig_local; p.sig_epoch = a->sig_epoch; p.sig_iters = a->max_iter + 1; p.stopped = w.ctl + a->max_iter + 2;
        for (int r = 0; r < p.n_peers; ++r) p.sig_peer[r] = a->sig_peer[r];
    }
    BnApplyKernel bn_apply = kernel_set(lay.DP)->bn_apply;

    if (plan.tc) snprintf(g_last_kernel, sizeof(g_last_kernel), "state_iter_tc_kernel<%d>", lay.DP);
    else snprintf(g_last_kernel, sizeof(g_last_kernel), plan.ws ? "state_iter_ws_kernel<%d,%s>" : "state_iter_kernel<%d,%s,%d,%d>", lay.DP,
                  plan.has_val ? "true" : "false", plan.ts.tn, plan.ts.nt);
    if (g_profile.enabled) {
        GNN_CUDA(cudaEventRecord(g_profile.begin, stream));
        g_profile.launches = 0;
    }
    for (int t = 0; t < a->max_iter; ++t) {
        const float* x_in = w.X + (size_t)(save ? t : (t & 1)) * w.slab;
        float* x_next = w.X + (size_t)(save ? t + 1 : ((t + 1) & 1)) * w.slab;
        p.x_in = x_in;
        p.x_out = bn_train ? w.H + (size_t)(save ? t : 0) * w.slab : x_next;
        p.agg_save = save ? w.AGG + (size_t)t * w.slab : nullptr;
        p.go_cur = go + t;
        p.go_next = (t + 1 < a->max_iter) ? go + t + 1 : nullptr;
        p.t = t;
        p.done_ctr = w.ctl + a->max_iter + 3 + t;
        for (int r = 0; r < p.n_peers; ++r) p.peer_out[r] = a->peer_state[r] ? a->peer_state[r] + (size_t)((t + 1) & 1) * w.slab : nullptr;
        if (N > 0) {
            void* args[] = {(void*)&p};
            GNN_CUDA(cudaLaunchKernel((const void*)plan.kernel, dim3(plan.grid), dim3(plan.ts.nt), args, plan.smem, stream));
            GNN_LAUNCH_CHECK();
            if (g_profile.enabled) ++g_profile.launches;
        }
        if (a->exchange)
            a->exchange(a->exchange_user, t, (int64_t)((char*)x_next - (char*)workspace),
                        p.go_next ? (int64_t)((char*)p.go_next - (char*)workspace) : (int64_t)-1);
        if (bn_train) {
            float* stats = w.stats + (size_t)t * 4 * lay.DP;
            bn_stats_kernel<<<lay.DP, 64, 0, stream>>>(go + t, w.bn_partial, plan.grid, lay.DP, lay.D, N, net->bn_gamma,
                                                                  net->bn_beta, net->bn_moving_mean, net->bn_moving_var, net->bn_eps,
                                                                  net->bn_momentum, stats);
            GNN_LAUNCH_CHECK();
            const long long items = N * (lay.DP / 4);
            bn_apply<<<(unsigned)ceil_div(items, 256 * BN_APPLY_U), 256, 0, stream>>>(go + t, p.go_next, kptr, t, p.x_out, x_in, stats, N, a->threshold, x_next);
            GNN_LAUNCH_CHECK();
        }
    }
    if (g_profile.enabled) {
        GNN_CUDA(cudaEventRecord(g_profile.end, stream));
        g_profile.pending = true;
    }
    // (no padding: 16-byte pieces, FINALIZE_U per thread; otherwise one element per thread)
    const long long fin_items = lay.D == lay.DP ? ceil_div(NGLOB * lay.D / 4, (long long)FINALIZE_U) : NGLOB * lay.D;
    finalize_kernel<<<(unsigned)ceil_div(std::max<long long>(fin_items, 1), 256), 256, 0, stream>>>(
        kptr, w.X, (long long)w.slab, save ? 0 : 2, NGLOB, lay.D, lay.DP, a->x_out, a->k_out);
    GNN_LAUNCH_CHECK();
    return GNN_OK;
}

extern "C" int gnn_state_loop_layout(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, size_t* state_offset, size_t* state_bytes) {
    GNN_TRY(check_args(g, net, a));
    NetLayout lay;
    GNN_TRY(make_layout(net, 1, a->D, a->NL_self, a->NL_agg, a->AL, 128, 128, &lay));
    Workspace w;
    char probe[1];   // carve only computes offsets relative to the base
    GNN_TRY(carve(g, a, lay, probe, &w));
    if (state_offset) *state_offset = (size_t)((char*)w.X - probe);
    if (state_bytes) *state_bytes = w.slab * 4;
    return GNN_OK;
}

extern "C" const char* gnn_last_forward_kernel(void) { return g_last_kernel; }
extern "C" const char* gnn_last_backward_kernel(void) { return g_last_bwd_kernel; }

extern "C" int gnn_profile_iterations(int32_t enable) {
    if (enable && !g_profile.begin) {
        GNN_CUDA(cudaEventCreate(&g_profile.begin));
        GNN_CUDA(cudaEventCreate(&g_profile.end));
    }
    g_profile.enabled = enable != 0;
    g_profile.pending = false;
    return GNN_OK;
}

extern "C" int gnn_profile_last_iterations(float* elapsed_ms, int32_t* launches) {
    if (!g_profile.pending) GNN_FAIL(GNN_ERR_INVALID, "no profiled forward call is pending");
    GNN_CUDA(cudaEventSynchronize(g_profile.end));
    float ms = 0.f;
    GNN_CUDA(cudaEventElapsedTime(&ms, g_profile.begin, g_profile.end));
    if (elapsed_ms) *elapsed_ms = ms;
    if (launches) *launches = g_profile.launches;
    g_profile.pending = false;
    return GNN_OK;
}

#include "state_bwd_host.inl"
