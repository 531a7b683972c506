// state_bwd_host.inl -- host side of gnn_state_loop_backward (included at the end of state_loop.cu)

extern "C" int gnn_state_loop_backward(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, const float* g_x,
                                       gnn_mlp_grad* grad, float* g_x0, float* g_nodes, float* g_agg_nodes, float* g_agg_arcs,
                                       void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNN_TRY(check_args(g, net, a));
    if (!a->save_for_backward) GNN_FAIL(GNN_ERR_INVALID, "backward needs a forward call with save_for_backward");
    if (!g_x || !grad) GNN_FAIL(GNN_ERR_INVALID, "g_x / grad missing");
    if (g->n_arcs > 0 && (!g->rowptr_T || !g->col_T)) GNN_FAIL(GNN_ERR_INVALID, "transposed CSR missing");
    DeviceInfo di;
    GNN_TRY(device_info(&di));
    const long long N = g->n_nodes;
    // the backward node kernel keeps more tiles in shared memory: 64-node tiles unless the graph is tiny
    TileShape ts = ((N + 63) / 64 >= (long long)di.sms) ? TileShape{64, 128} : TileShape{32, 32};
    NetLayout lay;
    GNN_TRY(make_layout(net, 1, a->D, a->NL_self, a->NL_agg, a->AL, 128, 128, &lay));
    Workspace w;
    GNN_TRY(carve(g, a, lay, workspace, &w));
    if (!workspace || workspace_bytes < w.total) GNN_FAIL(GNN_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, w.total);
    const bool bn_train = a->training && lay.has_bn;
    const bool has_val = g->val != nullptr && g->row_scale == nullptr;
    if (has_val && g->n_arcs > 0 && !g->val_T) GNN_FAIL(GNN_ERR_INVALID, "val_T missing");
    const bool want_cst = g_nodes || g_agg_nodes || g_agg_arcs;
    int* kptr = w.ctl + a->max_iter + 1;
    const size_t pfloats = bwd_param_floats(lay);

    if (N == 0) {
        ReduceParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.grad = *grad; rp.lay = lay; rp.gpartial = w.gpartial; rp.nblocks = 0; rp.state_loop = 1;
        reduce_params_kernel<<<(lay.fwd_floats + 255) / 256, 256, 0, stream>>>(rp);
        GNN_LAUNCH_CHECK();
        return GNN_OK;
    }

    // The packed net in the workspace (weights, transposed copies, final affine) is the one the forward call packed: it IS the
    // value saved for backward.  Not re-packed from the live parameter pointers: an in-place update of the parameters between
    // forward and backward (an optimizer step of an interleaved user loop) must not mix new weights with the saved iterates.
    // kernel + shared memory plan of the node kernel
    BwdNodeParams p;
    memset(&p, 0, sizeof(p));
    const bool y_saved = !(lay.has_bn && !a->training);  // inference-BN stores a*y+c: recompute y instead
    const bool need_dB = lay.L >= 2 || !y_saved;
    int maxw = lay.DP;
    for (int l = 0; l < lay.L; ++l) if (lay.out_pad[l] > maxw) maxw = lay.out_pad[l];
    p.SU = odd_quad_stride(lay.KP);
    p.SD = odd_quad_stride(maxw);
    const KernelSet* ks = kernel_set(lay.DP);
    if (!ks) GNN_FAIL(GNN_ERR_UNSUPPORTED, "no kernels for padded state width %d", lay.DP);
    BwdNodeKernel node_kernel = nullptr;
    size_t smem = 0;
    int grid = 1;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int TN = ts.tn;
        // [packed net | CTA gradient accumulators | input tile | delta tile(s) | saved hidden activations]
        size_t fl = (size_t)lay.total_floats + lay.fwd_floats + (size_t)TN * p.SU + (size_t)TN * p.SD * (need_dB ? 2 : 1);
        int act_floats = 0;
        for (int l = 1; l < lay.L; ++l) {
            p.act_stride[l] = odd_quad_stride(lay.in_pad[l]);
            p.act_off[l] = act_floats;
            act_floats += TN * p.act_stride[l];
        }
        smem = (fl + act_floats) * 4;
        node_kernel = ks->bwd_node[ts.tn == 64 ? 0 : 1];
        if (smem <= (size_t)di.smem_optin) break;
        if (ts.tn == 32) GNN_FAIL(GNN_ERR_UNSUPPORTED, "net_state too large for the backward kernel: %zu bytes of shared memory", smem);
        ts = TileShape{32, 32};
    }
    if (!node_kernel) GNN_FAIL(GNN_ERR_UNSUPPORTED, "no backward kernel for padded state width %d", lay.DP);
    // single Dense layer without active dropout on a graph that fills the GPU: the pipelined kernel of state_bwd_l1.cuh
    // (GNN_B200_BWD=phased keeps the phase-structured kernel: comparison runs)
    {
        const char* env = getenv("GNN_B200_BWD");
        if (ks->bwd_node_l1 && bwd_l1_applicable(lay, a->training != 0, y_saved) && N >= 64LL * di.sms && !(env && !strcmp(env, "phased")) &&
            bwd_l1_smem_bytes(lay) <= (size_t)di.smem_optin) {
            node_kernel = ks->bwd_node_l1;
            ts = TileShape{BL_TN, BL_NT};
            smem = bwd_l1_smem_bytes(lay);
            p.SU = bwd_l1_su(lay);
            p.SD = bwd_l1_sd(lay);
        }
    }
    snprintf(g_last_bwd_kernel, sizeof(g_last_bwd_kernel), node_kernel == ks->bwd_node_l1 ? "state_bwd_node_l1_kernel<%d>" : "state_bwd_node_kernel<%d,%d,%d>",
             lay.DP, ts.tn, ts.nt);
    int occ = 0;
    GNN_TRY(kernel_occupancy((const void*)node_kernel, ts.nt, smem, &occ));
    const long long ntiles = (N + ts.tn - 1) / ts.tn;
    grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(ntiles, (long long)occ * di.sms), w.max_ctas));

    GNN_CUDA(cudaMemsetAsync(w.gpartial, 0, (size_t)grid * pfloats * 4, stream));
    if (want_cst) GNN_CUDA(cudaMemsetAsync(w.gcst, 0, (size_t)N * lay.CP * 4, stream));
    if (bn_train) GNN_CUDA(cudaMemsetAsync(w.bn_dgdb, 0, (size_t)2 * lay.DP * 8, stream));
    pad_rows_kernel<<<(unsigned)ceil_div(N * lay.DP, 256), 256, 0, stream>>>(g_x, N, lay.D, lay.DP, w.G);
    GNN_LAUNCH_CHECK();

    p.N = N; p.G = w.G; p.cst = w.cst; p.wpack = w.wpack; p.GS = w.GS; p.GA = w.GA; p.gcst = want_cst ? w.gcst : nullptr;
    p.gpartial = w.gpartial; p.k_ptr = kptr; p.seed = a->seed; p.seed_dev = a->seed_dev; p.training = a->training;
    p.row_scale_mode = has_val ? 0 : 1; p.bn_eps = net->bn_eps; p.net = lay; p.has_dB = need_dB ? 1 : 0;
    ScatterKernel scatter = ks->scatter[has_val ? 1 : 0];
    BnBwdReduceKernel bn_reduce = ks->bn_bwd_reduce;
    // grid-stride reduction: a few CTAs per SM; the finalize kernel walks the partials serially, so keep them few
    const int red_grid = (int)std::min<long long>(2LL * di.sms, std::max<long long>(1, ceil_div(N * (lay.DP / 4), 256)));
    const long long items = N * (lay.DP / 4);

    for (int t = a->max_iter - 1; t >= 0; --t) {
        p.t = t;
        p.x_t = w.X + (size_t)t * w.slab;
        p.agg_t = w.AGG + (size_t)t * w.slab;
        p.y_t = bn_train ? w.H + (size_t)t * w.slab : (y_saved ? w.X + (size_t)(t + 1) * w.slab : nullptr);
        if (bn_train) {
            const float* stats = w.stats + (size_t)t * 4 * lay.DP;
            bn_reduce<<<red_grid, 256, 0, stream>>>(kptr, t, w.G, p.y_t, stats, net->bn_eps, N, w.bn_bwd_partial);
            GNN_LAUNCH_CHECK();
            bn_bwd_finalize_kernel<<<lay.DP, 64, 0, stream>>>(kptr, t, w.bn_bwd_partial, red_grid, lay.DP, lay.D, N,
                                                                         w.bn_bwd_sums, w.bn_dgdb);
            GNN_LAUNCH_CHECK();
            p.bn_stats = stats;
            p.bn_sums = w.bn_bwd_sums;
        }
        void* args[] = {(void*)&p};
        GNN_CUDA(cudaLaunchKernel((const void*)node_kernel, dim3(grid), dim3(ts.nt), args, smem, stream));
        GNN_LAUNCH_CHECK();
        scatter<<<(unsigned)ceil_div(items, 256), 256, 0, stream>>>(kptr, t, g->rowptr_T, g->col_T, has_val ? g->val_T : nullptr, N, w.GS,
                                                                   w.GA, w.G);
        GNN_LAUNCH_CHECK();
    }

    ReduceParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.grad = *grad; rp.lay = lay; rp.gpartial = w.gpartial; rp.nblocks = grid; rp.state_loop = 1;
    rp.dgamma_dbeta = bn_train ? w.bn_dgdb : nullptr;
    reduce_params_kernel<<<(lay.fwd_floats + 255) / 256, 256, 0, stream>>>(rp);
    GNN_LAUNCH_CHECK();
    if (g_x0) {
        unpad_rows_kernel<<<(unsigned)ceil_div(N * lay.D, 256), 256, 0, stream>>>(w.G, N, lay.D, lay.DP, g_x0);
        GNN_LAUNCH_CHECK();
    }
    if (want_cst && lay.C > 0) {
        unpack_gcst_kernel<<<(unsigned)ceil_div(N * lay.C, 256), 256, 0, stream>>>(w.gcst, N, lay.NL_self, lay.NL_agg, lay.AL, lay.CP, g_nodes,
                                                                                 g_agg_nodes, g_agg_arcs);
        GNN_LAUNCH_CHECK();
    }
    return GNN_OK;
}
