// state_inst.cu -- instantiates every templated kernel for ONE padded state width (compile with -DGNN_DP=<4..128>)
#ifndef GNN_DP
#error "compile with -DGNN_DP=<padded state width>"
#endif
#include "state_kernels.h"

namespace gnn {

#define GNN_CAT2(a, b) a##b
#define GNN_CAT(a, b) GNN_CAT2(a, b)

const KernelSet* GNN_CAT(kernel_set_dp, GNN_DP)() {
    static const KernelSet set = {
        {{state_iter_kernel<GNN_DP, false, 128, 128>, state_iter_kernel<GNN_DP, true, 128, 128>},
         {state_iter_kernel<GNN_DP, false, 32, 32>, state_iter_kernel<GNN_DP, true, 32, 32>}},
#if GNN_DP >= 16 && GNN_DP <= 32
        {{state_iter_ws_kernel<GNN_DP, false, false>, state_iter_ws_kernel<GNN_DP, false, true>},
         {state_iter_ws_kernel<GNN_DP, true, false>, state_iter_ws_kernel<GNN_DP, true, true>}},
        {state_iter_tc_kernel<GNN_DP, false>, state_iter_tc_kernel<GNN_DP, true>},
#else
        {{nullptr, nullptr}, {nullptr, nullptr}},
        {nullptr, nullptr},
#endif
        {state_bwd_node_kernel<GNN_DP, 64, 128>, state_bwd_node_kernel<GNN_DP, 32, 32>},
#if GNN_DP >= 16 && GNN_DP <= 32
        state_bwd_node_l1_kernel<GNN_DP>,
#else
        nullptr,
#endif
        {state_bwd_scatter_kernel<GNN_DP, false>, state_bwd_scatter_kernel<GNN_DP, true>},
        bn_apply_kernel<GNN_DP>,
        bn_bwd_reduce_kernel<GNN_DP>,
    };
    return &set;
}

}  // namespace gnn
