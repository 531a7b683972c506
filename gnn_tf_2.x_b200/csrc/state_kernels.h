// state_kernels.h -- kernel-pointer accessors; the templated kernels are instantiated per padded state width
// in state_inst.cu (compiled once per GNN_DP so that the build parallelises), the host code links against these.
#pragma once
#include "state_aux.cuh"
#include "state_bwd.cuh"
#include "state_bwd_l1.cuh"
#include "state_fwd.cuh"
#include "state_fwd_ws.cuh"
#include "state_fwd_tc.cuh"

namespace gnn {

typedef void (*IterKernel)(const IterParams);
typedef void (*BnApplyKernel)(const int*, int*, int*, int, const float*, const float*, const float*, long long, float, float*);
typedef void (*BwdNodeKernel)(const BwdNodeParams);
typedef void (*ScatterKernel)(const int*, int, const int32_t*, const int32_t*, const float*, long long, const float*, const float*, float*);
typedef void (*BnBwdReduceKernel)(const int*, int, const float*, const float*, const float*, float, long long, double*);

struct KernelSet {
    IterKernel iter[2][2];      // [tile: 0 = 128 nodes x 128 threads, 1 = 32 x 32][has_val]
    IterKernel iter_ws[2][2];   // warp-specialised pipeline [has_val][node-range partition with peers]; NULL when the width is not covered
    IterKernel iter_tc[2];      // tcgen05 / tensor-memory pipeline (row-scale graphs) [node-range partition with peers]; NULL when the width is not covered
    BwdNodeKernel bwd_node[2];  // [tile: 0 = 64 nodes x 128 threads, 1 = 32 x 32]
    BwdNodeKernel bwd_node_l1;  // single Dense layer, no dropout: pipelined 128-node tiles (state_bwd_l1.cuh); NULL when the width is not covered
    ScatterKernel scatter[2];   // [has_val]
    BnApplyKernel bn_apply;
    BnBwdReduceKernel bn_bwd_reduce;
};

#define GNN_DECLARE_KERNEL_SET(DPV) const KernelSet* kernel_set_dp##DPV();
GNN_DECLARE_KERNEL_SET(4)
GNN_DECLARE_KERNEL_SET(8)
GNN_DECLARE_KERNEL_SET(16)
GNN_DECLARE_KERNEL_SET(32)
GNN_DECLARE_KERNEL_SET(64)
GNN_DECLARE_KERNEL_SET(128)

static inline const KernelSet* kernel_set(int DP) {
    switch (DP) {
        case 4: return kernel_set_dp4();
        case 8: return kernel_set_dp8();
        case 16: return kernel_set_dp16();
        case 32: return kernel_set_dp32();
        case 64: return kernel_set_dp64();
        case 128: return kernel_set_dp128();
    }
    return nullptr;
}

}  // namespace gnn
