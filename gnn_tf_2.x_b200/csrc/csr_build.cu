// csr_build.cu -- arc preprocessing on the GPU: COO (already transposed: row = destination) -> row-major CSR
// in the order tf.sparse.reorder produces (reference GNN/graph_class.py:364-372), plus the source-sorted
// transposed structure used by the backward pass, plus the per-row weight when rows are uniform.
//
// Sorting is one stable LSD radix sort of 64-bit keys (row << 32 | col) carrying the input position
// (cub::DeviceRadixSort -- toolkit header library, used as plumbing for the once-per-graph preprocessing);
// the rest (key packing, row pointers with empty rows, gathers, uniformity test) are our kernels.
#include <cub/device/device_radix_sort.cuh>

#include <atomic>

#include "common.cuh"

namespace gnn {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};   // process-wide: the backward pass runs on autograd's own thread
void set_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches += n; }

namespace {

constexpr int kThreads = 256;

__global__ void pack_keys_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t nnz,
                                 uint64_t* __restrict__ keys, int32_t* __restrict__ iota) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    keys[i] = ((uint64_t)(uint32_t)row[i] << 32) | (uint32_t)col[i];
    iota[i] = (int32_t)i;
}

// rowptr from sorted row ids: entry j opens every row in (row[j-1], row[j]]; the tail rows get nnz.
// rows_of: functor giving the row of sorted entry j
template <typename RowOf>
__global__ void rowptr_kernel(RowOf row_of, int64_t nnz, int64_t n_rows, int32_t* __restrict__ rowptr) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j > nnz) return;
    int64_t prev = (j == 0) ? -1 : (int64_t)row_of(j - 1);
    int64_t cur = (j == nnz) ? n_rows : (int64_t)row_of(j);
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)j;
}

struct RowFromKey {
    const uint64_t* keys;
    __device__ int32_t operator()(int64_t j) const { return (int32_t)(keys[j] >> 32); }
};
struct RowFromArray {
    const int32_t* rows;
    __device__ int32_t operator()(int64_t j) const { return rows[j]; }
};

__global__ void unpack_sorted_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ perm,
                                     const float* __restrict__ val, int64_t nnz, int32_t* __restrict__ col_sorted,
                                     float* __restrict__ val_sorted, int32_t* __restrict__ rows_sorted,
                                     int32_t* __restrict__ iota) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    uint64_t k = keys[j];
    col_sorted[j] = (int32_t)(uint32_t)k;
    if (val_sorted) val_sorted[j] = val ? val[perm[j]] : 1.0f;
    if (rows_sorted) rows_sorted[j] = (int32_t)(k >> 32);
    if (iota) iota[j] = (int32_t)j;
}

__global__ void transpose_fill_kernel(const int32_t* __restrict__ perm_T, const int32_t* __restrict__ rows_sorted,
                                      const float* __restrict__ val_sorted, int64_t nnz, int32_t* __restrict__ col_T,
                                      float* __restrict__ val_T) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int32_t j = perm_T[p];
    col_T[p] = rows_sorted[j];
    if (val_T) val_T[p] = val_sorted[j];
}

__global__ void row_scale_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ val_sorted,
                                 int64_t n_rows, float* __restrict__ row_scale, int32_t* __restrict__ not_uniform) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int32_t e0 = rowptr[r], e1 = rowptr[r + 1];
    float first = e0 < e1 ? val_sorted[e0] : 0.f;
    bool differs = false;
    for (int32_t e = e0 + 1; e < e1; ++e) differs |= (val_sorted[e] != first);
    if (row_scale) row_scale[r] = first;
    if (differs) atomicOr(not_uniform, 1);
}

int end_bit_for(int64_t n) {
    int bits = 1;
    while (((int64_t)1 << bits) < n && bits < 32) ++bits;
    return bits;
}

}  // namespace
}  // namespace gnn

using namespace gnn;

extern "C" const char* gnn_last_error(void) { return g_last_error.c_str(); }
extern "C" int gnn_abi_version(void) { return GNN_B200_ABI_VERSION; }
extern "C" int64_t gnn_launch_count(int32_t reset) {
    int64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

extern "C" int gnn_device_info(int32_t* sm_count, int32_t* smem_optin_bytes, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    GNN_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) { GNN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (smem_optin_bytes) { GNN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)); *smem_optin_bytes = v; }
    if (cc_major) { GNN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { GNN_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return GNN_OK;
}

extern "C" int gnn_csr_build(const int32_t* row, const int32_t* col, const float* val, int64_t nnz, int64_t n_rows,
                             int64_t n_cols, int32_t* rowptr, int32_t* col_sorted, float* val_sorted, int32_t* perm,
                             float* row_scale, int32_t* rowptr_T, int32_t* col_T, int32_t* perm_T, float* val_T,
                             int32_t* rows_uniform, void* workspace, size_t* workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nnz < 0 || n_rows < 0 || n_cols < 0 || !workspace_bytes) GNN_FAIL(GNN_ERR_INVALID, "gnn_csr_build: bad sizes");
    if (nnz >= (int64_t)INT32_MAX || n_rows >= (int64_t)INT32_MAX || n_cols >= (int64_t)INT32_MAX)
        GNN_FAIL(GNN_ERR_UNSUPPORTED, "gnn_csr_build: int32 index range exceeded");
    const bool want_T = rowptr_T != nullptr;

    // workspace carving: keys in/out (u64), iota/perm scratch (i32), rows_sorted (i32), flag, cub temp
    size_t cub_bytes_a = 0, cub_bytes_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes_a, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)nnz, 0, 64, stream);
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes_b, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)nnz, 0, 32, stream);
    size_t cub_bytes = cub_bytes_a > cub_bytes_b ? cub_bytes_a : cub_bytes_b;
    size_t n8 = align_up((size_t)(nnz > 0 ? nnz : 1) * 8, 256), n4 = align_up((size_t)(nnz > 0 ? nnz : 1) * 4, 256);
    size_t need = 2 * n8 + 4 * n4 + 256 + align_up(cub_bytes, 256);
    if (!workspace) {
        *workspace_bytes = need;
        return GNN_OK;
    }
    if (*workspace_bytes < need) GNN_FAIL(GNN_ERR_WORKSPACE, "gnn_csr_build: workspace %zu < %zu", *workspace_bytes, need);
    if (!row || !col || !rowptr || !col_sorted || !val_sorted || !perm)
        if (nnz > 0 || !rowptr) GNN_FAIL(GNN_ERR_INVALID, "gnn_csr_build: NULL argument");

    char* w = (char*)workspace;
    uint64_t* keys_in = (uint64_t*)w; w += n8;
    uint64_t* keys_out = (uint64_t*)w; w += n8;
    int32_t* iota = (int32_t*)w; w += n4;
    int32_t* rows_sorted = (int32_t*)w; w += n4;
    int32_t* colkey_out = (int32_t*)w; w += n4;
    int32_t* perm_tmp = (int32_t*)w; w += n4;
    int32_t* flag = (int32_t*)w; w += 256;
    void* cub_tmp = (void*)w;

    GNN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), stream));
    const int grid_nnz = (int)ceil_div(nnz > 0 ? nnz : 1, kThreads);
    const int grid_nnz1 = (int)ceil_div(nnz + 1, kThreads);

    if (nnz > 0) {
        pack_keys_kernel<<<grid_nnz, kThreads, 0, stream>>>(row, col, nnz, keys_in, iota);
        GNN_LAUNCH_CHECK();
        int end_bit = 32 + end_bit_for(n_rows);
        GNN_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, iota, perm, (int)nnz, 0, end_bit, stream));
        count_launch(4);
        unpack_sorted_kernel<<<grid_nnz, kThreads, 0, stream>>>(keys_out, perm, val, nnz, col_sorted, val_sorted, rows_sorted, iota);
        GNN_LAUNCH_CHECK();
    }
    rowptr_kernel<<<grid_nnz1, kThreads, 0, stream>>>(RowFromKey{keys_out}, nnz, n_rows, rowptr);
    GNN_LAUNCH_CHECK();
    if (n_rows > 0) {
        row_scale_kernel<<<(int)ceil_div(n_rows, kThreads), kThreads, 0, stream>>>(rowptr, val_sorted, n_rows, row_scale, flag);
        GNN_LAUNCH_CHECK();
    }

    if (want_T) {
        if (!col_T || !perm_T) GNN_FAIL(GNN_ERR_INVALID, "gnn_csr_build: transposed outputs incomplete");
        if (nnz > 0) {
            // stable sort of CSR positions by source column: ties keep ascending CSR position
            GNN_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, col_sorted, colkey_out, iota, perm_tmp, (int)nnz, 0,
                                                     end_bit_for(n_cols), stream));
            count_launch(3);
            GNN_CUDA(cudaMemcpyAsync(perm_T, perm_tmp, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, stream));
            transpose_fill_kernel<<<grid_nnz, kThreads, 0, stream>>>(perm_T, rows_sorted, val_sorted, nnz, col_T, val_T);
            GNN_LAUNCH_CHECK();
        }
        rowptr_kernel<<<grid_nnz1, kThreads, 0, stream>>>(RowFromArray{colkey_out}, nnz, n_cols, rowptr_T);
        GNN_LAUNCH_CHECK();
    }

    if (!rows_uniform) return GNN_OK;   // the caller knows (or does not care): nothing to report, no synchronisation
    int32_t host_flag = 0;
    GNN_CUDA(cudaMemcpyAsync(&host_flag, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    GNN_CUDA(cudaStreamSynchronize(stream));
    *rows_uniform = host_flag ? 0 : 1;
    return GNN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// SpMM with a narrow dense operand (prologue: arc labels / node labels; GNN/GNN.py:259,263)
// one thread per (row, 4-column chunk); entries accumulated in stored order
// ---------------------------------------------------------------------------------------------------------
namespace gnn {
namespace {
__global__ void spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                            const float* __restrict__ val, int64_t n_rows, const float* __restrict__ dense,
                            int64_t ld_dense, int F, float* __restrict__ out, int64_t ld_out, int accumulate) {
    const int chunks = (F + 3) / 4;
    int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (item >= n_rows * chunks) return;
    int64_t r = item / chunks;
    int c0 = (int)(item % chunks) * 4;
    int nc = min(4, F - c0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int32_t e0 = rowptr[r], e1 = rowptr[r + 1];
    for (int32_t e = e0; e < e1; ++e) {
        float w = val ? val[e] : 1.f;
        const float* src = dense + (int64_t)col[e] * ld_dense + c0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < nc) acc[c] = fmaf(w, __ldg(src + c), acc[c]);
    }
    float* dst = out + r * ld_out + c0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nc) dst[c] = accumulate ? dst[c] + acc[c] : acc[c];
}

// out[n, 0..T) = act([x[n, 0..D) | labels[n, 0..NL)] @ W[D + NL, T] + b): the output net of an inference Loop when it is one Dense layer
// (GNN/GNN.py:275-279 with the reference's default output MLP).  OUT_LPR lanes per row: the lanes of a row read consecutive 16-byte
// pieces (one 128-byte line per row at D = 32), keep partial sums for all T outputs and add them with a fixed xor-shuffle tree
// (deterministic); lane 0 adds the bias, applies the activation (softmax in registers) and stores.  Weights in shared memory.
// One pass over the state instead of cat + GEMM + bias + softmax.  (Thread-per-row was measured first: 105 us for 1M rows of 35
// floats, every lane walking its own 128-byte line.)
// MAXT: compile-time bound of T (2 / 4 / 8 / 16): with one bound of 16 the predicated-off multiply-adds of a two-unit net were most
// of the instructions and the kernel took 246 us.
constexpr int OUT_MAXT = 16;
constexpr int OUT_LPR = 8;
constexpr int OUT_RPG = 8;      // rows per lane group and block (OUT_ROWS_PER_LANE_GROUP)
template <int MAXT>
__global__ void output_dense_kernel(const float* __restrict__ x, long long n, int D, long long ldx, const float* __restrict__ labels, int NL,
                                    long long ldl, const float* __restrict__ W, const float* __restrict__ b, int T, int act,
                                    float* __restrict__ out) {
    extern __shared__ float sw[];       // W [D + NL][T], b [T]
    const int F = D + NL;
    for (int i = threadIdx.x; i < F * T; i += blockDim.x) sw[i] = W[i];
    for (int i = threadIdx.x; i < T; i += blockDim.x) sw[F * T + i] = b[i];
    __syncthreads();
    const int l = threadIdx.x % OUT_LPR;
    // OUT_ROWS_PER_LANE_GROUP rows per lane group, one after the other: the weights are staged once per 256 rows, not once per 32
    for (int rr = 0; rr < OUT_RPG; ++rr) {
    const long long row = (blockIdx.x * (long long)OUT_RPG + rr) * (blockDim.x / OUT_LPR) + threadIdx.x / OUT_LPR;
    const bool valid = row < n;
    float acc[MAXT];
#pragma unroll
    for (int o = 0; o < MAXT; ++o) acc[o] = 0.f;
    if (valid) {
        const float* xr = x + row * ldx;
        const bool vec = (D % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
        if (vec) {
            for (int q = l; q < D / 4; q += OUT_LPR) {
                const float4 v4 = ldg4(xr + 4 * q);
                const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* wr = sw + (size_t)(4 * q + c) * T;
#pragma unroll
                    for (int o = 0; o < MAXT; ++o) if (o < T) acc[o] = fmaf(v[c], wr[o], acc[o]);
                }
            }
        } else {
            for (int j = l; j < D; j += OUT_LPR) {
                const float v = __ldg(xr + j);
                const float* wr = sw + (size_t)j * T;
#pragma unroll
                for (int o = 0; o < MAXT; ++o) if (o < T) acc[o] = fmaf(v, wr[o], acc[o]);
            }
        }
        for (int j = l; j < NL; j += OUT_LPR) {
            const float v = __ldg(labels + row * ldl + j);
            const float* wr = sw + (size_t)(D + j) * T;
#pragma unroll
            for (int o = 0; o < MAXT; ++o) if (o < T) acc[o] = fmaf(v, wr[o], acc[o]);
        }
    }
#pragma unroll
    for (int off = OUT_LPR / 2; off > 0; off >>= 1)
#pragma unroll
        for (int o = 0; o < MAXT; ++o)
            if (o < T) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);      // (T is uniform: every lane takes the same branch)
    if (!valid || l != 0) continue;
#pragma unroll
    for (int o = 0; o < MAXT; ++o) if (o < T) acc[o] += sw[F * T + o];
    if (act == GNN_ACT_SOFTMAX) {
        float m = -INFINITY, sum = 0.f;
#pragma unroll
        for (int o = 0; o < MAXT; ++o) if (o < T) m = fmaxf(m, acc[o]);
#pragma unroll
        for (int o = 0; o < MAXT; ++o) if (o < T) { acc[o] = expf(acc[o] - m); sum += acc[o]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int o = 0; o < MAXT; ++o) acc[o] *= inv;
    } else {
#pragma unroll
        for (int o = 0; o < MAXT; ++o) acc[o] = act_apply(act, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < MAXT; ++o) if (o < T) out[row * T + o] = acc[o];
    }
}
}  // namespace
}  // namespace gnn

extern "C" int gnn_spmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, const float* dense,
                        int64_t ld_dense, int32_t F, float* out, int64_t ld_out, int32_t accumulate, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows < 0 || F < 0) GNN_FAIL(GNN_ERR_INVALID, "gnn_spmm: bad sizes");
    if (n_rows == 0 || F == 0) return GNN_OK;
    if (!rowptr || !out) GNN_FAIL(GNN_ERR_INVALID, "gnn_spmm: NULL argument");  // col / dense may be NULL when there are no entries
    int64_t items = n_rows * ((F + 3) / 4);
    spmm_kernel<<<(unsigned)ceil_div(items, 256), 256, 0, stream>>>(rowptr, col, val, n_rows, dense, ld_dense, F, out, ld_out, accumulate);
    GNN_LAUNCH_CHECK();
    return GNN_OK;
}

extern "C" int gnn_output_dense(const float* x, int64_t n_rows, int32_t D, int64_t ld_x, const float* labels, int32_t NL, int64_t ld_labels,
                                const float* W, const float* b, int32_t T, int32_t act, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows < 0 || D < 0 || NL < 0 || D + NL < 1 || T < 1) GNN_FAIL(GNN_ERR_INVALID, "gnn_output_dense: bad sizes");
    if (T > gnn::OUT_MAXT) GNN_FAIL(GNN_ERR_UNSUPPORTED, "gnn_output_dense: at most %d output units", gnn::OUT_MAXT);
    if (act < GNN_ACT_LINEAR || act > GNN_ACT_SOFTPLUS) GNN_FAIL(GNN_ERR_INVALID, "gnn_output_dense: unknown activation %d", act);
    if (n_rows == 0) return GNN_OK;
    if ((D > 0 && !x) || (NL > 0 && !labels) || !W || !b || !out) GNN_FAIL(GNN_ERR_INVALID, "gnn_output_dense: NULL argument");
    const size_t smem = ((size_t)(D + NL) * T + T) * sizeof(float);
    if (smem > 48 * 1024) GNN_FAIL(GNN_ERR_UNSUPPORTED, "gnn_output_dense: %zu bytes of weights do not fit", smem);
    const unsigned grid = (unsigned)ceil_div(n_rows, (256 / gnn::OUT_LPR) * gnn::OUT_RPG);
    if (T <= 2) gnn::output_dense_kernel<2><<<grid, 256, smem, stream>>>(x, n_rows, D, ld_x, labels, NL, ld_labels, W, b, T, act, out);
    else if (T <= 4) gnn::output_dense_kernel<4><<<grid, 256, smem, stream>>>(x, n_rows, D, ld_x, labels, NL, ld_labels, W, b, T, act, out);
    else if (T <= 8) gnn::output_dense_kernel<8><<<grid, 256, smem, stream>>>(x, n_rows, D, ld_x, labels, NL, ld_labels, W, b, T, act, out);
    else gnn::output_dense_kernel<16><<<grid, 256, smem, stream>>>(x, n_rows, D, ld_x, labels, NL, ld_labels, W, b, T, act, out);
    GNN_LAUNCH_CHECK();
    return GNN_OK;
}
