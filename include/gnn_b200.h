/*
 * gnn_b200.h -- C ABI of the B200-native Scarselli-GNN state-convergence engine (libgnn_b200.so).
 *
 * The reference (sailab-code/GNN_tf_2.x) has no FFI: its boundary is the Python method
 * GNNnodeBased.Loop(g, training) (GNN/GNN.py:251-280) and the third-party TensorFlow calls inside it.
 * Each entry point below replaces one of those call sites; the citation says which.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the CALLER (PyTorch); the library
 *     borrows it for the duration of the call and allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream); all work is enqueued on
 *     it and the call returns without synchronising unless stated;
 *   - return value: 0 on success, negative GNN_ERR_* otherwise; gnn_last_error() gives the message of
 *     the last failure on the calling thread; nothing is thrown across the boundary;
 *   - all floating point is fp32 (as the reference: graph_class.py:42-47), indices are int32 on the device.
 *   - matrices are row-major, dense, with the leading dimension given where it can differ from the width.
 */
#ifndef GNN_B200_H
#define GNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNN_B200_ABI_VERSION 6
#define GNN_MAX_LAYERS 4 /* Dense layers per MLP */
#define GNN_MAX_PEERS 8  /* GPUs of one NVSwitch domain */

enum gnn_error {
    GNN_OK = 0,
    GNN_ERR_INVALID = -1,     /* bad argument */
    GNN_ERR_UNSUPPORTED = -2, /* shape / option outside what the kernels implement */
    GNN_ERR_WORKSPACE = -3,   /* workspace too small */
    GNN_ERR_CUDA = -4         /* CUDA runtime error */
};

/* Keras activation strings accepted by MLP() (GNN/MLP.py:33) */
enum gnn_activation {
    GNN_ACT_LINEAR = 0, GNN_ACT_RELU = 1, GNN_ACT_TANH = 2, GNN_ACT_SIGMOID = 3,
    GNN_ACT_SELU = 4, GNN_ACT_ELU = 5, GNN_ACT_SOFTMAX = 6, GNN_ACT_SOFTPLUS = 7
};

const char* gnn_last_error(void);
int gnn_abi_version(void);
/* number of SMs and opt-in shared memory per block of the current device */
int gnn_device_info(int32_t* sm_count, int32_t* smem_optin_bytes, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------------------------------------------
 * Arc preprocessing.  Replaces GraphTensor.COO2SparseTransposedTensor (GNN/graph_class.py:364-372: python
 * list(zip(col,row)) -> tf.SparseTensor -> tf.sparse.reorder) and, for the backward pass, what TF derives
 * internally for the gradient of sparse_dense_matmul (A^T g).
 *
 * Input: COO entries (row[i], col[i], val[i]) of the ALREADY TRANSPOSED matrix (row = destination node).
 * Output: CSR in row-major order (entries sorted by (row, col), ties in input order = stable):
 *   rowptr[n_rows+1], col_sorted[nnz], val_sorted[nnz], perm[nnz] (input position of each stored entry),
 *   row_scale[n_rows] = value of the first entry of each row (0 for empty rows), *rows_uniform = 1 when every
 *   row holds one repeated value (true for 'sum' / 'average' / 'normalized' aggregation).
 * Optional transposed structure (pass NULL rowptr_T to skip): entries sorted by (col, CSR position):
 *   rowptr_T[n_cols+1], col_T[nnz] (= row of the entry), perm_T[nnz] (= CSR position), val_T[nnz].
 * Call with workspace == NULL to get the required size in *workspace_bytes.
 * Synchronises the stream once (to report rows_uniform to the host); pass rows_uniform == NULL to skip the report and
 * the synchronisation (e.g. for the default ArcNode of the three aggregation modes, whose rows are uniform by construction).
 */
int gnn_csr_build(const int32_t* row, const int32_t* col, const float* val, /* device [nnz] */
                  int64_t nnz, int64_t n_rows, int64_t n_cols,
                  int32_t* rowptr, int32_t* col_sorted, float* val_sorted, int32_t* perm, float* row_scale,
                  int32_t* rowptr_T, int32_t* col_T, int32_t* perm_T, float* val_T,
                  int32_t* rows_uniform, /* host out */
                  void* workspace, size_t* workspace_bytes, void* stream);

/* Sparse x dense, out[r, 0:F] = sum_e val[e] * dense[col[e], 0:F] over the stored entries of row r in stored
 * order (deterministic, no atomics).  Replaces tf.sparse.sparse_dense_matmul at GNN/GNN.py:259 (ArcNode^T x arc
 * labels) and :263 (Adjacency^T x node labels); with the transposed structure it is also their gradient.
 * val may be NULL (all ones).  accumulate != 0 adds to `out` instead of overwriting it. */
int gnn_spmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_rows,
             const float* dense, int64_t ld_dense, int32_t F, float* out, int64_t ld_out, int32_t accumulate,
             void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * MLP description: what GNN/MLP.py:11-64 builds -- Dense chain, Dropout in front of Dense l (drop_rate[l]) or
 * behind the last Dense (drop_rate[n_layers]), optional trailing BatchNormalization (Keras defaults).
 */
typedef struct gnn_mlp {
    int32_t n_layers;                    /* 1..GNN_MAX_LAYERS */
    int32_t dims[GNN_MAX_LAYERS + 1];    /* dims[0] = input width, dims[l+1] = units of Dense l */
    int32_t act[GNN_MAX_LAYERS];         /* gnn_activation */
    const float* W[GNN_MAX_LAYERS];      /* device, Keras kernel layout [dims[l], dims[l+1]] */
    const float* b[GNN_MAX_LAYERS];      /* device [dims[l+1]] */
    float drop_rate[GNN_MAX_LAYERS + 1]; /* 0 = no dropout at that position */
    int32_t has_bn;
    const float* bn_gamma;               /* device [dims[n_layers]] */
    const float* bn_beta;
    float* bn_moving_mean;               /* device; updated in place once per loop iteration when training */
    float* bn_moving_var;
    float bn_eps;
    float bn_momentum;
} gnn_mlp;

/* gradients of the trainable variables, Keras layouts; every non-NULL buffer is OVERWRITTEN */
typedef struct gnn_mlp_grad {
    float* dW[GNN_MAX_LAYERS];
    float* db[GNN_MAX_LAYERS];
    float* dgamma;
    float* dbeta;
} gnn_mlp_grad;

/* destination-sorted CSR of Adjacency^T (+ its transpose), as produced by gnn_csr_build */
typedef struct gnn_graph {
    int64_t n_nodes;
    int64_t n_arcs;
    const int32_t* rowptr;    /* [n_nodes+1] */
    const int32_t* col;       /* [n_arcs] source node of every stored arc */
    const float* val;         /* [n_arcs] per-arc weight; NULL => use row_scale */
    const float* row_scale;   /* [n_nodes] common weight of the arcs entering a node; NULL => use val */
    const int32_t* rowptr_T;  /* [n_nodes+1] (backward only) */
    const int32_t* col_T;     /* [n_arcs] destination node, source-sorted */
    const float* val_T;       /* [n_arcs] weights in transposed order; NULL with row_scale */
    int32_t max_block16_arcs; /* largest rowptr[16(b+1)] - rowptr[16b] over the aligned blocks of 16 rows (the last block may
                                 be shorter); 0 = unknown.  Sizes the slots of the forward kernel's landing ring exactly */
} gnn_graph;

/* ------------------------------------------------------------------------------------------------------------
 * The state-convergence loop.  Replaces tf.while_loop(self.condition, self.convergence, ...) at
 * GNN/GNN.py:271-272 together with condition (:202-220) and convergence (:223-242), and -- for
 * gnn_state_loop_backward -- what tf.GradientTape replays for it (GNN/GNN_BaseClass.py:233-237).
 *
 * net_state input columns, in the reference's order (GNN.py:228-237):
 *   [ state(D) | nodes(NL_self) | agg_state(D) | agg_nodes(NL_agg) | agg_arcs(AL) ],  D = state width
 *   (NL_self = NL_agg = NL when state_vect_dim > 0, else 0).
 * One iteration: agg_state = Adjacency^T x state (segment sum over incoming arcs, stored order),
 * state_new = net_state(input).  Iterations run while any node has
 * ||state - state_old||_2 > threshold * ||state_old||_2 and k < max_iter; state_old starts as ones (:266).
 * The whole loop is enqueued without host synchronisation; *k_out (device float) is the iteration count.
 */
typedef struct gnn_loop_args {
    int32_t D, NL_self, NL_agg, AL;
    const float* x0;          /* device [N, D] */
    const float* nodes;       /* device [N, NL_self] (unused when NL_self == 0) */
    const float* agg_nodes;   /* device [N, NL_agg] */
    const float* agg_arcs;    /* device [N, AL] */
    int32_t max_iter;
    float threshold;
    int32_t training;         /* Keras `training` flag: dropout active, BatchNormalization uses batch statistics */
    int32_t save_for_backward;/* keep the iterates in the workspace so that gnn_state_loop_backward can run */
    uint32_t seed;            /* dropout generator seed of this call */
    float* x_out;             /* device [N, D] converged state ([n_global, D] when partitioned) */
    float* k_out;             /* device [1] number of iterations, float32 as GNN.py:267 */
    /* Node-range partition of one graph over several GPUs (all zero / NULL on a single GPU).  The graph handed to the
     * call holds the rows (destination nodes) [row_offset, row_offset + g->n_nodes) of a graph with n_global nodes;
     * its column indices are GLOBAL node ids.  x0 and x_out are full [n_global, D] arrays (replicated), nodes /
     * agg_nodes / agg_arcs hold the local rows only.  After the kernels of iteration t have been enqueued the library
     * calls exchange(user, t, x_next_offset, go_next_offset): byte offsets INTO THE WORKSPACE of the full state buffer
     * [n_global, DP] whose local rows were just written and of the int32 flag of iteration t+1 (-1 after the last
     * iteration).  The callback enqueues, on the same stream, the exchange of the boundary rows and the max-reduction
     * of the flag (e.g. NCCL through torch.distributed), so that every rank runs the same number of iterations.
     * Partitioned calls are forward-only (training = 0, save_for_backward = 0). */
    int64_t n_global;
    int64_t row_offset;
    void (*exchange)(void* user, int32_t t, int64_t x_next_offset, int64_t go_next_offset);
    void* exchange_user;
    /* Fused exchange over NVLink peer memory (optional, partitioned calls only).  When n_peers > 1 the workspace of
     * every rank lives in peer-mapped (symmetric) memory and peer_state[r] is the address, valid on THIS device, of the
     * state-buffer area of rank r's workspace (workspace_r + state offset from gnn_state_loop_layout).  The iteration
     * kernel then stores every new state row it produces directly into the buffers of the peers that gather from it
     * (peer_mask[local row] bit r; NULL = every peer needs every row), overlapping the transfer with the MLP of the
     * following tiles; the exchange callback only has to max-reduce the flag (which also orders the iterations). */
    int32_t n_peers;
    int32_t rank;
    float* peer_state[GNN_MAX_PEERS];
    const uint32_t* peer_mask;
    /* In-kernel signalling (optional, with the fused exchange): no collective and no callback between the iterations.
     * sig_local / sig_peer[r] address one signal area per rank in peer-mapped memory (sig_peer[r] = rank r's area as seen from
     * THIS device; sig_local = this rank's own), each 2 * 8 * (max_iter + 1) int32, zero-filled when allocated and never reset:
     * arrive[t][r] | flag[t][r].  The last CTA of iteration t on rank r writes its convergence flag and then the call's
     * sig_epoch (release, system scope) into slot [t + 1][r] of every rank; the CTAs of iteration t + 1 wait until all ranks'
     * marks have reached sig_epoch (compared modulo 2^32: use a counter that grows with every partitioned call of the group)
     * and OR the flags, so every rank runs the same iterations.  exchange may then be NULL. */
    int32_t* sig_local;
    int32_t* sig_peer[GNN_MAX_PEERS];
    uint32_t sig_epoch;
    /* Optional: device [1].  When non-NULL the dropout generator takes the seed of the call from device memory at kernel
     * time instead of `seed`: a forward / backward pair captured in a CUDA graph (arguments frozen at capture) then draws
     * new masks at every replay as long as the caller advances the value between replays. */
    const uint32_t* seed_dev;
} gnn_loop_args;

int gnn_state_loop_workspace_bytes(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, size_t* bytes);

/* where the state buffers sit inside the workspace: byte offset of the first one and bytes per buffer ([rows, DP] fp32,
 * rows = n_global when partitioned); forward-only calls ping-pong between buffer 0 and 1 */
int gnn_state_loop_layout(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, size_t* state_offset, size_t* state_bytes);

int gnn_state_loop_forward(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a,
                           void* workspace, size_t workspace_bytes, void* stream);

/* BPTT through the iterations saved by the matching forward call (same g/net/a/workspace).
 * g_x: device [N, D] gradient of the loss wrt the converged state (x_out).
 * Outputs (NULL to skip): grad of the net_state variables, g_x0 [N, D], g_nodes [N, NL_self],
 * g_agg_nodes [N, NL_agg], g_agg_arcs [N, AL] (the last three are needed only by LGNN: LGNN.py:258-259). */
int gnn_state_loop_backward(const gnn_graph* g, const gnn_mlp* net, const gnn_loop_args* a, const float* g_x,
                            gnn_mlp_grad* grad, float* g_x0, float* g_nodes, float* g_agg_nodes, float* g_agg_arcs,
                            void* workspace, size_t workspace_bytes, void* stream);

/* number of kernels the library has launched since the last reset (for bench.py's gpu_launches) */
int64_t gnn_launch_count(int32_t reset);

/* Measurement aid for bench.py's roofline: when enabled, gnn_state_loop_forward brackets the sequence of
 * iteration-kernel launches (not the prologue / epilogue kernels) with CUDA events on the launching stream.
 * gnn_profile_last_iterations waits for the closing event and returns the elapsed milliseconds and the number of
 * iteration launches enqueued between the two events (launches that found the loop stopped return at once). */
int gnn_profile_iterations(int32_t enable);
int gnn_profile_last_iterations(float* elapsed_ms, int32_t* launches);
/* out[n_rows, T] = act([x[n_rows, D] | labels[n_rows, NL]] @ W[D + NL, T] + b[T]) : the output net of an inference Loop when it is ONE
 * Dense layer over every node (replaces the tf.concat + Keras call of GNN/GNN.py:245-248, 279).  Row-major, ld_* in floats; W in
 * Keras order (input x output); act = GNN_ACT_* incl. softmax; T <= 16.  Enqueues one kernel on the stream. */
int gnn_output_dense(const float* x, int64_t n_rows, int32_t D, int64_t ld_x, const float* labels, int32_t NL, int64_t ld_labels,
                     const float* W, const float* b, int32_t T, int32_t act, float* out, void* stream);

/* name of the iteration kernel the last gnn_state_loop_forward call launched ("" before the first call) */
const char* gnn_last_forward_kernel(void);
/* name of the node kernel the last gnn_state_loop_backward call launched ("" before the first call) */
const char* gnn_last_backward_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* GNN_B200_H */
