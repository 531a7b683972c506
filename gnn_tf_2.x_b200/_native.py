# coding=utf-8
"""ctypes binding of ``libgnn_b200.so`` (C ABI declared in ``include/gnn_b200.h``).

PyTorch is used only as the owner of device memory and streams: every call hands raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream to the library.  There is NO fallback: if the shared library is
missing, or no CUDA device is present, the functions raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

GNN_MAX_LAYERS = 4
# GNN_B200_LIBRARY: another build of the same ABI (comparison runs of an older kernel on the same box)
_LIB_PATH = os.environ.get('GNN_B200_LIBRARY') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libgnn_b200.so')
_lib: Optional[C.CDLL] = None

ACT_CODES = {'linear': 0, 'relu': 1, 'tanh': 2, 'sigmoid': 3, 'selu': 4, 'elu': 5, 'softmax': 6, 'softplus': 7}


class gnn_mlp(C.Structure):
    _fields_ = [('n_layers', C.c_int32),
                ('dims', C.c_int32 * (GNN_MAX_LAYERS + 1)),
                ('act', C.c_int32 * GNN_MAX_LAYERS),
                ('W', C.c_void_p * GNN_MAX_LAYERS),
                ('b', C.c_void_p * GNN_MAX_LAYERS),
                ('drop_rate', C.c_float * (GNN_MAX_LAYERS + 1)),
                ('has_bn', C.c_int32),
                ('bn_gamma', C.c_void_p), ('bn_beta', C.c_void_p),
                ('bn_moving_mean', C.c_void_p), ('bn_moving_var', C.c_void_p),
                ('bn_eps', C.c_float), ('bn_momentum', C.c_float)]


class gnn_mlp_grad(C.Structure):
    _fields_ = [('dW', C.c_void_p * GNN_MAX_LAYERS), ('db', C.c_void_p * GNN_MAX_LAYERS),
                ('dgamma', C.c_void_p), ('dbeta', C.c_void_p)]


class gnn_graph(C.Structure):
    _fields_ = [('n_nodes', C.c_int64), ('n_arcs', C.c_int64),
                ('rowptr', C.c_void_p), ('col', C.c_void_p), ('val', C.c_void_p), ('row_scale', C.c_void_p),
                ('rowptr_T', C.c_void_p), ('col_T', C.c_void_p), ('val_T', C.c_void_p), ('max_block16_arcs', C.c_int32)]


class gnn_loop_args(C.Structure):
    _fields_ = [('D', C.c_int32), ('NL_self', C.c_int32), ('NL_agg', C.c_int32), ('AL', C.c_int32),
                ('x0', C.c_void_p), ('nodes', C.c_void_p), ('agg_nodes', C.c_void_p), ('agg_arcs', C.c_void_p),
                ('max_iter', C.c_int32), ('threshold', C.c_float), ('training', C.c_int32),
                ('save_for_backward', C.c_int32), ('seed', C.c_uint32),
                ('x_out', C.c_void_p), ('k_out', C.c_void_p),
                ('n_global', C.c_int64), ('row_offset', C.c_int64), ('exchange', C.c_void_p), ('exchange_user', C.c_void_p),
                ('n_peers', C.c_int32), ('rank', C.c_int32), ('peer_state', C.c_void_p * 8), ('peer_mask', C.c_void_p),
                ('sig_local', C.c_void_p), ('sig_peer', C.c_void_p * 8), ('sig_epoch', C.c_uint32),
                ('seed_dev', C.c_void_p)]


# callback type of gnn_loop_args.exchange: (user, t, x_next_offset, go_next_offset)
EXCHANGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int64, C.c_int64)


# every symbol include/gnn_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = ['gnn_last_error', 'gnn_abi_version', 'gnn_device_info', 'gnn_csr_build', 'gnn_spmm',
                    'gnn_state_loop_workspace_bytes', 'gnn_state_loop_layout', 'gnn_state_loop_forward', 'gnn_state_loop_backward',
                    'gnn_launch_count', 'gnn_profile_iterations', 'gnn_profile_last_iterations', 'gnn_last_forward_kernel',
                    'gnn_last_backward_kernel', 'gnn_output_dense']


def library_path() -> str: return _LIB_PATH


def lib() -> C.CDLL:
    """ load the shared library once; loud failure when it has not been built """
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f'{_LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               f'(make -C gnn_tf_2.x_b200/csrc). There is no CPU fallback.')
        l = C.CDLL(_LIB_PATH)
        l.gnn_last_error.restype = C.c_char_p
        l.gnn_last_forward_kernel.restype = C.c_char_p
        l.gnn_last_backward_kernel.restype = C.c_char_p
        l.gnn_abi_version.restype = C.c_int
        l.gnn_launch_count.restype = C.c_int64
        l.gnn_launch_count.argtypes = [C.c_int32]
        l.gnn_device_info.argtypes = [C.POINTER(C.c_int32)] * 4
        l.gnn_csr_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
        l.gnn_spmm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                               C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        l.gnn_output_dense.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        l.gnn_state_loop_workspace_bytes.argtypes = [C.POINTER(gnn_graph), C.POINTER(gnn_mlp), C.POINTER(gnn_loop_args),
                                                     C.POINTER(C.c_size_t)]
        l.gnn_state_loop_layout.argtypes = [C.POINTER(gnn_graph), C.POINTER(gnn_mlp), C.POINTER(gnn_loop_args),
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        l.gnn_state_loop_layout.restype = C.c_int
        l.gnn_state_loop_forward.argtypes = [C.POINTER(gnn_graph), C.POINTER(gnn_mlp), C.POINTER(gnn_loop_args),
                                             C.c_void_p, C.c_size_t, C.c_void_p]
        l.gnn_state_loop_backward.argtypes = [C.POINTER(gnn_graph), C.POINTER(gnn_mlp), C.POINTER(gnn_loop_args), C.c_void_p,
                                              C.POINTER(gnn_mlp_grad), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_size_t, C.c_void_p]
        l.gnn_profile_iterations.argtypes = [C.c_int32]
        l.gnn_profile_last_iterations.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        for name in ('gnn_profile_iterations', 'gnn_profile_last_iterations', 'gnn_device_info', 'gnn_csr_build', 'gnn_spmm', 'gnn_output_dense', 'gnn_state_loop_workspace_bytes',
                     'gnn_state_loop_forward', 'gnn_state_loop_backward'):
            getattr(l, name).restype = C.c_int
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f'{what} failed ({rc}): {lib().gnn_last_error().decode()}')


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('gnn_b200 needs a CUDA device (B200, sm_100a): there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def set_seed(args: 'gnn_loop_args', seed) -> None:
    """ dropout seed of a call: python int (by value) or an int64 CUDA tensor holding a value in [0, 2^32) -- the kernels then read
    its low 32 bits from device memory at run time (gnn_loop_args.seed_dev), which is what a CUDA-graph replay needs """
    if isinstance(seed, torch.Tensor):
        if seed.dtype != torch.int64 or not seed.is_cuda or seed.numel() != 1: raise TypeError('device seed: one int64 CUDA element')
        args.seed, args.seed_dev = 0, seed.data_ptr()
    else:
        args.seed, args.seed_dev = int(seed) & 0xFFFFFFFF, None


def last_forward_kernel() -> str:
    return lib().gnn_last_forward_kernel().decode()


def last_backward_kernel() -> str:
    return lib().gnn_last_backward_kernel().decode()


def profile_iterations(enable: bool) -> None:
    check(lib().gnn_profile_iterations(1 if enable else 0), 'gnn_profile_iterations')


def profile_last_iterations() -> tuple[float, int]:
    """ (milliseconds between the first and the last iteration-kernel launch of the last forward, launches) """
    ms, n = C.c_float(0), C.c_int32(0)
    check(lib().gnn_profile_last_iterations(C.byref(ms), C.byref(n)), 'gnn_profile_last_iterations')
    return float(ms.value), int(n.value)


def launch_count(reset: bool = False) -> int:
    return int(lib().gnn_launch_count(1 if reset else 0))


# ---------------------------------------------------------------------------------------------------------------------
def csr_from_coo_transposed(coo, *, device=None, with_transpose: bool = False):
    """ GraphTensor.COO2SparseTransposedTensor (graph_class.py:364-372) on the GPU: the COO matrix is transposed
    (new row = coo.col) and stored row-major, entries sorted by (row, col), ties in input order. """
    from .graph_class import SparseCSR
    device = default_device() if device is None else torch.device(device)
    rows = torch.as_tensor(np.ascontiguousarray(coo.col, dtype=np.int32), device=device)
    cols = torch.as_tensor(np.ascontiguousarray(coo.row, dtype=np.int32), device=device)
    vals = torch.as_tensor(np.ascontiguousarray(coo.data, dtype=np.float32), device=device)
    n_rows, n_cols = int(coo.shape[1]), int(coo.shape[0])
    return csr_build(rows, cols, vals, n_rows, n_cols, with_transpose=with_transpose)


def csr_build(rows: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, n_rows: int, n_cols: int, *,
              with_transpose: bool = False, assume_uniform: bool = False):
    """ gnn_csr_build on device int32/float32 tensors.  assume_uniform: the caller guarantees that every row holds one
    repeated value (default ArcNode of the three aggregation modes): no flag is read back, the call does not synchronise """
    from .graph_class import SparseCSR
    l = lib()
    device = rows.device
    if device.type != 'cuda': raise RuntimeError('csr_build needs CUDA tensors')
    nnz = int(rows.shape[0])
    i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=device)
    f32 = lambda n: torch.empty(max(n, 1), dtype=torch.float32, device=device)
    rowptr, col_s, val_s, perm, row_scale = i32(n_rows + 1), i32(nnz), f32(nnz), i32(nnz), f32(n_rows)
    rowptr_T = col_T = perm_T = val_T = None
    if with_transpose: rowptr_T, col_T, perm_T, val_T = i32(n_cols + 1), i32(nnz), i32(nnz), f32(nnz)
    uniform = C.c_int32(0)
    ws_bytes = C.c_size_t(0)
    with torch.cuda.device(device):
        args = [_ptr(rows), _ptr(cols), _ptr(vals), nnz, n_rows, n_cols, _ptr(rowptr), _ptr(col_s), _ptr(val_s), _ptr(perm),
                _ptr(row_scale), _ptr(rowptr_T), _ptr(col_T), _ptr(perm_T), _ptr(val_T), None if assume_uniform else C.byref(uniform)]
        check(l.gnn_csr_build(*args, None, C.byref(ws_bytes), _stream(device)), 'gnn_csr_build(size query)')
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=device)
        check(l.gnn_csr_build(*args, _ptr(ws), C.byref(ws_bytes), _stream(device)), 'gnn_csr_build')
    trim = lambda t, n: None if t is None else t[:n]
    return SparseCSR(rowptr[:n_rows + 1], trim(col_s, nnz), trim(val_s, nnz), trim(perm, nnz), (n_rows, n_cols),
                     rowptr_T=trim(rowptr_T, n_cols + 1), col_T=trim(col_T, nnz), perm_T=trim(perm_T, nnz),
                     values_T=trim(val_T, nnz), row_scale=row_scale[:n_rows] if (assume_uniform or uniform.value) else None)


def spmm(rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], dense: torch.Tensor,
         out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """ out[r] = sum_e val[e] * dense[col[e]] over the stored entries of row r (gnn_spmm) """
    l = lib()
    device = dense.device
    dense = dense.contiguous()
    n_rows, F = int(rowptr.shape[0]) - 1, int(dense.shape[1])
    if out is None: out = torch.empty((n_rows, F), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(l.gnn_spmm(_ptr(rowptr), _ptr(col), _ptr(val), n_rows, _ptr(dense), dense.stride(0) if F else 0, F, _ptr(out),
                         out.stride(0) if F else 0, 1 if accumulate else 0, _stream(device)), 'gnn_spmm')
    return out


def output_dense(x: Optional[torch.Tensor], labels: Optional[torch.Tensor], kernel: torch.Tensor, bias: torch.Tensor, activation: str) -> torch.Tensor:
    """ act([x | labels] @ kernel + bias) in one pass over the rows (gnn_output_dense); inference only (no autograd) """
    first = x if x is not None else labels
    n, device = int(first.shape[0]), first.device
    D = 0 if x is None else int(x.shape[1])
    NL = 0 if labels is None else int(labels.shape[1])
    T = int(kernel.shape[1])
    if x is not None and x.stride(1) != 1: x = x.contiguous()
    if labels is not None and labels.stride(1) != 1: labels = labels.contiguous()
    kernel, bias = kernel.detach().contiguous(), bias.detach().contiguous()
    out = torch.empty((n, T), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib().gnn_output_dense(_ptr(x) if D else None, n, D, x.stride(0) if D else 0, _ptr(labels) if NL else None, NL,
                                     labels.stride(0) if NL else 0, _ptr(kernel), _ptr(bias), T, ACT_CODES[activation], _ptr(out),
                                     _stream(device)), 'gnn_output_dense')
    return out


# ---------------------------------------------------------------------------------------------------------------------
def make_graph(adj) -> gnn_graph:
    """ gnn_graph view of a SparseCSR (Adjacency^T); uses the per-row weight when every row is uniform """
    g = gnn_graph()
    g.n_nodes, g.n_arcs = adj.dense_shape[0], adj.nnz
    g.rowptr, g.col = _ptr(adj.rowptr), _ptr(adj.col)
    if adj.row_scale is not None:
        g.val, g.row_scale, g.val_T = None, _ptr(adj.row_scale), None
    else:
        g.val, g.row_scale, g.val_T = _ptr(adj.values), None, _ptr(adj.values_T)
    g.rowptr_T, g.col_T = _ptr(adj.rowptr_T), _ptr(adj.col_T)
    g.max_block16_arcs = adj.max_block16_arcs
    return g


def make_mlp(spec, keepalive: list) -> gnn_mlp:
    """ gnn_mlp view of a keras_compat.MLPSpec; tensors that must outlive the call are appended to keepalive """
    m = gnn_mlp()
    L = spec.n_layers
    if L > GNN_MAX_LAYERS: raise NotImplementedError(f'at most {GNN_MAX_LAYERS} Dense layers per MLP are supported, got {L}')
    if spec.alpha_dropout and spec.has_dropout: raise NotImplementedError('AlphaDropout is not implemented in the CUDA path')
    m.n_layers = L
    for i, d in enumerate(spec.dims): m.dims[i] = int(d)
    for i in range(L):
        if spec.activations[i] not in ACT_CODES: raise NotImplementedError(f'activation {spec.activations[i]!r} not implemented')
        m.act[i] = ACT_CODES[spec.activations[i]]
        w = spec.dense_layers[i].kernel.detach().contiguous()
        b = spec.dense_layers[i].bias.detach().contiguous()
        keepalive += [w, b]
        m.W[i], m.b[i] = w.data_ptr(), b.data_ptr()
    for i in range(L + 1): m.drop_rate[i] = float(spec.dropout_rates[i])
    bn = spec.batchnorm
    m.has_bn = 0 if bn is None else 1
    if bn is not None:
        m.bn_gamma, m.bn_beta = bn.gamma.data_ptr(), bn.beta.data_ptr()
        m.bn_moving_mean, m.bn_moving_var = bn.moving_mean.data_ptr(), bn.moving_variance.data_ptr()
        m.bn_eps, m.bn_momentum = bn.epsilon, bn.momentum
    return m
