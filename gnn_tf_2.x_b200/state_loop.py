# coding=utf-8
"""Autograd glue around the C-ABI state loop: one ``torch.autograd.Function`` per ``Loop`` call.

``state_loop(...)`` replaces ``tf.while_loop(self.condition, self.convergence, ...)`` (GNN/GNN.py:271-272) and
``sparse_dense(...)`` replaces ``tf.sparse.sparse_dense_matmul`` (GNN/GNN.py:259,263).  Forward and backward both run
entirely inside ``libgnn_b200.so``; PyTorch only allocates the buffers and records the graph of tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as N
from .graph_class import SparseCSR
from .keras_compat import Sequential


class _SparseDense(torch.autograd.Function):
    """ out = S @ dense with S a row-major SparseCSR (already transposed matrix of the reference) """

    @staticmethod
    def forward(ctx, sp: SparseCSR, dense: torch.Tensor):
        ctx.sp = sp
        return N.spmm(sp.rowptr, sp.col, sp.values, dense)

    @staticmethod
    def backward(ctx, g_out):
        sp = ctx.sp
        g_out = g_out.contiguous()
        if sp.rowptr_T is not None:
            return None, N.spmm(sp.rowptr_T, sp.col_T, sp.values_T, g_out)
        # no transposed structure (ArcNode): every column holds at most a few entries -> index_add
        rows = sp.indices[:, 0]
        g = torch.zeros((sp.dense_shape[1], g_out.shape[1]), dtype=torch.float32, device=g_out.device)
        g.index_add_(0, sp.col.to(torch.int64), sp.values[:, None] * g_out[rows])
        return None, g


def sparse_dense(sp: SparseCSR, dense: torch.Tensor) -> torch.Tensor:
    if dense.shape[1] == 0:
        return torch.zeros((sp.dense_shape[0], 0), dtype=torch.float32, device=dense.device)
    return _SparseDense.apply(sp, dense.to(torch.float32))


class _SegmentPool(torch.autograd.Function):
    """ NodeGraph^T @ out_nodes for a block-diagonal NodeGraph kept as (graph id, coefficient) per node, graph ids
    non-decreasing: a CSR over graphs evaluated by gnn_spmm (stored order, deterministic); GNN.py:331-332 """

    @staticmethod
    def forward(ctx, rowptr, col, coeff, ids, out_nodes):
        ctx.save_for_backward(coeff, ids)
        return N.spmm(rowptr, col, coeff, out_nodes)

    @staticmethod
    def backward(ctx, g_pooled):
        coeff, ids = ctx.saved_tensors
        return None, None, None, None, g_pooled.index_select(0, ids) * coeff[:, None]


def segment_pool(rowptr, col, coeff, ids, out_nodes):
    return _SegmentPool.apply(rowptr, col, coeff, ids, out_nodes.to(torch.float32))


class _LoopConfig:
    __slots__ = ('adj', 'net', 'spec', 'D', 'NL_self', 'NL_agg', 'AL', 'max_iter', 'threshold', 'training', 'seed', 'save', 'partition')


class _StateLoop(torch.autograd.Function):

    @staticmethod
    def forward(ctx, cfg: _LoopConfig, x0, nodes, agg_nodes, agg_arcs, *params):
        lib = N.lib()
        device = x0.device
        n_nodes = int(x0.shape[0])   # rows of the state: all nodes of the graph (also when partitioned)
        x0c = x0.detach().contiguous()
        nodes_c = None if nodes is None else nodes.detach().contiguous()
        agg_nodes_c, agg_arcs_c = agg_nodes.detach().contiguous(), agg_arcs.detach().contiguous()
        save = cfg.save   # decided by the caller: grad mode is always off inside Function.forward
        part = cfg.partition

        keep = []
        mlp = N.make_mlp(cfg.spec, keep)
        graph = N.make_graph(cfg.adj)
        x_out = torch.empty((n_nodes, cfg.D), dtype=torch.float32, device=device)
        k_out = torch.zeros((), dtype=torch.float32, device=device)
        args = N.gnn_loop_args()
        args.D, args.NL_self, args.NL_agg, args.AL = cfg.D, cfg.NL_self, cfg.NL_agg, cfg.AL
        args.x0, args.nodes = x0c.data_ptr(), None if nodes_c is None else nodes_c.data_ptr()
        args.agg_nodes, args.agg_arcs = agg_nodes_c.data_ptr(), agg_arcs_c.data_ptr()
        args.max_iter, args.threshold = int(cfg.max_iter), float(cfg.threshold)
        args.training, args.save_for_backward = int(bool(cfg.training)), int(save)
        N.set_seed(args, cfg.seed)
        args.x_out, args.k_out = x_out.data_ptr(), k_out.data_ptr()

        if part is not None:
            args.n_global, args.row_offset = int(part.n_global), int(part.row_offset)
        nbytes = C.c_size_t(0)
        N.check(lib.gnn_state_loop_workspace_bytes(C.byref(graph), C.byref(mlp), C.byref(args), C.byref(nbytes)),
                'gnn_state_loop_workspace_bytes')
        if part is not None and hasattr(part, 'alloc_workspace'):
            workspace = part.alloc_workspace(nbytes.value)        # peer-mapped (symmetric) memory when the exchange is fused
        else:
            workspace = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
        callback, failure = None, []
        if part is not None and hasattr(part, 'peer_setup'):
            off, sbytes = C.c_size_t(0), C.c_size_t(0)
            N.check(lib.gnn_state_loop_layout(C.byref(graph), C.byref(mlp), C.byref(args), C.byref(off), C.byref(sbytes)),
                    'gnn_state_loop_layout')
            part.peer_setup(args, workspace, int(off.value))      # fills n_peers / rank / peer_state / peer_mask, orders the call
        if part is not None and getattr(part, 'needs_callback', True):
            DP = 4
            while DP < cfg.D: DP *= 2
            slab_bytes = int(part.n_global) * DP * 4

            def _exchange(_user, t, x_off, go_off):
                # called by the library between the enqueues of two iterations: enqueue the collectives on the same stream
                try:
                    x_full = workspace[x_off:x_off + slab_bytes].view(torch.float32).view(int(part.n_global), DP)
                    go = None if go_off < 0 else workspace[go_off:go_off + 4].view(torch.int32)
                    part.exchange(int(t), x_full, go)
                except BaseException as exc:   # never let an exception cross the C boundary
                    failure.append(exc)

            callback = N.EXCHANGE_FN(_exchange)
            args.exchange = C.cast(callback, C.c_void_p)
        with torch.cuda.device(device):
            N.check(lib.gnn_state_loop_forward(C.byref(graph), C.byref(mlp), C.byref(args), workspace.data_ptr(), nbytes.value,
                                               N._stream(device)), 'gnn_state_loop_forward')
        if failure: raise failure[0]
        if save:
            ctx.cfg, ctx.workspace, ctx.nbytes = cfg, workspace, nbytes.value
            # inputs only: keeping the Function's own outputs on ctx would close an output -> grad_fn -> ctx -> output cycle that
            # pins the workspace (12.9 GB at C4) until the cycle collector runs; backward reads k from the workspace
            ctx.held = (x0c, nodes_c, agg_nodes_c, agg_arcs_c, keep)
            ctx.need = (x0.requires_grad, nodes is not None and nodes.requires_grad, agg_nodes.requires_grad,
                        agg_arcs.requires_grad)
            ctx.n_params = len(params)
        ctx.mark_non_differentiable(k_out)
        return x_out, k_out

    @staticmethod
    def backward(ctx, g_x, _g_k):
        lib = N.lib()
        cfg = ctx.cfg
        x0c, nodes_c, agg_nodes_c, agg_arcs_c, _ = ctx.held
        device = x0c.device
        n_nodes = int(x0c.shape[0])
        g_x = g_x.contiguous().to(torch.float32)

        keep = []
        mlp = N.make_mlp(cfg.spec, keep)
        graph = N.make_graph(cfg.adj)
        args = N.gnn_loop_args()
        args.D, args.NL_self, args.NL_agg, args.AL = cfg.D, cfg.NL_self, cfg.NL_agg, cfg.AL
        args.x0, args.nodes = x0c.data_ptr(), None if nodes_c is None else nodes_c.data_ptr()
        args.agg_nodes, args.agg_arcs = agg_nodes_c.data_ptr(), agg_arcs_c.data_ptr()
        args.max_iter, args.threshold = int(cfg.max_iter), float(cfg.threshold)
        args.training, args.save_for_backward = int(bool(cfg.training)), 1
        N.set_seed(args, cfg.seed)
        args.x_out = args.k_out = None        # not touched by the backward sweep

        # gradients of the trainable variables, Keras order: [kernel, bias] per Dense (+ gamma, beta)
        grad = N.gnn_mlp_grad()
        param_grads = []
        for i, dense in enumerate(cfg.spec.dense_layers):
            dW, db = torch.empty_like(dense.kernel), torch.empty_like(dense.bias)
            grad.dW[i], grad.db[i] = dW.data_ptr(), db.data_ptr()
            param_grads += [dW, db]
        bn = cfg.spec.batchnorm
        if bn is not None:
            if cfg.training:
                dg, dbeta = torch.empty_like(bn.gamma), torch.empty_like(bn.beta)
                grad.dgamma, grad.dbeta = dg.data_ptr(), dbeta.data_ptr()
            else:
                # inference-mode normalisation inside a differentiated call: gamma/beta gradients are not produced
                dg, dbeta = torch.zeros_like(bn.gamma), torch.zeros_like(bn.beta)
            param_grads += [dg, dbeta]

        need_x0, need_nodes, need_an, need_aa = ctx.need
        new = lambda ref: torch.empty_like(ref)
        g_x0 = new(x0c) if need_x0 else None
        g_nodes = new(nodes_c) if need_nodes else None
        g_an = new(agg_nodes_c) if need_an and agg_nodes_c.shape[1] > 0 else None
        g_aa = new(agg_arcs_c) if need_aa and agg_arcs_c.shape[1] > 0 else None
        with torch.cuda.device(device):
            N.check(lib.gnn_state_loop_backward(C.byref(graph), C.byref(mlp), C.byref(args), g_x.data_ptr(), C.byref(grad),
                                                N._ptr(g_x0), N._ptr(g_nodes), N._ptr(g_an), N._ptr(g_aa),
                                                ctx.workspace.data_ptr(), ctx.nbytes, N._stream(device)), 'gnn_state_loop_backward')
        if need_an and g_an is None: g_an = torch.zeros_like(agg_nodes_c)
        if need_aa and g_aa is None: g_aa = torch.zeros_like(agg_arcs_c)
        ctx.workspace = None
        assert len(param_grads) == ctx.n_params
        return (None, g_x0, g_nodes, g_an, g_aa, *param_grads)


def state_loop(adjacency: SparseCSR, net_state: Sequential, x0: torch.Tensor, nodes: Optional[torch.Tensor],
               aggregated_nodes: torch.Tensor, aggregated_arcs: torch.Tensor, *, max_iteration: int, threshold: float,
               training: bool, seed: int = 0, partition=None):
    """ run the whole state-convergence loop on the device.

    :param partition: None, or an object with ``n_global``, ``row_offset`` and ``exchange(t, x_full, go_flag)`` for one
        node range of a graph split over several GPUs (``dist_graph.GraphPartition``): ``adjacency`` then holds the local
        rows with GLOBAL column ids, ``x0`` is the full (replicated) initial state, labels are local rows; forward only.

    :param adjacency: Adjacency^T as SparseCSR (rows = destination), with the transposed structure when gradients are needed
    :param net_state: keras_compat.Sequential; its parameters receive gradients through autograd
    :param x0: initial state [N, D]
    :param nodes: node labels [N, NL] concatenated to the own state when state_vect_dim > 0, else None
    :param aggregated_nodes: [N, NL] or [N, 0]; aggregated_arcs: [N, AL]
    :return: (k  -- device float32 scalar, number of iterations;  state [N, D])
    """
    if x0.device.type != 'cuda': raise RuntimeError('state_loop needs CUDA tensors: there is no CPU path')
    cfg = _LoopConfig()
    cfg.adj, cfg.net, cfg.spec = adjacency, net_state, net_state.spec()
    cfg.D = int(x0.shape[1])
    cfg.NL_self = 0 if nodes is None else int(nodes.shape[1])
    cfg.NL_agg, cfg.AL = int(aggregated_nodes.shape[1]), int(aggregated_arcs.shape[1])
    cfg.max_iter, cfg.threshold, cfg.training = int(max_iteration), float(threshold), bool(training)
    cfg.seed = seed if isinstance(seed, torch.Tensor) else int(seed)     # device tensor: seed read at kernel time (captured steps)
    cfg.partition = partition
    if cfg.spec.dims[0] != 2 * cfg.D + cfg.NL_self + cfg.NL_agg + cfg.AL:
        raise ValueError(f'net_state input width {cfg.spec.dims[0]} does not match the graph: expected '
                         f'{2 * cfg.D + cfg.NL_self + cfg.NL_agg + cfg.AL} = AL + 2*(NL + state_vect_dim)')
    if cfg.spec.dims[-1] != cfg.D:
        raise ValueError(f'net_state output width {cfg.spec.dims[-1]} != state width {cfg.D}')
    params = net_state.trainable_variables
    cfg.save = bool(torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                                    for t in [x0, nodes, aggregated_nodes, aggregated_arcs] + list(params)))
    if partition is not None and (cfg.save or training):
        raise NotImplementedError('a partitioned state loop is forward-only: call it under torch.no_grad() with training=False')
    x, k = _StateLoop.apply(cfg, x0.to(torch.float32), nodes, aggregated_nodes, aggregated_arcs, *params)
    return k, x
