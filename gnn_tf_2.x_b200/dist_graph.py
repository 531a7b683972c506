# coding=utf-8
"""Multi-GPU execution (the reference itself is single-process).  One process per GPU,
``torch.distributed`` (NCCL over NVLink/NVSwitch) for the plumbing.

* **Batches of graphs** shard by whole graph: every rank runs the full loop on its own merged batch and the flat MLP
  gradient is summed with one all-reduce per step (``allreduce_gradients``, used by ``BaseClass.training_step`` once
  ``model.distributed = True``).  No data-path collective; iteration count and BatchNormalization statistics are those of
  the rank's own batch (= the reference run on that shard).
* **One giant graph** is split by contiguous node ranges (``GraphPartition``): rank r owns the destination rows
  [lo_r, hi_r) of the CSR (global column ids) and the matching rows of the state; every rank keeps a full-size state
  buffer.  After each iteration the ranks exchange the rows other ranks gather from (``HaloPlan``: all-gather when almost
  everything is needed, all-to-all of the packed boundary rows otherwise) and max-reduce the convergence flag, all
  enqueued on the compute stream by the library's ``exchange`` callback -- still no host synchronisation in the loop.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------------------------
# graph batches: gradient all-reduce
# ---------------------------------------------------------------------------------------------------------------------
def allreduce_gradients(grads: list[torch.Tensor], group=None) -> list[torch.Tensor]:
    """ SUM the gradients of all ranks with ONE collective over a flat buffer (the loss is a sum over targets,
    GNN.py:199, so the gradient of the global batch is the sum of the per-shard gradients) """
    if not dist.is_initialized() or dist.get_world_size(group) == 1: return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, pos = [], 0
    for g in grads:
        out.append(flat[pos:pos + g.numel()].view_as(g))
        pos += g.numel()
    return out


def shard_graphs(graphs: list, rank: int, world: int) -> list:
    """ whole graphs per rank, contiguous blocks of (almost) equal length """
    bounds = np.linspace(0, len(graphs), world + 1).astype(int)
    return graphs[bounds[rank]:bounds[rank + 1]]


# ---------------------------------------------------------------------------------------------------------------------
# one giant graph: node-range partition
# ---------------------------------------------------------------------------------------------------------------------
def partition_bounds(n_nodes: int, world: int) -> list[int]:
    """ contiguous node ranges of equal length (the last one takes the remainder) """
    step = -(-n_nodes // world)
    return [min(r * step, n_nodes) for r in range(world)] + [n_nodes]


class HaloPlan:
    """ which state rows a rank must receive from / send to every other rank after each iteration.
    Device-agnostic (CUDA + NCCL in production, CPU + gloo in the tests). """

    def __init__(self, cols: torch.Tensor, bounds: list[int], rank: int, world: int, group=None, allgather_threshold: float = 0.5):
        self.rank, self.world, self.group, self.bounds = rank, world, group, bounds
        lo, hi = bounds[rank], bounds[rank + 1]
        device = cols.device
        self.lo, self.hi = lo, hi
        # remote rows this rank gathers from: a bitmap over the nodes (one scatter over the arcs, no sort)
        wanted = torch.zeros(bounds[-1], dtype=torch.bool, device=device)
        if cols.numel(): wanted[cols.to(torch.int64)] = True
        wanted[lo:hi] = False
        # all-gather instead of all-to-all when most remote rows are needed anyway (uniform-random graphs): decided first, because
        # then nobody needs the row lists and their two all-to-all exchanges
        sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
        remote = bounds[-1] - (hi - lo)
        frac = (wanted.sum().to(torch.float32) / max(remote, 1)).reshape(1)
        dist.all_reduce(frac, op=dist.ReduceOp.MAX, group=group)
        self.use_allgather = bool(frac.item() > allgather_threshold) and len(set(sizes)) == 1
        if self.use_allgather:
            self.recv_rows = self.send_rows = None
            self.recv_counts = self.send_counts = None
            return
        needed = torch.nonzero(wanted, as_tuple=False)[:, 0]                         # sorted
        edges = torch.as_tensor(bounds[1:], dtype=torch.int64, device=device)
        owner = torch.bucketize(needed, edges, right=True)
        recv_counts = torch.bincount(owner, minlength=world)[:world]
        send_counts = torch.empty_like(recv_counts)
        dist.all_to_all_single(send_counts, recv_counts, group=group)
        self.recv_counts, self.send_counts = recv_counts.tolist(), send_counts.tolist()
        send_rows = torch.empty(int(sum(self.send_counts)), dtype=torch.int64, device=device)
        dist.all_to_all_single(send_rows, needed, output_split_sizes=self.send_counts, input_split_sizes=self.recv_counts, group=group)
        self.recv_rows, self.send_rows = needed, send_rows                          # global ids, grouped by peer rank

    def exchange(self, x_full: torch.Tensor) -> None:
        """ make the rows of x_full that this rank gathers from valid (its own rows [lo, hi) were just computed) """
        if self.world == 1: return
        if self.use_allgather:
            dist.all_gather_into_tensor(x_full, x_full[self.lo:self.hi], group=self.group)
            return
        send = x_full.index_select(0, self.send_rows)
        recv = torch.empty((self.recv_rows.numel(), x_full.shape[1]), dtype=x_full.dtype, device=x_full.device)
        dist.all_to_all_single(recv, send, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts, group=self.group)
        x_full.index_copy_(0, self.recv_rows, recv)

    def bytes_received_per_exchange(self, row_bytes: int) -> int:
        if self.use_allgather: return (self.bounds[-1] - (self.hi - self.lo)) * row_bytes
        return int(self.recv_rows.numel()) * row_bytes


_SYMMETRIC_WORKSPACES: dict = dict()
_SIGNALS: dict = dict()          # (device, group) -> [symmetric int32 signal area, handle, capacity in iterations, epoch of the last call]


class GraphPartition:
    """ rows [lo, hi) of one GraphObject on this rank: local CSRs with global column ids + replicated labels.
    Passed as ``partition=`` to ``state_loop`` (attributes n_global, row_offset, exchange).

    Exchange of the boundary state rows after every iteration, two implementations:
      * fused (default when peer memory is available): the loop's workspace lives in symmetric memory
        (torch.distributed._symmetric_memory), the iteration kernel stores each new row straight into the state buffers of
        the peers that gather from it (NVLink P2P stores overlapped with the MLP of the next tiles) and the callback only
        max-reduces the convergence flag -- which is also the barrier that orders the iterations across ranks;
      * NCCL: all-gather / all-to-all of the rows after the kernel (``fused=False`` or when symmetric memory is refused). """

    def __init__(self, g, rank: int, world: int, device=None, group=None, fused: bool = True):
        from . import _native
        from .graph_class import GraphObject
        assert isinstance(g, GraphObject)
        self.rank, self.world, self.group = rank, world, group
        self.device = _native.default_device() if device is None else torch.device(device)
        self.n_global = int(g.nodes.shape[0])
        self.bounds = partition_bounds(self.n_global, world)
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        self.row_offset, self.n_local = lo, hi - lo
        adj, an = g.Adjacency, g.ArcNode                          # COO: Adjacency (src, dst), ArcNode (arc, dst)
        dev = self.device
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)
        labels = g._arc_labels if getattr(g, '_arc_labels_of', None) is g.arcs else g.arcs[:, 2:]     # page-locked copy when pinned
        # Arcs sorted by destination (what merge / the generators produce): the arcs entering my node range are ONE slice of
        # every array, selected on the host for free -- only those rows cross PCIe.  Otherwise the whole COO goes to the device
        # once and the selection runs there.
        sorted_dst = getattr(g, '_dst_sorted', None)
        if sorted_dst is None:
            sorted_dst = g._dst_sorted = bool(adj.col.shape[0] < 2 or np.all(adj.col[1:] >= adj.col[:-1]))
        if sorted_dst:
            # (the bounds in the array's own dtype: a Python int makes numpy convert all E destinations to int64 first -- 13 ms per call at C4)
            key = adj.col.dtype.type
            a0, a1 = int(np.searchsorted(adj.col, key(lo), side='left')), int(np.searchsorted(adj.col, key(min(hi, np.iinfo(adj.col.dtype).max)), side='left'))
            n_mine = a1 - a0
            dst_loc = (up(adj.col[a0:a1], torch.int32) - lo).to(torch.int32)
            src_mine = up(adj.row[a0:a1], torch.int32)
            vals, an_vals = up(adj.data[a0:a1], torch.float32), up(an.data[a0:a1], torch.float32)
            self.arc_labels = up(labels[a0:a1], torch.float32)
            self.h2d_bytes = 4 * n_mine * (4 + int(labels.shape[1]))
        else:
            dst_all, src_all = up(adj.col, torch.int32), up(adj.row, torch.int32)
            mine = torch.nonzero((dst_all >= lo) & (dst_all < hi), as_tuple=False)[:, 0]      # arc order kept
            n_mine = int(mine.numel())
            dst_loc = (dst_all.index_select(0, mine) - lo).to(torch.int32)
            src_mine = src_all.index_select(0, mine)
            vals, an_vals = up(adj.data, torch.float32).index_select(0, mine), up(an.data, torch.float32).index_select(0, mine)
            self.arc_labels = up(labels, torch.float32).index_select(0, mine)
            self.h2d_bytes = 4 * int(adj.col.shape[0]) * (4 + int(labels.shape[1]))
        # Adjacency^T rows = local destination, columns = GLOBAL source; ArcNode^T rows = local destination, columns =
        # position of the arc among this rank's arcs (its labels are kept in that order)
        self.Adjacency = _native.csr_build(dst_loc, src_mine, vals, self.n_local, self.n_global)
        self.ArcNode = _native.csr_build(dst_loc, torch.arange(n_mine, dtype=torch.int32, device=dev), an_vals, self.n_local, max(n_mine, 1))
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
        self.nodes = f32(g.nodes)                                 # replicated (label widths are small)
        self.h2d_bytes += int(g.nodes.size) * 4
        self.n_arcs_local = n_mine
        self.halo = HaloPlan(self.Adjacency.col, self.bounds, rank, world, group) if world > 1 else None
        import os
        env = os.environ.get('GNN_B200_FUSED')                    # '0' / '1': force the NCCL / the fused exchange (experiments)
        if env is not None: fused = env != '0'
        # round 1 fell back to an NCCL all-gather at 8 ranks when every row travels (14.5 ms per 50 iterations against 16.1 ms for 16-byte
        # peer stores); with 32-byte row pieces and no per-iteration host collective the fused path is used at every N
        self.fused = bool(fused and world > 1 and world <= 8 and self.device.type == 'cuda')
        self.in_kernel_signals = self.fused and os.environ.get('GNN_B200_SIGNALS', '1') != '0'     # '0': per-iteration NCCL all-reduce of the flag
        self._ws = self._handle = self._peer_mask = self._state_offsets = None
        self._order = torch.zeros(1, dtype=torch.int32, device=self.device) if world > 1 else None
        if self.fused: self._build_peer_mask()

    # ---- fused exchange over peer memory --------------------------------------------------------------------------
    def _build_peer_mask(self):
        """ bit r of peer_mask[local row] = rank r gathers from that row; None when every peer needs (almost) every row """
        if self.halo.use_allgather:
            self._peer_mask = None
            return
        mask = torch.zeros(self.n_local, dtype=torch.int64, device=self.device)
        pos = 0
        for r, count in enumerate(self.halo.send_counts):
            if count: mask.index_add_(0, self.halo.send_rows[pos:pos + count] - self.row_offset, torch.full((count,), 1 << r, dtype=torch.int64, device=self.device))
            pos += count
        self._peer_mask = mask.to(torch.int32)     # rows are unique per peer, so the sum is the OR of the bits

    def alloc_workspace(self, nbytes: int) -> torch.Tensor:
        if not self.fused: return torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        # one symmetric workspace per (device, group), shared by every partition object of the process: the collective
        # allocation + rendezvous costs milliseconds and the loop calls are ordered on the stream anyway
        key = (str(self.device), id(self.group))
        cached = _SYMMETRIC_WORKSPACES.get(key)
        # a repeated call of THIS partition object with a size it has already agreed on with the peers: every rank takes this branch
        # together (same program, same objects), no collective and no host synchronisation
        if cached is not None and self._ws is cached[0] and nbytes <= getattr(self, '_ws_agreed', -1): return self._ws
        need = torch.tensor([nbytes], dtype=torch.int64, device=self.device)
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)              # same decision on every rank
        if cached is None or cached[0].numel() < int(need.item()):
            try:
                import torch.distributed._symmetric_memory as symm
                ws = symm.empty(int(need.item()), dtype=torch.uint8, device=self.device)
                handle = symm.rendezvous(ws, group=self.group if self.group is not None else dist.group.WORLD)
                cached = _SYMMETRIC_WORKSPACES[key] = (ws, handle)
            except Exception as exc:                                                 # no peer mapping on this system: NCCL exchange
                self.fused, self._fused_error = False, repr(exc)
                return torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if self._ws is not cached[0]: self._state_offsets = None
        self._ws, self._handle = cached
        self._ws_agreed = nbytes
        return self._ws

    def _signals(self, iters: int):
        """ the group's signal area for the in-kernel cross-GPU signalling (gnn_loop_args.sig_*): symmetric memory, zero-filled ONCE
        (marks are epochs that only grow), re-allocated when a loop needs more iterations than it holds """
        import torch.distributed._symmetric_memory as symm
        key = (str(self.device), id(self.group))
        entry = _SIGNALS.get(key)
        if entry is None or entry[2] < iters:
            cap = max(64, 2 * iters)
            sig = symm.empty(2 * 8 * cap, dtype=torch.int32, device=self.device)
            sig.zero_()
            handle = symm.rendezvous(sig, group=self.group if self.group is not None else dist.group.WORLD)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)              # every rank's area is zero before anybody signals
            entry = _SIGNALS[key] = [sig, handle, cap, 0 if entry is None else entry[3]]
        return entry

    needs_callback = property(lambda self: self.world > 1 and not (self.fused and self.in_kernel_signals))

    def peer_setup(self, args, workspace, state_offset: int) -> None:
        if self.world == 1: return
        # order this call after everything the peers did with the shared buffers in the previous call
        dist.all_reduce(self._order, op=dist.ReduceOp.MAX, group=self.group)
        if not self.fused: return
        if self.in_kernel_signals:
            try:
                entry = self._signals(int(args.max_iter) + 1)
            except Exception as exc:                        # no symmetric memory for the signals: host-side flag reduction
                self.in_kernel_signals, self._signal_error = False, repr(exc)
            else:
                entry[3] += 1                               # one epoch per partitioned call, the same on every rank
                args.sig_local = int(entry[1].buffer_ptrs[self.rank])
                for r in range(self.world): args.sig_peer[r] = int(entry[1].buffer_ptrs[r])
                args.sig_epoch = entry[3] & 0xFFFFFFFF
        # The offset of the state buffers inside the workspace depends on max_iteration, the size of the packed net and the
        # width of the constant rows: another GNN / LGNN layer over the same partition moves it.  The cache is keyed on the
        # local layout; every rank runs the same program (same nets, same max_iteration), so all ranks re-gather together.
        key = (id(self._ws), int(state_offset))
        if self._state_offsets is None or self._state_offsets[0] != key:
            mine = torch.tensor([state_offset], dtype=torch.int64, device=self.device)
            every = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=self.group)
            self._state_offsets = (key, [int(t.item()) for t in every])
        offsets = self._state_offsets[1]
        args.n_peers, args.rank = self.world, self.rank
        for r in range(self.world): args.peer_state[r] = int(self._handle.buffer_ptrs[r]) + offsets[r]
        args.peer_mask = None if self._peer_mask is None else self._peer_mask.data_ptr()

    def exchange(self, t: int, x_full: torch.Tensor, go_flag: Optional[torch.Tensor]) -> None:
        if self.world == 1: return
        if not self.fused: self.halo.exchange(x_full)      # fused: the kernel has already stored the rows into the peers
        # max-reduce of the flag = every rank runs the same iterations; it is also the cross-GPU barrier between iterations
        dist.all_reduce(go_flag if go_flag is not None else self._order, op=dist.ReduceOp.MAX, group=self.group)


def partitioned_loop(gnn, part: GraphPartition, x0: Optional[torch.Tensor] = None, seed: int = 0):
    """ GNNnodeBased.Loop (inference) on one node range; returns (k, state [n_global, D], outputs of the LOCAL rows).
    In the returned state the rows of this rank and the rows it gathers from are valid (all rows when every row travels).
    Same arithmetic as the single-GPU call: the result does not depend on the number of ranks. """
    from .state_loop import state_loop, sparse_dense
    gnn.to(part.device)
    lo, hi = part.row_offset, part.row_offset + part.n_local
    with torch.no_grad():
        aggregated_arcs = sparse_dense(part.ArcNode, part.arc_labels)
        if gnn.state_vect_dim > 0:
            state0 = x0 if x0 is not None else gnn.initial_state
            if state0 is None: raise ValueError('a partitioned loop needs a replicated initial state (x0 or gnn.initial_state)')
            state0 = state0.to(part.device, torch.float32)
            aggregated_nodes = sparse_dense(part.Adjacency, part.nodes)
            node_self = part.nodes[lo:hi].contiguous()
        else:
            state0 = part.nodes
            aggregated_nodes = torch.zeros((part.n_local, 0), dtype=torch.float32, device=part.device)
            node_self = None
        k, state = state_loop(part.Adjacency, gnn.net_state, state0, node_self, aggregated_nodes, aggregated_arcs,
                              max_iteration=gnn.max_iteration, threshold=gnn.state_threshold, training=False, seed=seed, partition=part)
        local = state[lo:hi]
        net_in = torch.cat([local, part.nodes[lo:hi]], dim=1) if gnn.state_vect_dim else local
        out = gnn.net_output(net_in, training=False)
    return k, state, out


# ---------------------------------------------------------------------------------------------------------------------
# bench.py leg for N > 1 (strong scaling of the C4 graph)
# ---------------------------------------------------------------------------------------------------------------------
def bench_partitioned(g_host, wl, build_gnn, args, device, rank, world, with_e2e: bool = True):
    from . import _native
    gnn = build_gnn()
    part = GraphPartition(g_host, rank, world, device=device)
    E = wl['E']
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup):
        for _ in range(warmup): fn()
        torch.cuda.synchronize(); dist.barrier()
        start.record()
        for _ in range(steps): fn()
        stop.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([start.elapsed_time(stop) / steps], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)       # the job is as slow as its slowest rank
        return float(ms.item())

    ks = []

    def fwd():
        k, state, out = partitioned_loop(gnn, part)
        ks.append(k)

    _native.launch_count(reset=True)
    warm = max(3, args.warmup)
    ms_fwd = timed(fwd, args.steps, warm)
    launches = _native.launch_count() // (args.steps + warm)          # library kernels per step (one partitioned loop), as at N = 1
    k_fwd = float(ks[-1])

    halo = part.halo.bytes_received_per_exchange(128) if part.halo is not None else 0
    partition = {'ranks': world, 'rows_per_rank': part.n_local,
                 'exchange': ('fused NVLink peer stores in the iteration kernel' if part.fused else 'NCCL ') +
                             (' (every row to every peer)' if part.halo.use_allgather else ' (boundary rows only)'),
                 'halo_bytes_received_per_iteration_per_rank': int(halo)}
    if not with_e2e:
        return {'value': E * k_fwd / (ms_fwd * 1e-3), 'ms_per_step': ms_fwd, 'iterations': k_fwd, 'gpu_launches': int(launches), 'partition': partition}

    # e2e: host buffers of the local rows -> device (CSR build included) -> loop -> local outputs back on the host
    g_host.pin_host_buffers()
    x0_host = torch.from_numpy(wl['x0']).pin_memory()
    held = []

    def e2e_step():
        gnn.initial_state = x0_host.to(device, non_blocking=True)
        p = GraphPartition(g_host, rank, world, device=device)
        k, state, out = partitioned_loop(gnn, p)
        held.append(out.cpu())
        if len(held) > 2: held.pop(0)

    ms_e2e = timed(e2e_step, args.steps, 2)   # two warm-ups: the first partitions pay the symmetric-memory rendezvous
    local_bytes = part.h2d_bytes + x0_host.numel() * 4           # this rank's arcs + the replicated node labels and initial state
    return {'value': E * k_fwd / (ms_fwd * 1e-3), 'ms_per_step': ms_fwd, 'iterations': k_fwd, 'gpu_launches': int(launches),
            'e2e': {'value': E * k_fwd / (ms_e2e * 1e-3), 'unit': 'arc-updates/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': int(local_bytes), 'd2h_bytes_per_step': int(held[-1].numel() * 4)},
            'partition': partition}
