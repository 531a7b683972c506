# coding=utf-8
"""ORACLE (test infrastructure, NOT product code) -- integer/index part of the hot path.

CPU restatement in NumPy of the reference's arc preprocessing.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the product package never does.

Parity status: the GraphObject-level functions below (arcnode_coo, adjacency_coo, nodegraph, merge) are PINNED against
the reference's own ``GraphObject`` executed in the build container (``oracle/gen_golden.py`` ->
``tests/golden/graphobject_*.npz``).  ``transposed_row_major`` restates ``tf.sparse.reorder`` (TensorFlow is not
installable here): PARITY UNPINNED at that boundary, documented behaviour = sort entries by (row, col).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
def arcnode_coo(dst: np.ndarray, n_nodes: int, aggregation_mode: str):
    """ GNN/graph_class.py:98-121 -- ArcNode (E, N): row = arc id, col = destination node, value by aggregation mode.
    :return: (row int64[E], col int64[E], data float32[E]) """
    dst = np.asarray(dst, dtype=np.int64)
    n_arcs = len(dst)
    values = np.ones(n_arcs, dtype=np.float64)                                     # :106
    if aggregation_mode == 'normalized':
        values = values * float(1 / n_arcs)                                         # :110-113 (1/len(col) = 1/#arcs)
    elif aggregation_mode == 'average':
        _, inverse, counts = np.unique(dst, return_inverse=True, return_counts=True)  # :116-118
        values = values / counts[inverse]
    elif aggregation_mode != 'sum':
        raise ValueError('ERROR: Unknown aggregation mode')
    return np.arange(n_arcs, dtype=np.int64), dst, values.astype(np.float32)        # :120 (dtype float32)


def adjacency_coo(src: np.ndarray, dst: np.ndarray, arcnode_data: np.ndarray):
    """ GNN/graph_class.py:90-95 -- Adjacency (N, N): entry (src, dst) = ArcNode value of the same arc, arc order kept,
    duplicates kept. :return: (row=src, col=dst, data) """
    return np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64), np.asarray(arcnode_data, dtype=np.float32)


def transposed_row_major(row: np.ndarray, col: np.ndarray, data: np.ndarray, shape: tuple[int, int]):
    """ GNN/graph_class.py:364-372 -- COO2SparseTransposedTensor: indices = zip(col, row), dense_shape swapped, then
    tf.sparse.reorder = canonical row-major order, i.e. sorted by (new_row, new_col) = (col, row); ties (duplicate
    arcs) in original order.
    :return: dict(indices int64[nnz, 2], values float32[nnz], dense_shape, perm int64[nnz], rowptr int64[rows+1]) """
    new_row, new_col = np.asarray(col, dtype=np.int64), np.asarray(row, dtype=np.int64)
    perm = np.lexsort((new_col, new_row))            # last key is primary; lexsort is stable
    n_rows = shape[1]
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, new_row + 1, 1)
    rowptr = np.cumsum(rowptr)
    return dict(indices=np.stack([new_row[perm], new_col[perm]], axis=1), values=np.asarray(data, dtype=np.float32)[perm],
                dense_shape=(shape[1], shape[0]), perm=perm, rowptr=rowptr)


def csr_transpose(rowptr: np.ndarray, col: np.ndarray, n_cols: int):
    """ source-sorted CSR^T of a destination-sorted CSR (our own structure for the backward pass, no reference line):
    entries ordered by (col, position in CSR). :return: (rowptr_T, col_T = destination row, perm_T = CSR position) """
    nnz = len(col)
    rows = np.repeat(np.arange(len(rowptr) - 1, dtype=np.int64), np.diff(rowptr))
    perm_T = np.lexsort((np.arange(nnz), np.asarray(col, dtype=np.int64)))
    rowptr_T = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(rowptr_T, np.asarray(col, dtype=np.int64) + 1, 1)
    return np.cumsum(rowptr_T), rows[perm_T], perm_T


def nodegraph(n_nodes: int, problem_based: str):
    """ GNN/graph_class.py:132-144 -- (N, 1) matrix of 1/N for graph-based problems, None otherwise """
    if problem_based != 'g': return None
    return np.ones((n_nodes, 1), dtype=np.float32) * 1 / n_nodes


def merge(graphs: list[dict], problem_based: str):
    """ GNN/graph_class.py:284-319 -- disjoint union. Each graph is a dict(arcs, nodes, targets, set_mask, output_mask,
    sample_weights, NodeGraph). Arc ids (float32 columns 0-1) are shifted by the cumulated node counts (:304); NodeGraph
    becomes block-diagonal (:313-315). """
    offsets = np.cumsum([0] + [g['nodes'].shape[0] for g in graphs])
    arcs = []
    for g, off in zip(graphs, offsets):
        block = g['arcs'].astype(np.float32).copy()
        block[:, :2] += off
        arcs.append(block)
    out = dict(arcs=np.concatenate(arcs, axis=0))
    for key in ('nodes', 'targets', 'set_mask', 'output_mask', 'sample_weights'):
        out[key] = np.concatenate([g[key] for g in graphs], axis=0)
    out['NodeGraph'] = None
    if problem_based == 'g':
        sizes = [g['NodeGraph'].shape for g in graphs]
        dense = np.zeros((sum(s[0] for s in sizes), sum(s[1] for s in sizes)), dtype=np.float32)
        r = c = 0
        for g, (nr, nc) in zip(graphs, sizes):
            dense[r:r + nr, c:c + nc] = g['NodeGraph']
            r, c = r + nr, c + nc
        out['NodeGraph'] = dense
    return out


# ---------------------------------------------------------------------------------------------------------------------
def spmm_rows(rowptr, col, val, dense):
    """ ``tf.sparse.sparse_dense_matmul`` on CPU as documented in SURVEY 8c: sequential over the stored entries, float32
    accumulation, i.e. each output row sums its entries in ascending stored order (GNN/GNN.py:234,259,263). """
    rowptr, col = np.asarray(rowptr), np.asarray(col)
    dense = np.asarray(dense, dtype=np.float32)
    out = np.zeros((len(rowptr) - 1, dense.shape[1]), dtype=np.float32)
    for r in range(len(rowptr) - 1):
        acc = np.zeros(dense.shape[1], dtype=np.float32)
        for e in range(rowptr[r], rowptr[r + 1]):
            acc = acc + np.float32(val[e]) * dense[col[e]]
        out[r] = acc
    return out
