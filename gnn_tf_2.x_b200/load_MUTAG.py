# coding=utf-8
"""MUTAG ("Mutagenicity", TU format) loader with the semantics of the reference's ``load_MUTAG.py`` restated for NumPy 2
(the reference's ``delimiter=', '`` crashes on NumPy >= 1.23).  Host-side fixture code for config 2, not on the hot path.

Mirrored behaviour (load_MUTAG.py:14-52): node / edge / graph labels one-hot; the edge list is re-sorted with
``np.unique(edges, axis=0)`` (``:28``) while the edge-label file stays in raw order (so labels are paired with the sorted
list, as the reference does); node ids of a graph are renumbered to 0..n-1 over the ids that appear in its edges
(``:33-36``); one GraphObject per graph, problem_based='g'.
"""
from __future__ import annotations

import os

import numpy as np

from .graph_class import GraphObject


def load_MUTAG(path: str = 'MUTAG_raw/', aggregation_mode: str = 'average') -> list[GraphObject]:
    if path[-1] != '/': path += '/'
    edges = np.loadtxt(path + 'Mutagenicity_edges.txt', dtype=int, delimiter=',')
    edge_labels = np.loadtxt(path + 'Mutagenicity_edge_labels.txt', dtype=int)
    node_labels = np.loadtxt(path + 'Mutagenicity_node_labels.txt', dtype=int)
    graph_of_node = np.loadtxt(path + 'Mutagenicity_graph_indicator.txt', dtype=int)
    graph_targets = np.loadtxt(path + 'Mutagenicity_graph_labels.txt', dtype=int)

    _, first = np.unique(graph_of_node, return_index=True)
    bounds = np.concatenate([first, [len(graph_of_node)]])               # node i (1-based) belongs to graph g iff bounds[g] < i <= bounds[g+1]
    onehot = lambda labels: np.eye(len(np.unique(labels)), dtype=int)[labels]
    nL, eL, targets = onehot(node_labels), onehot(edge_labels), onehot(graph_targets)

    edges = np.unique(edges, axis=0)                                     # sorted edge list (labels stay in file order)
    owner = np.searchsorted(bounds, edges[:, 0], side='left') - 1        # graph of the first endpoint (ids are 1-based)
    same = owner == np.searchsorted(bounds, edges[:, 1], side='left') - 1
    graphs = []
    for gidx in range(len(bounds) - 1):
        sel = np.nonzero((owner == gidx) & same)[0]
        ids = edges[sel]
        uniq, renumbered = np.unique(ids, return_inverse=True)           # compress the ids that appear in edges
        arcs = np.concatenate([renumbered.reshape(ids.shape), eL[sel]], axis=1)
        nodes = nL[bounds[gidx]:bounds[gidx + 1]]
        graphs.append(GraphObject(arcs=arcs, nodes=nodes, targets=targets[gidx][np.newaxis, ...], problem_based='g',
                                  aggregation_mode=aggregation_mode))
    return graphs
