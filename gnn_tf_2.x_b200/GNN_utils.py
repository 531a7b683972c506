# coding=utf-8
"""Host-side dataset helpers with the reference's names and semantics (``GNN/GNN_utils.py``): random / debug graphs,
index splitting, batching by merge, MinMax normalisation, leave-K-out fold builder.  NumPy only; nothing here is on
the GPU hot path.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np

from .graph_class import GraphObject, GraphTensor


# ---------------------------------------------------------------------------------------------------------------------
def _cluster_targets(features: np.ndarray, n_clusters: int) -> np.ndarray:
    """ 1-hot targets from agglomerative clustering of the labels (GNN_utils.py:66-70) """
    from sklearn.cluster import AgglomerativeClustering
    labels = AgglomerativeClustering(n_clusters=n_clusters).fit(features).labels_
    targets = np.zeros((features.shape[0], n_clusters))
    targets[np.arange(features.shape[0]), labels] = 1
    return targets


# ---------------------------------------------------------------------------------------------------------------------
def randomGraph(nodes_number: int, dim_node_label: int, dim_arc_label: int, dim_target: int, density: float,
                *, normalize_features: bool = False, aggregation_mode: str = 'average',
                problem_based: str = 'n') -> GraphObject:
    """ Random symmetric graph, label of arc (i,j) == label of (j,i) (GNN_utils.py:16-84).
    Consumes ``np.random`` in the same order as the reference, so a seeded run gives the same graph. """
    assert problem_based in ('n', 'a', 'g')
    nodes = 2 * np.random.random((nodes_number, dim_node_label)) - 1

    # arcs: pick half of the requested arcs as (low id -> higher id), drop duplicates, mirror them
    arcs_number = round(density * nodes_number * (nodes_number - 1) / 2)
    sources = np.random.choice(range(nodes_number)[:-1], arcs_number // 2)
    room = nodes_number - sources - 1
    destinations = sources + np.ceil(room * np.random.random(len(sources)))
    ascending = np.unique(np.stack([sources, destinations], axis=1).astype(float), axis=0)
    labels = 2 * np.random.random((ascending.shape[0], dim_arc_label)) - 1
    ids = np.concatenate([ascending, ascending[:, ::-1]])
    arcs = np.unique(np.concatenate([ids, np.concatenate([labels, labels])], axis=1), axis=0)

    if problem_based == 'g':
        targets = np.zeros((1, dim_target))
        targets[0, np.random.choice(range(dim_target))] = 1
    else:
        targets = _cluster_targets(arcs[:, 2:] if problem_based == 'a' else nodes, dim_target)

    output_mask = np.ones(arcs.shape[0] if problem_based == 'a' else nodes.shape[0], dtype=bool)
    if normalize_features:
        nodes = nodes / np.max(nodes, axis=0)
        arcs[:, 2:] = arcs[:, 2:] / np.max(arcs[:, 2:], axis=0)
    return GraphObject(arcs=arcs, nodes=nodes, targets=targets, problem_based=problem_based,
                       output_mask=output_mask, aggregation_mode=aggregation_mode)


# ---------------------------------------------------------------------------------------------------------------------
def simple_graph(problem_based: str, aggregation_mode: str = 'average') -> GraphObject:
    """ the fixed 4-node / 8-arc debug graph (GNN_utils.py:88-105) """
    nodes = np.array([[11, 21], [12, 22], [13, 23], [14, 24]])
    arcs = np.array([[0, 1, 10], [0, 2, 40], [1, 0, 10], [1, 2, 20], [2, 0, 40], [2, 1, 20], [2, 3, 30], [3, 2, 30]])
    if problem_based == 'g':
        targets = np.array([[0., 1.]])
    else:
        targets = _cluster_targets(arcs[:, 2:] if problem_based == 'a' else nodes, 2)
    return GraphObject(arcs=arcs, nodes=nodes, targets=targets, problem_based=problem_based, aggregation_mode=aggregation_mode)


# ---------------------------------------------------------------------------------------------------------------------
def progressbar(percent: float, width: int = 30) -> None:
    left = round(width * percent / 100)
    print('\r[', '#' * left, ' ' * int(width - left), ']', f' {percent:.1f}%', sep='', end='', flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def getindices(len_dataset: int, perc_Train: float = 0.7, perc_Valid: float = 0.1, seed=None):
    """ shuffled (train, test, validation) index lists -- note the order (GNN_utils.py:117-149).
    seed: number -> fixed shuffle; None -> random shuffle; False -> no shuffle """
    if perc_Train < 0 or perc_Valid < 0 or perc_Train + perc_Valid > 1:
        raise ValueError('Error - percentage must stay in [0-1] and their sum must be <= 1')
    idx = list(range(len_dataset))
    if seed: np.random.seed(seed)
    if seed is not False: np.random.shuffle(idx)
    n_test = round(len_dataset * (1 - perc_Train - perc_Valid))
    n_valid = round(len_dataset * perc_Valid)
    return idx[n_test + n_valid:], idx[:n_test], idx[n_test:n_test + n_valid]


# ---------------------------------------------------------------------------------------------------------------------
def getSet(glist: list[str], set_indices: list[int], problem_based: str, aggregation_mode: str,
           verbose: bool = False) -> list[GraphObject]:
    """ load the graphs whose folder paths are glist[set_indices] (GNN_utils.py:153-173) """
    if not (type(glist) == list and all(isinstance(x, str) for x in glist)):
        raise TypeError('type of param <glist> must be list of str \'path-like\' or GraphObjects')
    chosen = []
    for i, elem in enumerate(set_indices):
        chosen.append(glist[elem])
        if verbose: progressbar((i + 1) * 100 / len(set_indices))
    return [GraphObject.load(i, problem_based=problem_based, aggregation_mode=aggregation_mode) for i in chosen]


# ---------------------------------------------------------------------------------------------------------------------
def getbatches(glist: list[GraphObject], problem_based: str, aggregation_mode: str, batch_size: int = 32,
               number_of_batches=None, one_graph_per_batch=True):
    """ split a list of graphs into batches; each batch merged into one GraphObject by default (GNN_utils.py:177-194) """
    if number_of_batches is None:
        batches = [glist[i:i + batch_size] for i in range(0, len(glist), batch_size)]
    else:
        sizes = [len(part) for part in np.array_split(np.arange(len(glist)), number_of_batches)]
        starts = np.concatenate([[0], np.cumsum(sizes)])
        batches = [list(glist[a:b]) for a, b in zip(starts[:-1], starts[1:])]
    if one_graph_per_batch:
        batches = [GraphObject.merge(b, problem_based=problem_based, aggregation_mode=aggregation_mode) for b in batches]
    return batches


# ---------------------------------------------------------------------------------------------------------------------
def normalize_graphs(gTr, gVa, gTe, based_on: str = 'gTr', norm_rangeN=None, norm_rangeA=None) -> None:
    """ in-place MinMax normalisation of node labels and of ALL arc columns (ids included, as the reference does,
    GNN_utils.py:198-234); the graph structure is held separately by GraphObject, so it is not affected. """

    def as_list(g, name):
        if g is None: return []
        if not (type(g) == GraphObject or (type(g) == list and all(isinstance(x, GraphObject) for x in g))):
            raise TypeError(f'type of param <{name}> must be GraphObject or list of Graphobjects')
        return g if type(g) == list else [g]

    gTr, gVa, gTe = as_list(gTr, 'gTr'), as_list(gVa, 'gVa'), as_list(gTe, 'gTe')
    if based_on not in ['gTr', 'all']: raise ValueError('param <based_on> must be \'gTr\' or \'all\'')
    fit_on = gTr if based_on == 'gTr' else gTr + gVa + gTe

    from sklearn.preprocessing import MinMaxScaler
    node_scaler = MinMaxScaler(feature_range=(0, 1) if norm_rangeN is None else norm_rangeN)
    arcs_scaler = MinMaxScaler(feature_range=(0, 1) if norm_rangeA is None else norm_rangeA)
    merged = GraphObject.merge(fit_on, problem_based='n', aggregation_mode='sum')
    node_scaler.fit(merged.nodes)
    arcs_scaler.fit(merged.arcs)
    for g in gTr + gVa + gTe:
        g.nodes = node_scaler.transform(g.nodes)
        g.arcs = arcs_scaler.transform(g.arcs)


# ---------------------------------------------------------------------------------------------------------------------
def prepare_LKO_data(dataset, problem_based: str, number_of_batches: int = 10, useVa: bool = False,
                     seed: Optional[float] = None, normalize_method: str = 'gTr', aggregation_mode: str = 'average'):
    """ folds for ``model.LKO`` (GNN_utils.py:238-353): returns (gTRs, gTEs, gVAs).
    Single GraphObject: folds differ only by ``set_mask`` (here each set gets ITS OWN mask; the reference assigns the
    test mask to all three, GNN_utils.py:299,306 -- see DESIGN.md quirks). List (or list of per-class lists) of
    GraphObjects: graphs are shuffled, split into ``number_of_batches`` merged batches; fold i tests on batch i,
    validates on the last remaining batch if ``useVa``. """
    assert number_of_batches > 1 + useVa
    if seed: np.random.seed(seed)
    gTRs, gTEs, gVAs = [], [], []

    if isinstance(dataset, GraphObject):
        if normalize_method: normalize_graphs(dataset, None, None, based_on=normalize_method)
        base = GraphTensor.fromGraphObject(dataset)
        order = np.arange(len(dataset.set_mask))
        np.random.shuffle(order)
        chunks = np.array_split(order, number_of_batches)

        def with_mask(indices):
            import torch
            mask = np.zeros(len(dataset.set_mask), dtype=bool)
            mask[indices] = True
            g = base.copy()
            g.set_mask = torch.as_tensor(mask, device=base.device)
            return g

        for i in range(number_of_batches):
            rest = [c for j, c in enumerate(chunks) if j != i]
            gTEs.append(with_mask(chunks[i]))
            gVAs.append(with_mask(rest.pop(-1)) if useVa else None)
            gTRs.append(with_mask(np.concatenate(rest)))

    elif isinstance(dataset, list):
        if all(isinstance(i, GraphObject) for i in dataset): dataset = [dataset]
        assert all(isinstance(i, list) for i in dataset) and all(isinstance(j, GraphObject) for i in dataset for j in i)
        assert all(len(i) > number_of_batches for i in dataset)
        for group in dataset: np.random.shuffle(group)
        per_class = [getbatches(group, problem_based, aggregation_mode, -1, number_of_batches, False) for group in dataset]
        folds = [[g for cls in per_class for g in cls[j]] for j in range(number_of_batches)]
        for fold in folds: np.random.shuffle(fold)
        merged = [GraphObject.merge(fold, problem_based=problem_based, aggregation_mode=aggregation_mode) for fold in folds]
        for i in range(number_of_batches):
            # normalisation is in place: every fold works on its own copies (the reference re-normalises shared objects)
            gTr = [g.copy_with_problem(problem_based) for g in merged]
            gTe = gTr.pop(i)
            gVa = gTr.pop(-1) if useVa else None
            if normalize_method: normalize_graphs(gTr, gVa, gTe, based_on=normalize_method)
            gTRs.append([GraphTensor.fromGraphObject(g) for g in gTr])
            gTEs.append(GraphTensor.fromGraphObject(gTe))
            gVAs.append(GraphTensor.fromGraphObject(gVa) if gVa is not None else None)
    else:
        raise TypeError('param <dataset> must be a GraphObject, a list of GraphObjects or a list of lists of Graphobjects')
    return gTRs, gTEs, gVAs
