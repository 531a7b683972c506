# coding=utf-8
"""Generate tests/golden/*.npz by running the REFERENCE's own NumPy/SciPy code (GraphObject, GNN_utils helpers,
MLP.get_inout_dims) from /root/reference in the build container.

TensorFlow is not installable here; the reference modules used below only need ``tf.keras.backend.floatx()``,
``tf.Tensor`` as an annotation and importable ``tensorflow.keras.layers/models`` names, so a stub package is put on
sys.path (nothing TF-numerical is executed -- everything that would need real TF is left "parity unpinned").

Run:  python oracle/gen_golden.py        (needs /root/reference; the fixtures are committed, tests never need it)
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def install_tf_stub():
    tf = types.ModuleType('tensorflow')
    keras = types.ModuleType('tensorflow.keras')
    layers = types.ModuleType('tensorflow.keras.layers')
    models = types.ModuleType('tensorflow.keras.models')

    class _Backend:
        floatx = staticmethod(lambda: 'float32')

    class _Missing:
        def __init__(self, *a, **k): raise RuntimeError('TensorFlow stub: numerical TF objects are not available')

    for name in ('Dense', 'Dropout', 'AlphaDropout', 'BatchNormalization'): setattr(layers, name, _Missing)
    models.Sequential = _Missing
    keras.backend, keras.layers, keras.models = _Backend, layers, models
    tf.keras, tf.Tensor = keras, type('Tensor', (), {})
    tf.constant = lambda v, dtype=None: v
    sys.modules.update({'tensorflow': tf, 'tensorflow.keras': keras, 'tensorflow.keras.layers': layers,
                        'tensorflow.keras.models': models})


def graph_fields(g):
    d = dict(arcs=g.arcs, nodes=g.nodes, targets=g.targets, set_mask=g.set_mask, output_mask=g.output_mask,
             sample_weights=g.sample_weights, arcnode_row=g.ArcNode.row, arcnode_col=g.ArcNode.col,
             arcnode_data=g.ArcNode.data, arcnode_shape=np.array(g.ArcNode.shape), adj_row=g.Adjacency.row,
             adj_col=g.Adjacency.col, adj_data=g.Adjacency.data, adj_shape=np.array(g.Adjacency.shape),
             dims=np.array([g.DIM_NODE_LABEL, g.DIM_ARC_LABEL, g.DIM_TARGET]))
    if g.NodeGraph is not None: d['NodeGraph'] = g.NodeGraph
    # reference products with scipy (no TF involved): Adjacency^T @ nodes, ArcNode^T @ arc labels
    d['adjT_nodes'] = (g.Adjacency.T.tocsr() @ g.nodes).astype(np.float32)
    d['arcnodeT_labels'] = (g.ArcNode.T.tocsr() @ g.arcs[:, 2:]).astype(np.float32)
    return d


def main():
    install_tf_stub()
    sys.path.insert(0, REF)
    from GNN.graph_class import GraphObject
    from GNN import GNN_utils as utils
    from GNN.MLP import get_inout_dims
    os.makedirs(OUT, exist_ok=True)

    # 1. simple_graph: every problem type x aggregation mode (GNN_utils.py:88-105)
    for pb in ('n', 'a', 'g'):
        for mode in ('average', 'normalized', 'sum'):
            g = utils.simple_graph(pb, mode)
            np.savez(f'{OUT}/graphobject_simple_{pb}_{mode}.npz', **graph_fields(g))

    # 2. randomGraph with the legacy NumPy seed (GNN_utils.py:16-84) + merge of 5 graphs (graph_class.py:284-319)
    for pb in ('n', 'g'):
        np.random.seed(7)
        glist = [utils.randomGraph(int(n), 3, 1, 2, 0.7, aggregation_mode='average', problem_based=pb) for n in (15, 22, 17, 30, 9)]
        np.savez(f'{OUT}/graphobject_random_{pb}.npz', **graph_fields(glist[1]))
        merged = GraphObject.merge(glist, problem_based=pb, aggregation_mode='average')
        fields = graph_fields(merged)
        for i, gi in enumerate(glist):
            fields[f'part{i}_arcs'], fields[f'part{i}_nodes'], fields[f'part{i}_targets'] = gi.arcs, gi.nodes, gi.targets
        np.savez(f'{OUT}/graphobject_merge_{pb}.npz', **fields)
        # setAggregation on the merged graph
        for mode in ('sum', 'normalized'):
            merged.setAggregation(mode)
            np.savez(f'{OUT}/graphobject_merge_{pb}_{mode}.npz', arcnode_data=merged.ArcNode.data, adj_data=merged.Adjacency.data)

    # 3. isolated nodes + duplicate arcs + custom masks / weights
    nodes = np.arange(12, dtype=float).reshape(6, 2) / 10
    arcs = np.array([[0, 1, .1, .2], [0, 1, .3, .4], [2, 1, .5, .6], [4, 0, .7, .8], [1, 4, .9, 1.], [4, 4, .2, .1]])
    targs = np.eye(2)[[0, 1, 1, 0, 1, 0]]
    for mode in ('average', 'normalized', 'sum'):
        g = GraphObject(arcs=arcs, nodes=nodes, targets=targs[:4], problem_based='n', set_mask=np.array([1, 1, 0, 1, 1, 1]),
                        output_mask=np.array([1, 0, 1, 1, 1, 0]), sample_weights=np.array([1., 2., .5, 3.]), aggregation_mode=mode)
        np.savez(f'{OUT}/graphobject_irregular_{mode}.npz', **graph_fields(g))

    # 4. save / load round trip through the reference's own writer: byte layout of the folder
    with tempfile.TemporaryDirectory() as tmp:
        g = utils.simple_graph('g', 'average')
        g.save(tmp + '/g_npy')
        g.savetxt(tmp + '/g_txt')
        np.savez(f'{OUT}/graphobject_saved_files.npz', npy=np.array(sorted(os.listdir(tmp + '/g_npy'))),
                 txt=np.array(sorted(os.listdir(tmp + '/g_txt'))),
                 arcs_txt=np.array(open(tmp + '/g_txt/arcs.txt').read()))

    # 5. get_inout_dims table (MLP.py:68-122)
    rows = []
    for net in ('state', 'output'):
        for pb in ('n', 'a', 'g'):
            for ds in (0, 3):
                for hidden in (None, 7, [5, 4]):
                    for layer in (0, 1, 3):
                        for gs in (False, True):
                            for go in (False, True):
                                inp, layers = get_inout_dims(net, 3, 2, 4, pb, ds, hidden, layer=layer, get_state=gs, get_output=go)
                                rows.append([net == 'output', 'nag'.index(pb), ds, -1 if hidden is None else (hidden if isinstance(hidden, int) else 54),
                                             layer, gs, go, inp, len(layers), layers[-1]])
    np.savez(f'{OUT}/inout_dims.npz', table=np.array(rows, dtype=np.int64))

    # 6. getindices / getbatches (GNN_utils.py:117-194)
    tr, te, va = utils.getindices(100, 0.7, 0.2, seed=3)
    np.savez(f'{OUT}/getindices.npz', tr=np.array(tr), te=np.array(te), va=np.array(va))
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()
