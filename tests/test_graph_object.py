"""GraphObject (host data layer) and the integer oracle against fixtures produced by the REFERENCE's own GraphObject
(oracle/gen_golden.py -> tests/golden/graphobject_*.npz). Bit-exact for indices, exact float32 for weights."""
import glob
import os

import numpy as np
import pytest

import gnn_b200
from gnn_b200.graph_class import GraphObject
from gnn_b200 import GNN_utils as utils
from oracle import graph_oracle as GO


def _load(path): return dict(np.load(path, allow_pickle=False))


def _check_graph(g: GraphObject, ref: dict):
    assert g.arcs.dtype == np.float32 and g.nodes.dtype == np.float32 and g.targets.dtype == np.float32
    np.testing.assert_array_equal(g.arcs, ref['arcs'])
    np.testing.assert_array_equal(g.nodes, ref['nodes'])
    np.testing.assert_array_equal(g.targets, ref['targets'])
    np.testing.assert_array_equal(g.set_mask, ref['set_mask'])
    np.testing.assert_array_equal(g.output_mask, ref['output_mask'])
    np.testing.assert_array_equal(g.sample_weights, ref['sample_weights'])
    assert g.sample_weights.dtype == np.float64
    for name, m in (('arcnode', g.ArcNode), ('adj', g.Adjacency)):
        np.testing.assert_array_equal(m.row, ref[f'{name}_row'])
        np.testing.assert_array_equal(m.col, ref[f'{name}_col'])
        np.testing.assert_array_equal(m.data, ref[f'{name}_data'])
        assert m.data.dtype == np.float32
        assert tuple(m.shape) == tuple(ref[f'{name}_shape'])
    assert [g.DIM_NODE_LABEL, g.DIM_ARC_LABEL, g.DIM_TARGET] == list(ref['dims'])
    if 'NodeGraph' in ref:
        np.testing.assert_array_equal(g.NodeGraph, ref['NodeGraph'])
        assert g.NodeGraph.dtype == np.float32
    else:
        assert g.NodeGraph is None


@pytest.mark.parametrize('pb', ['n', 'a', 'g'])
@pytest.mark.parametrize('mode', ['average', 'normalized', 'sum'])
def test_simple_graph_matches_reference(golden_dir, pb, mode):
    ref = _load(f'{golden_dir}/graphobject_simple_{pb}_{mode}.npz')
    g = utils.simple_graph(pb, mode)
    _check_graph(g, ref)


def test_simple_graph_known_answers():
    """ SURVEY 8c known-answer vectors """
    g = utils.simple_graph('n', 'average')
    np.testing.assert_array_equal(g.ArcNode.col, [1, 2, 0, 2, 0, 1, 3, 2])
    np.testing.assert_allclose(g.ArcNode.data, [.5, 1 / 3, .5, 1 / 3, .5, .5, 1, 1 / 3], rtol=1e-7)
    t = GO.transposed_row_major(g.Adjacency.row, g.Adjacency.col, g.Adjacency.data, g.Adjacency.shape)
    np.testing.assert_array_equal(t['rowptr'], [0, 2, 4, 7, 8])
    np.testing.assert_array_equal(t['indices'][:, 1], [1, 2, 0, 2, 0, 1, 3, 2])
    np.testing.assert_allclose(t['values'], [.5, .5, .5, .5, 1 / 3, 1 / 3, 1 / 3, 1], rtol=1e-7)
    agg = GO.spmm_rows(t['rowptr'], t['indices'][:, 1], t['values'], g.nodes)
    np.testing.assert_allclose(agg, [[12.5, 22.5], [12, 22], [12.333334, 22.333334], [13, 23]], rtol=1e-6)


@pytest.mark.parametrize('mode', ['average', 'normalized', 'sum'])
def test_irregular_graph_matches_reference(golden_dir, mode):
    ref = _load(f'{golden_dir}/graphobject_irregular_{mode}.npz')
    nodes = np.arange(12, dtype=float).reshape(6, 2) / 10
    arcs = np.array([[0, 1, .1, .2], [0, 1, .3, .4], [2, 1, .5, .6], [4, 0, .7, .8], [1, 4, .9, 1.], [4, 4, .2, .1]])
    targs = np.eye(2)[[0, 1, 1, 0, 1, 0]]
    g = GraphObject(arcs=arcs, nodes=nodes, targets=targs[:4], problem_based='n', set_mask=np.array([1, 1, 0, 1, 1, 1]),
                    output_mask=np.array([1, 0, 1, 1, 1, 0]), sample_weights=np.array([1., 2., .5, 3.]), aggregation_mode=mode)
    _check_graph(g, ref)
    # oracle products vs reference (scipy) products
    for name, coo, dense, key in (('adj', g.Adjacency, g.nodes, 'adjT_nodes'), ('arcnode', g.ArcNode, g.arcs[:, 2:], 'arcnodeT_labels')):
        t = GO.transposed_row_major(coo.row, coo.col, coo.data, coo.shape)
        np.testing.assert_allclose(GO.spmm_rows(t['rowptr'], t['indices'][:, 1], t['values'], dense), ref[key], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('pb', ['n', 'g'])
def test_merge_matches_reference(golden_dir, pb):
    ref = _load(f'{golden_dir}/graphobject_merge_{pb}.npz')
    parts = []
    for i in range(5):
        t = ref[f'part{i}_targets']
        parts.append(GraphObject(arcs=ref[f'part{i}_arcs'], nodes=ref[f'part{i}_nodes'], targets=t, problem_based=pb,
                                 output_mask=np.ones(ref[f'part{i}_nodes'].shape[0], dtype=bool), aggregation_mode='average'))
    merged = GraphObject.merge(parts, problem_based=pb, aggregation_mode='average')
    _check_graph(merged, ref)
    for mode in ('sum', 'normalized'):
        merged.setAggregation(mode)
        r2 = _load(f'{golden_dir}/graphobject_merge_{pb}_{mode}.npz')
        np.testing.assert_array_equal(merged.ArcNode.data, r2['arcnode_data'])
        np.testing.assert_array_equal(merged.Adjacency.data, r2['adj_data'])
    # oracle merge restatement
    om = GO.merge([dict(arcs=p.arcs, nodes=p.nodes, targets=p.targets, set_mask=p.set_mask, output_mask=p.output_mask,
                        sample_weights=p.sample_weights, NodeGraph=p.NodeGraph) for p in parts], pb)
    np.testing.assert_array_equal(om['arcs'], ref['arcs'])
    if pb == 'g': np.testing.assert_array_equal(om['NodeGraph'], ref['NodeGraph'])


def test_oracle_structures_match_reference(golden_dir):
    for path in sorted(glob.glob(f'{golden_dir}/graphobject_*_*.npz')):
        ref = _load(path)
        if 'arcs' not in ref or 'part0_arcs' in ref: continue
        name = os.path.basename(path)
        mode = name.rsplit('_', 1)[1][:-4]
        if mode not in ('average', 'normalized', 'sum'): mode = 'average'
        src, dst = ref['arcs'][:, 0].astype(int), ref['arcs'][:, 1].astype(int)
        row, col, data = GO.arcnode_coo(dst, ref['nodes'].shape[0], mode)
        np.testing.assert_array_equal(col, ref['arcnode_col'])
        np.testing.assert_array_equal(data, ref['arcnode_data'])
        arow, acol, adata = GO.adjacency_coo(src, dst, data)
        np.testing.assert_array_equal(arow, ref['adj_row'])
        np.testing.assert_array_equal(acol, ref['adj_col'])
        t = GO.transposed_row_major(arow, acol, adata, ref['adj_shape'])
        np.testing.assert_allclose(GO.spmm_rows(t['rowptr'], t['indices'][:, 1], t['values'], ref['nodes']), ref['adjT_nodes'],
                                   rtol=1e-6, atol=1e-6)


def test_default_structure_flag():
    """ the fast upload path of GraphTensor.fromGraphObject is only taken for matrices the builders made themselves """
    rng = np.random.default_rng(0)
    arcs = np.concatenate([rng.integers(0, 20, (60, 2)).astype(float), rng.random((60, 1))], axis=1)
    nodes, targets = rng.random((20, 3)), rng.random((20, 2))
    g = GraphObject(arcs, nodes, targets)
    assert g.has_default_structure()
    g.setAggregation('sum')
    assert g.has_default_structure()
    custom = GraphObject(arcs, nodes, targets, ArcNode=g.getArcNode())
    assert not custom.has_default_structure()
    g.ArcNode = g.ArcNode.copy()                       # replaced by the user: no assumption survives
    assert not g.has_default_structure()
    g.setAggregation('average')
    assert g.has_default_structure()
    g.Adjacency.data = g.Adjacency.data * 2            # a new array object
    assert not g.has_default_structure()
    assert g.host_bytes() > 0


def test_error_behaviour():
    nodes, arcs, targs = np.zeros((3, 2)), np.array([[0, 1, 1.], [1, 2, 1.]]), np.zeros((3, 2))
    with pytest.raises(ValueError): GraphObject(arcs, nodes, targs, aggregation_mode='max')
    with pytest.raises(ValueError): GraphObject(arcs, nodes, targs, set_mask=np.ones(3), output_mask=np.ones(2))
    g = GraphObject(arcs, nodes, targs)
    with pytest.raises(ValueError): g.setAggregation('mean')
    with pytest.raises(TypeError): GraphObject.merge((g, g), 'n', 'sum')
    assert g.copy().aggregation_mode == 'average'
    # isolated node 0 has an all-zero ArcNode column
    assert g.ArcNode.tocsc()[:, 0].nnz == 0


def test_save_load_roundtrip(tmp_path, golden_dir):
    ref = _load(f'{golden_dir}/graphobject_saved_files.npz')
    g = utils.simple_graph('g', 'average')
    g.save(str(tmp_path / 'npy'))
    g.savetxt(str(tmp_path / 'txt'))
    assert sorted(os.listdir(tmp_path / 'npy')) == list(ref['npy'])
    assert sorted(os.listdir(tmp_path / 'txt')) == list(ref['txt'])
    assert open(tmp_path / 'txt' / 'arcs.txt').read() == str(ref['arcs_txt'])
    g2 = GraphObject.load(str(tmp_path / 'npy'), problem_based='g', aggregation_mode='average')
    np.testing.assert_array_equal(g2.arcs, g.arcs)
    np.testing.assert_array_equal(g2.NodeGraph, g.NodeGraph)
    g3 = GraphObject.load_txt(str(tmp_path / 'txt'), problem_based='g', aggregation_mode='sum')
    np.testing.assert_array_equal(g3.nodes, g.nodes)
    # masks / weights only written when non-default
    n = utils.simple_graph('n', 'average')
    n.set_mask[0] = False
    n.sample_weights[1] = 2.0
    n.save(str(tmp_path / 'npy2'))
    assert sorted(os.listdir(tmp_path / 'npy2')) == ['arcs.npy', 'nodes.npy', 'sample_weights.npy', 'set_mask.npy', 'targets.npy']
    n2 = GraphObject.load(str(tmp_path / 'npy2'), problem_based='n', aggregation_mode='average')
    np.testing.assert_array_equal(n2.set_mask, n.set_mask)
    np.testing.assert_array_equal(n2.sample_weights, n.sample_weights)


def test_large_merge_has_no_dense_nodegraph():
    rng = np.random.default_rng(0)
    parts = []
    for _ in range(300):
        n = int(rng.integers(4, 9))
        arcs = np.stack([rng.integers(0, n, 2 * n), rng.integers(0, n, 2 * n), rng.random(2 * n)], axis=1)
        parts.append(GraphObject(arcs, rng.random((n, 2)), np.eye(2)[[0]], problem_based='g'))
    m = GraphObject.merge(parts, 'g', 'average')
    ids, coeff, n_graphs = m.nodegraph_segments()
    assert n_graphs == 300 and len(ids) == m.nodes.shape[0]
    sizes = np.bincount(ids)
    np.testing.assert_allclose(coeff, (1.0 / sizes[ids]).astype(np.float32))
    dense = m.NodeGraph
    assert dense.shape == (m.nodes.shape[0], 300)
    np.testing.assert_allclose(dense.sum(axis=0), 1.0, rtol=1e-6)


def test_inout_dims_match_reference(golden_dir):
    from gnn_b200.MLP import get_inout_dims
    table = _load(f'{golden_dir}/inout_dims.npz')['table']
    for is_out, pb, ds, hid, layer, gs, go, inp, nlay, last in table:
        hidden = None if hid == -1 else ([5, 4] if hid == 54 else int(hid))
        got_inp, got_layers = get_inout_dims('output' if is_out else 'state', 3, 2, 4, 'nag'[pb], int(ds), hidden,
                                             layer=int(layer), get_state=bool(gs), get_output=bool(go))
        assert (got_inp, len(got_layers), got_layers[-1]) == (inp, nlay, last)


def test_getindices_matches_reference(golden_dir):
    ref = _load(f'{golden_dir}/getindices.npz')
    tr, te, va = utils.getindices(100, 0.7, 0.2, seed=3)
    np.testing.assert_array_equal(tr, ref['tr'])
    np.testing.assert_array_equal(te, ref['te'])
    np.testing.assert_array_equal(va, ref['va'])


@pytest.mark.parametrize('pb', ['n', 'g'])
def test_random_graph_reproduces_reference_with_seed(golden_dir, pb):
    """ same legacy NumPy seed, same consumption order -> identical graph as the reference's randomGraph """
    ref = _load(f'{golden_dir}/graphobject_random_{pb}.npz')
    np.random.seed(7)
    glist = [utils.randomGraph(int(n), 3, 1, 2, 0.7, aggregation_mode='average', problem_based=pb) for n in (15, 22, 17, 30, 9)]
    _check_graph(glist[1], ref)


@pytest.mark.skipif(not os.path.exists('/root/reference/MUTAG_raw/Mutagenicity_edges.txt'), reason='MUTAG raw files are not shipped')
def test_mutag_loader_counts():
    """ restated loader (NumPy 2) reproduces the dataset statistics measured with the reference recipe """
    from gnn_b200.load_MUTAG import load_MUTAG
    graphs = load_MUTAG('/root/reference/MUTAG_raw/')
    assert len(graphs) == 4337
    assert sum(g.nodes.shape[0] for g in graphs) == 131488 and sum(g.arcs.shape[0] for g in graphs) == 266894
    assert graphs[0].DIM_NODE_LABEL == 14 and graphs[0].DIM_ARC_LABEL == 3 and graphs[0].DIM_TARGET == 2
    assert graphs[0].NodeGraph.shape == (graphs[0].nodes.shape[0], 1)
