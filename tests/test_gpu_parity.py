"""GPU parity tests proper (run on the B200 box with -m gpu): the CUDA product path, called through the C ABI, against
the CPU oracle on identical seeded inputs.  CSR / index construction and iteration counts exact; states, outputs, loss and
gradients within 1e-4 relative (fp32)."""
import os

import numpy as np
import pytest
import torch

from tests.parity import random_case, run_cuda, run_oracle, assert_parity, build_product, rel_err, TOL

pytestmark = pytest.mark.gpu


def _require_gpu():
    if not torch.cuda.is_available(): pytest.fail('GPU test selected but no CUDA device is visible')


# ---------------------------------------------------------------------------------------------------------------------
# arc preprocessing
# ---------------------------------------------------------------------------------------------------------------------
def _check_csr(coo, with_transpose):
    from gnn_b200 import _native
    from oracle import graph_oracle as GO
    sp = _native.csr_from_coo_transposed(coo, with_transpose=with_transpose)
    want = GO.transposed_row_major(coo.row, coo.col, coo.data, coo.shape)
    np.testing.assert_array_equal(sp.rowptr.cpu().numpy(), want['rowptr'])
    np.testing.assert_array_equal(sp.col.cpu().numpy(), want['indices'][:, 1])
    np.testing.assert_array_equal(sp.values.cpu().numpy(), want['values'])
    np.testing.assert_array_equal(sp.perm.cpu().numpy(), want['perm'])
    np.testing.assert_array_equal(sp.indices.cpu().numpy(), want['indices'])
    assert sp.dense_shape == tuple(want['dense_shape'])
    if with_transpose:
        rowptr_T, col_T, perm_T = GO.csr_transpose(want['rowptr'], want['indices'][:, 1], coo.shape[0])
        np.testing.assert_array_equal(sp.rowptr_T.cpu().numpy(), rowptr_T)
        np.testing.assert_array_equal(sp.col_T.cpu().numpy(), col_T)
        np.testing.assert_array_equal(sp.perm_T.cpu().numpy(), perm_T)
        np.testing.assert_array_equal(sp.values_T.cpu().numpy(), want['values'][perm_T])
    return sp


@pytest.mark.parametrize('mode', ['average', 'normalized', 'sum'])
def test_csr_simple_graph_known_answer(mode):
    _require_gpu()
    from gnn_b200 import GNN_utils as utils
    g = utils.simple_graph('n', mode)
    sp = _check_csr(g.Adjacency, True)
    np.testing.assert_array_equal(sp.rowptr.cpu().numpy(), [0, 2, 4, 7, 8])
    np.testing.assert_array_equal(sp.col.cpu().numpy(), [1, 2, 0, 2, 0, 1, 3, 2])
    assert sp.row_scale is not None
    _check_csr(g.ArcNode, False)


def test_csr_golden_merge(golden_dir):
    _require_gpu()
    from scipy.sparse import coo_matrix
    ref = dict(np.load(f'{golden_dir}/graphobject_merge_n.npz'))
    adj = coo_matrix((ref['adj_data'], (ref['adj_row'], ref['adj_col'])), shape=tuple(ref['adj_shape']))
    sp = _check_csr(adj, True)
    from gnn_b200 import _native
    got = _native.spmm(sp.rowptr, sp.col, sp.values, torch.as_tensor(ref['nodes'], device='cuda'))
    np.testing.assert_allclose(got.cpu().numpy(), ref['adjT_nodes'], rtol=1e-6, atol=1e-6)
    an = coo_matrix((ref['arcnode_data'], (ref['arcnode_row'], ref['arcnode_col'])), shape=tuple(ref['arcnode_shape']))
    sp2 = _check_csr(an, False)
    got = _native.spmm(sp2.rowptr, sp2.col, sp2.values, torch.as_tensor(ref['arcs'][:, 2:].copy(), device='cuda'))
    np.testing.assert_allclose(got.cpu().numpy(), ref['arcnodeT_labels'], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('n,e,seed', [(1, 0, 0), (5, 0, 1), (7, 3, 2), (1000, 20000, 3), (50000, 400000, 4)])
def test_csr_random(n, e, seed):
    _require_gpu()
    from scipy.sparse import coo_matrix
    rng = np.random.default_rng(seed)
    row, col = rng.integers(0, n, e), rng.integers(0, max(n // 2, 1), e)   # upper half of the nodes gets no arc (empty rows)
    if e > 3: row[2], col[2] = row[0], col[0]                              # duplicate entries
    data = rng.random(e).astype(np.float32)
    sp = _check_csr(coo_matrix((data, (row, col)), shape=(n, n)), True)
    if e > 3: assert sp.row_scale is None      # the duplicated entry makes one row non-uniform


def test_row_scale_detection():
    _require_gpu()
    from gnn_b200.graph_class import GraphObject, GraphTensor
    c = random_case(seed=5, n_nodes=300, n_arcs=2000)
    for mode in ('average', 'normalized', 'sum'):
        g = GraphObject(c['arcs'], c['nodes'], c['targets'], set_mask=c['set_mask'], output_mask=c['output_mask'], aggregation_mode=mode)
        gt = GraphTensor.fromGraphObject(g)
        assert gt.Adjacency.row_scale is not None and gt.ArcNode.row_scale is not None
        indeg = np.bincount(c['arcs'][:, 1].astype(int), minlength=300)
        want = {'sum': np.where(indeg > 0, 1.0, 0.0), 'normalized': np.where(indeg > 0, 1 / 2000, 0.0),
                'average': np.where(indeg > 0, 1 / np.maximum(indeg, 1), 0.0)}[mode]
        np.testing.assert_allclose(gt.Adjacency.row_scale.cpu().numpy(), want.astype(np.float32), rtol=1e-7)


def test_default_structure_upload_equals_generic_path():
    """ GraphTensor.fromGraphObject of a default-built graph (endpoints once, arc ids on the device, no read-back) must give
    bit-identical structures and the same arcs matrix as the generic COO path; pinned host buffers change nothing """
    _require_gpu()
    from gnn_b200.graph_class import GraphObject, GraphTensor
    c = random_case(seed=6, n_nodes=500, n_arcs=4000, AL=2)
    for mode in ('average', 'normalized', 'sum'):
        g = GraphObject(c['arcs'], c['nodes'], c['targets'], set_mask=c['set_mask'], output_mask=c['output_mask'], aggregation_mode=mode)
        assert g.has_default_structure()
        fast = GraphTensor.fromGraphObject(g)
        g.pin_host_buffers()
        assert g.has_default_structure()
        pinned = GraphTensor.fromGraphObject(g)
        g._struct_refs = None                               # force the generic path
        assert not g.has_default_structure()
        slow = GraphTensor.fromGraphObject(g)
        for gt in (fast, pinned):
            for name in ('Adjacency', 'ArcNode'):
                a, b = getattr(gt, name), getattr(slow, name)
                for field in ('rowptr', 'col', 'values', 'perm', 'row_scale', 'rowptr_T', 'col_T', 'perm_T'):
                    x, y = getattr(a, field), getattr(b, field)
                    assert (x is None) == (y is None), (name, field)
                    if x is not None: assert torch.equal(x, y), (name, field)
            assert torch.equal(gt.arc_labels, slow.arcs[:, 2:])
            assert torch.equal(gt.arcs, slow.arcs)


# ---------------------------------------------------------------------------------------------------------------------
# forward (inference) parity
# ---------------------------------------------------------------------------------------------------------------------
FWD_CASES = {
    'ds0_nl3': dict(NL=3, AL=1, DS=0, act='tanh', max_iter=5),
    'ds0_nl14_mutag': dict(NL=14, AL=3, DS=0, act='selu', max_iter=5, n_nodes=970, n_arcs=1970),
    'ds0_nl5': dict(NL=5, AL=1, DS=0, act='sigmoid', max_iter=8),
    'ds8': dict(NL=3, AL=2, DS=8, act='tanh', max_iter=30, threshold=0.001),
    'ds32_c4like': dict(NL=3, AL=1, DS=32, act='selu', max_iter=50, threshold=0.0, n_nodes=700, n_arcs=7000, bn=True),
    'ds20_pad32': dict(NL=2, AL=2, DS=20, act='relu', max_iter=10),
    'ds64': dict(NL=3, AL=1, DS=64, act='tanh', max_iter=6),
    'ds100_pad128': dict(NL=1, AL=1, DS=100, act='tanh', max_iter=4, n_nodes=150, n_arcs=900),
    'hidden2': dict(NL=3, AL=2, DS=6, hidden=(10,), act='tanh', max_iter=12),
    'hidden3_elu': dict(NL=4, AL=1, DS=0, hidden=(9, 7), act='elu', max_iter=10),
    'hidden4': dict(NL=2, AL=1, DS=5, hidden=(8, 12, 6), act='softplus', max_iter=7),
    'sum_mode': dict(NL=3, AL=1, DS=4, act='tanh', aggregation='sum', max_iter=10, weight_scale=0.05),
    'normalized_mode': dict(NL=3, AL=1, DS=4, act='tanh', aggregation='normalized', max_iter=10),
    'custom_arcnode': dict(NL=3, AL=2, DS=8, act='tanh', custom_arcnode=True, max_iter=10),
    'bn_inference': dict(NL=3, AL=1, DS=0, act='selu', bn=True, max_iter=5),
    'linear_converges': dict(NL=3, AL=1, DS=3, act='linear', max_iter=60, threshold=0.01, weight_scale=0.05),
    'max_iter0': dict(NL=3, AL=1, DS=2, act='tanh', max_iter=0),
    'tiny': dict(NL=2, AL=1, DS=0, act='tanh', n_nodes=3, n_arcs=4, max_iter=5, isolated=False, duplicates=False, masks=False),
    'no_arcs': dict(NL=2, AL=1, DS=3, act='tanh', n_nodes=40, n_arcs=0, max_iter=5, duplicates=False),
}


@pytest.mark.parametrize('name', sorted(FWD_CASES))
def test_forward_parity(name):
    _require_gpu()
    case = random_case(seed=sorted(FWD_CASES).index(name), **FWD_CASES[name])
    got, want = run_cuda(case, training=False), run_oracle(case, training=False)
    assert_parity(got, want)


@pytest.mark.parametrize('tile', ['128', '32'])
def test_forward_parity_both_tile_shapes(tile, monkeypatch):
    _require_gpu()
    monkeypatch.setenv('GNN_B200_TILE', tile)
    case = random_case(seed=77, n_nodes=1111, n_arcs=9000, NL=3, AL=1, DS=32, act='selu', max_iter=9, threshold=0.0, bn=True)
    assert_parity(run_cuda(case, training=False), run_oracle(case, training=False))
    case = random_case(seed=78, n_nodes=777, n_arcs=5000, NL=14, AL=3, DS=0, act='selu', max_iter=5)
    assert_parity(run_cuda(case, training=False), run_oracle(case, training=False))


def test_forward_large_tiles_and_determinism():
    """ 40k nodes -> 128-node tiles on every SM; two runs must agree bit for bit (no atomics in the data path) """
    _require_gpu()
    case = random_case(seed=9, n_nodes=40000, n_arcs=400000, NL=3, AL=1, DS=32, act='selu', max_iter=6, threshold=0.0, bn=True)
    a = run_cuda(case, training=False)
    b = run_cuda(case, training=False)
    assert a['k'] == b['k'] == 6
    np.testing.assert_array_equal(a['state'], b['state'])
    assert_parity(a, run_oracle(case, training=False))


def test_forward_many_tiles_per_cta():
    """ 150k nodes = 1172 tiles of 128 nodes: more tiles than resident CTAs, every CTA loops over several tiles """
    _require_gpu()
    case = random_case(seed=10, n_nodes=150000, n_arcs=1200000, NL=3, AL=1, DS=32, act='selu', max_iter=4, threshold=0.0, bn=True)
    assert_parity(run_cuda(case, training=False), run_oracle(case, training=False))
    case = random_case(seed=11, n_nodes=150000, n_arcs=600000, NL=14, AL=3, DS=0, act='selu', max_iter=3, threshold=0.0)
    got, want = run_cuda(case, training=True), run_oracle(case, training=True)
    assert_parity(got, want, want64=lambda: run_oracle(case, training=True, float64=True))


def test_graph_and_edge_based_forward():
    _require_gpu()
    case = random_case(seed=21, n_nodes=400, n_arcs=1600, NL=4, AL=2, DS=0, act='tanh', problem='g', n_graphs=13, max_iter=5)
    assert_parity(run_cuda(case, training=False), run_oracle(case, training=False))
    case = random_case(seed=22, n_nodes=120, n_arcs=700, NL=3, AL=2, DS=4, act='tanh', problem='a', max_iter=6)
    assert_parity(run_cuda(case, training=False), run_oracle(case, training=False))


# ---------------------------------------------------------------------------------------------------------------------
# training: BPTT gradients
# ---------------------------------------------------------------------------------------------------------------------
TRAIN_CASES = {
    'ds0_tanh': dict(NL=3, AL=1, DS=0, act='tanh', max_iter=5),
    'ds0_selu_mutag': dict(NL=14, AL=3, DS=0, act='selu', max_iter=5, n_nodes=600, n_arcs=1300, problem='g', n_graphs=20),
    'ds8_tanh': dict(NL=3, AL=2, DS=8, act='tanh', max_iter=12, threshold=0.01),
    'ds32_selu': dict(NL=3, AL=1, DS=32, act='selu', max_iter=10, threshold=0.0, n_nodes=500, n_arcs=5000),
    'custom_arcnode': dict(NL=3, AL=2, DS=8, act='sigmoid', custom_arcnode=True, max_iter=8),
    'sum_mode': dict(NL=2, AL=1, DS=4, act='tanh', aggregation='sum', max_iter=6, weight_scale=0.05),
    'hidden2': dict(NL=3, AL=2, DS=6, hidden=(10,), act='tanh', max_iter=8),
    'hidden3': dict(NL=4, AL=1, DS=0, hidden=(9, 7), act='elu', max_iter=6),
    'dropout_in': dict(NL=3, AL=1, DS=0, act='selu', max_iter=5, drop=[0.1, 0.0]),
    'dropout_everywhere': dict(NL=3, AL=2, DS=6, hidden=(10,), act='tanh', max_iter=6, drop=[0.2, 0.3, 0.1], out_drop=[0.1, 0.0]),
    'bn_train': dict(NL=3, AL=1, DS=0, act='selu', bn=True, max_iter=5),
    'bn_train_ds16': dict(NL=3, AL=1, DS=16, act='tanh', bn=True, max_iter=7, threshold=0.0, n_nodes=900, n_arcs=6000),
    'starter_default': dict(NL=3, AL=1, DS=0, act='selu', bn=True, out_bn=True, drop=[0.1, 0.0], out_drop=[0.1, 0.0], max_iter=5,
                            n_nodes=900, n_arcs=7000),
    'edge_based': dict(NL=3, AL=2, DS=4, act='tanh', problem='a', max_iter=5, n_nodes=100, n_arcs=600),
    'k_zero': dict(NL=3, AL=1, DS=2, act='tanh', max_iter=0),
}


@pytest.mark.parametrize('name', sorted(TRAIN_CASES))
def test_training_parity(name):
    _require_gpu()
    case = random_case(seed=100 + sorted(TRAIN_CASES).index(name), **TRAIN_CASES[name])
    mean = name != 'k_zero'    # gradients / k with k == 0 are not finite in the reference either
    got, want = run_cuda(case, training=True, mean=mean), run_oracle(case, training=True, mean=mean)
    assert_parity(got, want, want64=lambda: run_oracle(case, training=True, mean=mean, float64=True))


@pytest.mark.parametrize('tile', ['128', '32'])
def test_training_parity_tile_shapes(tile, monkeypatch):
    _require_gpu()
    monkeypatch.setenv('GNN_B200_TILE', tile)
    case = random_case(seed=300, n_nodes=10000, n_arcs=60000, NL=3, AL=1, DS=32, act='selu', bn=True, max_iter=4, threshold=0.0)
    assert_parity(run_cuda(case, training=True), run_oracle(case, training=True))


def test_backward_is_deterministic():
    _require_gpu()
    case = random_case(seed=301, n_nodes=5000, n_arcs=40000, NL=3, AL=1, DS=8, act='tanh', max_iter=5)
    a, b = run_cuda(case, training=True), run_cuda(case, training=True)
    for x, y in zip(a['gs'] + a['go'], b['gs'] + b['go']): np.testing.assert_array_equal(x, y)


def test_label_gradients_for_lgnn():
    """ gradients wrt node labels / aggregated labels / x0 (needed by LGNN parallel & residual modes) """
    _require_gpu()
    from gnn_b200.state_loop import state_loop, sparse_dense
    from oracle import gnn_oracle as O
    case = random_case(seed=400, n_nodes=150, n_arcs=900, NL=3, AL=2, DS=5, act='tanh', max_iter=6, masks=False)
    g, gt, gnn = build_product(case)
    nodes = gt.nodes.clone().requires_grad_()
    labels = gt.arcs[:, 2:].clone().requires_grad_()
    x0 = torch.as_tensor(case['x0'], device='cuda').requires_grad_()
    agg_arcs = sparse_dense(gt.ArcNode, labels)
    agg_nodes = sparse_dense(gt.Adjacency, nodes)
    k, x = state_loop(gt.Adjacency, gnn.net_state, x0, nodes, agg_nodes, agg_arcs, max_iteration=6, threshold=0.01, training=True)
    probe = torch.as_tensor(np.random.default_rng(0).standard_normal(tuple(x.shape)).astype(np.float32), device='cuda')
    g_nodes, g_labels, g_x0 = torch.autograd.grad((x * probe).sum(), [nodes, labels, x0])
    # oracle
    og = O.OracleGraph.build(case['arcs'], case['nodes'], case['targets'], 'n', None, None, 1, None, 'average')
    on = og.nodes.clone().requires_grad_()
    ol = og.arcs[:, 2:].clone().requires_grad_()
    ox0 = torch.tensor(case['x0']).requires_grad_()
    net_s = O.OracleMLP.from_weights(case['ws'], case['acts_state'])
    state, state_old, kk = ox0, torch.ones_like(ox0), 0.0
    aa, an = O.spmm(og.arcnode, ol), O.spmm(og.adj, on)
    while O.condition(kk, state.detach(), state_old.detach(), float(np.float32(0.01)), 6):
        inp = torch.cat([state, on, O.spmm(og.adj, state), an, aa], dim=1)
        state, state_old, kk = net_s(inp, True), state, kk + 1
    w_nodes, w_labels, w_x0 = torch.autograd.grad((state * probe.cpu()).sum(), [on, ol, ox0])
    assert float(k) == kk
    assert rel_err(g_nodes.cpu().numpy(), w_nodes.numpy()) < TOL
    assert rel_err(g_labels.cpu().numpy(), w_labels.numpy()) < TOL
    assert rel_err(g_x0.cpu().numpy(), w_x0.numpy()) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# node-range partition (multi-GPU path) on one GPU
# ---------------------------------------------------------------------------------------------------------------------
class _FakePartition:
    """ rows [lo, hi) of a graph, no peers: what dist_graph.GraphPartition hands to state_loop """

    def __init__(self, n_global, lo, calls):
        self.n_global, self.row_offset, self.calls = n_global, lo, calls

    def exchange(self, t, x_full, go_flag):
        self.calls.append((t, tuple(x_full.shape), None if go_flag is None else int(go_flag.numel())))


def test_partition_emulated_ranks_match_single_gpu():
    """ three 'ranks' advance in lock step, one iteration at a time, each computing only its own node range from the
    full previous state; the assembled trajectory must equal the single-GPU loop (same kernels, row_offset mechanics) """
    _require_gpu()
    from gnn_b200 import _native, dist_graph
    from gnn_b200.state_loop import state_loop, sparse_dense
    T = 4
    case = random_case(seed=500, n_nodes=3000, n_arcs=24000, NL=3, AL=2, DS=8, act='tanh', max_iter=T, threshold=0.0, masks=False)
    g, gt, gnn = build_product(case)
    with torch.no_grad():
        k_ref, x_ref, _ = gnn.Loop(gt, training=False)
    assert float(k_ref) == T
    n = 3000
    bounds = dist_graph.partition_bounds(n, 3)
    parts = [dist_graph.GraphPartition(g, r, 3, device='cuda') if False else None for r in range(3)]   # needs a process group: build by hand
    x = torch.as_tensor(case['x0'], device='cuda')
    adj, an = g.Adjacency, g.ArcNode
    local = []
    for r in range(3):
        lo, hi = bounds[r], bounds[r + 1]
        mine = np.nonzero((adj.col >= lo) & (adj.col < hi))[0]
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device='cuda')
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device='cuda')
        A = _native.csr_build(i32(adj.col[mine] - lo), i32(adj.row[mine]), f32(adj.data[mine]), hi - lo, n)
        AN = _native.csr_build(i32(an.col[mine] - lo), i32(np.arange(len(mine))), f32(an.data[mine]), hi - lo, max(len(mine), 1))
        labels = f32(g.arcs[mine, 2:])
        with torch.no_grad():
            local.append((lo, hi, A, sparse_dense(AN, labels), sparse_dense(A, gt.nodes), gt.nodes[lo:hi].contiguous()))
    calls = []
    for t in range(T):
        nxt = torch.empty_like(x)
        for lo, hi, A, agg_arcs, agg_nodes, nodes_self in local:
            with torch.no_grad():
                k, full = state_loop(A, gnn.net_state, x, nodes_self, agg_nodes, agg_arcs, max_iteration=1, threshold=0.0, training=False,
                                     partition=_FakePartition(n, lo, calls))
            assert float(k) == 1.0
            nxt[lo:hi] = full[lo:hi]
        x = nxt
    assert rel_err(x.cpu().numpy(), x_ref.cpu().numpy()) < 1e-6
    assert calls[0] == (0, (n, 8), None)       # one exchange call per iteration; no next flag after the last iteration
    with pytest.raises(NotImplementedError):
        state_loop(local[0][2], gnn.net_state, x, local[0][5], local[0][4], local[0][3], max_iteration=1, threshold=0.0, training=True,
                   partition=_FakePartition(n, 0, calls))


# ---------------------------------------------------------------------------------------------------------------------
# warp-specialised pipelined kernel (state_fwd_ws.cuh), forced on small cases too
# ---------------------------------------------------------------------------------------------------------------------
WS_CASES = {
    'dp32_rowscale': dict(n_nodes=5000, n_arcs=50000, NL=3, AL=1, DS=32, act='selu', max_iter=7, threshold=0.0, bn=True),
    'dp32_perarc': dict(n_nodes=3000, n_arcs=20000, NL=3, AL=2, DS=24, act='tanh', max_iter=6, custom_arcnode=True),
    'dp16_ds0': dict(n_nodes=4000, n_arcs=9000, NL=14, AL=3, DS=0, act='selu', max_iter=5),
    'dp8': dict(n_nodes=1000, n_arcs=6000, NL=5, AL=1, DS=0, act='sigmoid', max_iter=6),
    'tiny_partial_tile': dict(n_nodes=70, n_arcs=900, NL=3, AL=1, DS=16, act='tanh', max_iter=5),
    'hub_node_overflow': dict(n_nodes=600, n_arcs=30000, NL=3, AL=1, DS=32, act='tanh', max_iter=4, aggregation='average'),
    'converging': dict(n_nodes=2000, n_arcs=10000, NL=3, AL=1, DS=8, act='linear', max_iter=60, threshold=0.01, weight_scale=0.05),
}


@pytest.mark.parametrize('kernel', ['tc', 'ws'])
@pytest.mark.parametrize('name', sorted(WS_CASES))
def test_ws_kernel_forward_and_training(name, kernel, monkeypatch):
    """ both pipelined kernels forced on small cases: 'tc' = tcgen05 / tensor-memory Dense layer (state_fwd_tc.cuh; per-arc weights
    fall through to the mma.sync pipeline), 'ws' = mma.sync pipeline (state_fwd_ws.cuh) """
    _require_gpu()
    monkeypatch.setenv('GNN_B200_KERNEL', kernel)
    case = random_case(seed=700 + sorted(WS_CASES).index(name), **WS_CASES[name])
    assert_parity(run_cuda(case, training=False), run_oracle(case, training=False))
    got, want = run_cuda(case, training=True), run_oracle(case, training=True)
    assert_parity(got, want, want64=lambda: run_oracle(case, training=True, float64=True))


def test_ws_and_symmetric_kernels_agree(monkeypatch):
    """ both kernels sum the arcs in stored order; the MLP differs only in rounding (tensor-core 3xTF32 vs fp32 FMA) """
    _require_gpu()
    case = random_case(seed=710, n_nodes=30000, n_arcs=240000, NL=3, AL=1, DS=32, act='selu', max_iter=5, threshold=0.0, bn=True)
    from gnn_b200 import _native
    monkeypatch.setenv('GNN_B200_KERNEL', 'ws')
    a = run_cuda(case, training=False)
    assert _native.last_forward_kernel() == 'state_iter_ws_kernel<32,false>'
    monkeypatch.setenv('GNN_B200_KERNEL', 'sym')
    b = run_cuda(case, training=False)
    assert _native.last_forward_kernel().startswith('state_iter_kernel<32,false')
    monkeypatch.setenv('GNN_B200_KERNEL', 'tc')
    c = run_cuda(case, training=False)      # the tcgen05 pipeline
    assert _native.last_forward_kernel() == 'state_iter_tc_kernel<32>'
    c2 = run_cuda(case, training=False)
    monkeypatch.delenv('GNN_B200_KERNEL')
    d = run_cuda(case, training=False)      # the planner's own choice at this size: the mma.sync pipeline (measured faster, state_loop.cu)
    assert _native.last_forward_kernel() == 'state_iter_ws_kernel<32,false>'
    np.testing.assert_array_equal(d['state'], a['state'])
    np.testing.assert_array_equal(c['state'], c2['state'])      # deterministic
    assert a['k'] == b['k'] == c['k']
    assert rel_err(a["state"], b["state"]) < 2e-5   # 3xTF32 tensor-core product vs sequential fp32 FMA
    assert rel_err(c["state"], b["state"]) < 2e-5   # same, tcgen05


@pytest.mark.parametrize('kw', [dict(DS=32, out_act='softmax', T=2), dict(DS=0, out_act='tanh', T=3), dict(DS=5, out_act='sigmoid', T=16),
                                dict(DS=7, out_act='softmax', T=4, masks=True)],
                         ids=['d32-softmax2', 'labels-tanh3', 'd5-sigmoid16', 'masked-falls-back'])
def test_output_net_one_pass_equals_torch_path(kw):
    """ inference Loop: the one-kernel output net (gnn_output_dense: every node selected, one Dense layer) against the torch path the
    same Loop takes with autograd enabled, and against the oracle """
    _require_gpu()
    base = dict(seed=820, n_nodes=5000, n_arcs=30000, NL=3, AL=2, act='tanh', max_iter=4, threshold=0.0, masks=False)
    base.update(kw)
    case = random_case(**base)
    g, gt, gnn = build_product(case)
    with torch.no_grad():
        k1, x1, out1 = gnn.Loop(gt, training=False, seed=case['seed'])
    k2, x2, out2 = gnn.Loop(gt, training=False, seed=case['seed'])          # autograd on: concat + Dense through torch
    assert float(k1) == float(k2) and torch.equal(x1, x2)
    assert out1.shape == out2.shape
    assert rel_err(out1.cpu().numpy(), out2.detach().cpu().numpy()) < 2e-6
    want = run_oracle(case, training=False)
    assert rel_err(out1.cpu().numpy(), want['out']) < TOL
