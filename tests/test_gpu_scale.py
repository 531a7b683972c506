"""GPU parity at the sizes bench.py TIMES: the C4 graph (1M nodes / 10M arcs, both source
distributions) through the kernel plan the benchmark uses, one C5-shaped merged batch of 5 000 graphs in training, the
5-layer LGNN of config 3, and the data-parallel identity the graph-batch sharding relies on (sum of the per-shard gradients
== gradient of the merged batch).  Oracle: oracle/gnn_oracle.py with its multi-threaded CSR product (seconds per iteration)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path: sys.path.insert(0, ROOT)

from tests.parity import rel_err, TOL

pytestmark = pytest.mark.gpu


def _require_gpu():
    if not torch.cuda.is_available(): pytest.fail('GPU test selected but no CUDA device is visible')


def _c4(name):
    import bench
    from gnn_b200.graph_class import GraphObject, GraphTensor
    from gnn_b200.GNN import GNNnodeBased
    from gnn_b200.keras_compat import Dense, BatchNormalization, Sequential, Adam, categorical_crossentropy
    wl = bench.make_workload(name, 1_000_000, 10_000_000)
    net_s = Sequential([Dense(32, activation='selu'), BatchNormalization()], input_dim=71, device='cuda')
    net_o = Sequential([Dense(2, activation='softmax')], input_dim=35, device='cuda')
    net_s.set_weights(wl['ws']); net_o.set_weights(wl['wo'])
    gnn = GNNnodeBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=32, max_iteration=50,
                       threshold=0.0, addressed_problem='c', path_writer='/tmp/gnn_b200_scale/')
    gnn.initial_state = torch.as_tensor(wl['x0'], device='cuda')
    g = GraphObject(arcs=wl['arcs'], nodes=wl['nodes'], targets=wl['targets'], problem_based='n', aggregation_mode='average',
                    _endpoints=(wl['src'], wl['dst']))
    return bench, wl, gnn, GraphTensor.fromGraphObject(g, device='cuda')


@pytest.mark.parametrize('name,kernel', [('c4u', None), ('c4l', None), ('c4u', 'tc')])
def test_c4_full_size_forward_and_training_parity(name, kernel, monkeypatch):
    """ the timed configuration itself: 1M nodes / 10M arcs / D = 32 -> the pipelined kernel the planner picks (mma.sync pipeline), with
    the ring sized for this graph, and the pipelined backward kernel; once more through the tcgen05 pipeline """
    _require_gpu()
    from gnn_b200 import _native
    if kernel: monkeypatch.setenv('GNN_B200_KERNEL', kernel)
    else: monkeypatch.delenv('GNN_B200_KERNEL', raising=False)
    bench, wl, gnn, gt = _c4(name)
    res = bench.parity_check(wl, gnn, gt, torch.device('cuda'), 2, training=False)
    assert _native.last_forward_kernel() == ('state_iter_tc_kernel<32>' if kernel == 'tc' else 'state_iter_ws_kernel<32,false>')
    assert res['k_equal'] and res['k'] == 2.0
    assert res['max_rel'] <= TOL, res
    assert res['elementwise_rel_p99'] <= TOL, res
    res = bench.parity_check(wl, gnn, gt, torch.device('cuda'), 2, training=True)
    assert res['ok'], res
    assert _native.last_backward_kernel() == 'state_bwd_node_l1_kernel<32>'


def _c5_batch(n_graphs, seed):
    import bench
    return bench.make_graph_batches(n_graphs, n_graphs, seed=seed)[0]


def _c5_nets(device):
    from gnn_b200.keras_compat import Dense, Sequential
    rng = np.random.default_rng(0)
    ws = [(rng.standard_normal((31, 14)) / np.sqrt(31)).astype(np.float32), (0.1 * rng.standard_normal(14)).astype(np.float32)]
    wo = [(rng.standard_normal((14, 2)) / np.sqrt(8)).astype(np.float32), (0.1 * rng.standard_normal(2)).astype(np.float32)]
    net_s = Sequential([Dense(14, activation='selu')], input_dim=31, device=device)
    net_o = Sequential([Dense(2, activation='softmax')], input_dim=14, device=device)
    net_s.set_weights(ws); net_o.set_weights(wo)
    return net_s, net_o, ws, wo


def _gnn_gradients(gnn, gt):
    targs = gnn.get_filtered_tensor(gt, gt.targets)
    weights = gnn.get_filtered_tensor(gt, gt.sample_weights)
    k, state, out = gnn.Loop(gt, training=True)
    loss = (gnn.loss_function(targs, out, **gnn.loss_args) * weights).sum()
    ws, wo = gnn.net_state.trainable_variables, gnn.net_output.trainable_variables
    grads = torch.autograd.grad(loss, ws + wo)
    return float(k), float(loss), [g.cpu().numpy() for g in grads], out.detach().cpu().numpy()


def _c5_gnn(max_iter=5, threshold=0.01):
    from gnn_b200.GNN import GNNgraphBased
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    net_s, net_o, ws, wo = _c5_nets('cuda')
    gnn = GNNgraphBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=0, max_iteration=max_iter,
                        threshold=threshold, addressed_problem='c', path_writer='/tmp/gnn_b200_scale_c5/')
    return gnn, ws, wo


def test_c5_merged_batch_training_parity():
    """ config 5 as benchmarked: one merged batch of 5 000 MUTAG-shaped graphs (~150k nodes / ~290k arcs), graph-focused,
    NodeGraph pooling in segment form, forward + BPTT gradients against the oracle (sparse NodeGraph) """
    _require_gpu()
    from gnn_b200.graph_class import GraphTensor
    from oracle import gnn_oracle as O
    g = _c5_batch(5000, seed=3)
    gnn, ws, wo = _c5_gnn()
    k, loss, grads, out = _gnn_gradients(gnn, GraphTensor.fromGraphObject(g, device='cuda'))
    gid, coeff, G = g.nodegraph_segments()
    og = O.OracleGraph.build(g.arcs, g.nodes, g.targets, 'g', None, None, 1, ('segments', gid, coeff, G), 'average', endpoints=(g._src, g._dst))
    net_s, net_o = O.OracleMLP.from_weights(ws, ['selu']), O.OracleMLP.from_weights(wo, ['softmax'])
    k2, loss2, gs2, go2, out2, _ = O.training_gradients(og, net_s, net_o, O.categorical_crossentropy, mean=False, state_vect_dim=0, max_iteration=5,
                                                        threshold=0.01, problem_based='g', fast_spmm=True)
    assert k == float(k2)
    assert rel_err(out, out2.numpy()) < TOL
    assert abs(loss - float(loss2)) <= TOL * abs(float(loss2))
    for got, want in zip(grads, [t.numpy() for t in gs2 + go2]): assert rel_err(got, want) < TOL


def test_sharded_gradients_equal_merged_batch_gradient():
    """ data-parallel identity behind dist_graph.allreduce_gradients: the loss is a SUM over targets (GNN.py:199)
    and no arc crosses graphs, so with a common iteration count the gradient of the merged batch equals the sum of the
    gradients of its shards.  Two shards emulated on one GPU (threshold 0 => k = max_iteration on every shard) """
    _require_gpu()
    from gnn_b200.graph_class import GraphObject, GraphTensor
    a, b = _c5_batch(1500, seed=11), _c5_batch(1700, seed=12)
    merged = GraphObject.merge([a, b], 'g', 'average')
    gnn, _, _ = _c5_gnn(max_iter=4, threshold=0.0)
    res = [_gnn_gradients(gnn, GraphTensor.fromGraphObject(x, device='cuda')) for x in (a, b, merged)]
    assert res[0][0] == res[1][0] == res[2][0] == 4.0
    assert abs(res[0][1] + res[1][1] - res[2][1]) <= 1e-5 * abs(res[2][1])
    for ga, gb, gm in zip(res[0][2], res[1][2], res[2][2]):
        assert rel_err(ga + gb, gm) < 1e-5
    np.testing.assert_allclose(np.concatenate([res[0][3], res[1][3]]), res[2][3], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('mode', ['parallel', 'residual'])
def test_lgnn_five_layers_matches_oracle(mode):
    """ config 3: FIVE layers, get_state False / get_output True (starter.py:77-79), forward + cross-layer BPTT gradients """
    _require_gpu()
    from oracle import gnn_oracle as O
    from tests.test_gpu_models import _lgnn_setup
    lgnn, gt, og, onets, okw = _lgnn_setup('n', False, True, 0, layers=5, seed=5, n_nodes=400, n_arcs=2400)
    with torch.no_grad():
        K, state, outs = lgnn.Loop(gt, training=False)
        K2, state2, outs2 = O.lgnn_loop(og, onets, training=False, **okw)
    assert len(K) == 5 and [float(k) for k in K] == [float(k) for k in K2]
    assert rel_err(state.cpu().numpy(), state2.numpy()) < TOL
    for x, y in zip(outs, outs2): assert rel_err(x.cpu().numpy(), y.numpy()) < TOL
    lgnn.training_mode = mode
    iters, loss, targs, out = lgnn.evaluate_single_graph(gt, training=True)
    wS, wO = lgnn.trainable_variables()
    flat = [v for layer in wS + wO for v in layer]
    grads = torch.autograd.grad(loss, flat, allow_unused=True)
    grads = [torch.zeros_like(v) if gr is None else gr for v, gr in zip(flat, grads)]
    K2, loss2, gs2, go2, _ = O.lgnn_training_gradients(og, onets, O.categorical_crossentropy, training_mode=mode, mean=True, **okw)
    assert abs(float(loss.detach()) - float(loss2)) <= TOL * abs(float(loss2))
    pos = 0
    for li, layer in enumerate(wS):
        for j, _ in enumerate(layer):
            got, want = (grads[pos] / iters[li]).cpu().numpy(), gs2[li][j].numpy()
            assert rel_err(got, want) < TOL or float(np.max(np.abs(got - want))) < 1e-6, ('state', li, j)
            pos += 1
    for li, layer in enumerate(wO):
        for j, _ in enumerate(layer):
            got, want = grads[pos].cpu().numpy(), go2[li][j].numpy()
            assert rel_err(got, want) < TOL or float(np.max(np.abs(got - want))) < 1e-6, ('output', li, j)
            pos += 1


# ---------------------------------------------------------------------------------------------------------------------
# backward node kernel of state_bwd_l1.cuh (single Dense layer, no dropout, graphs that fill the GPU)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kw', [dict(DS=32, act='selu', bn=True, aggregation='average'),          # the C4 net, training BatchNormalization
                                dict(DS=20, act='tanh', bn=False, aggregation='sum'),             # padded width 32, D = 20, no BN
                                dict(DS=10, act='sigmoid', bn=True, aggregation='normalized'),    # padded width 16
                                dict(DS=0, act='relu', bn=False, aggregation='average', NL=14, AL=3)],   # state = labels (C5 shape), width 16
                         ids=['d32-selu-bn', 'd20-tanh', 'd10-sigmoid-bn', 'labels14-relu'])
def test_pipelined_backward_kernel_matches_oracle_and_phased_kernel(kw, monkeypatch):
    """ 12 345 nodes (not a multiple of the 128-node tile; more than 64 per SM, so the pipelined kernel is chosen): loss and BPTT
    gradients against the oracle, and against the phase-structured kernel (GNN_B200_BWD=phased) on the same inputs """
    _require_gpu()
    from gnn_b200 import _native
    from tests.parity import random_case, run_cuda, run_oracle, assert_parity
    base = dict(seed=900, n_nodes=12345, n_arcs=70000, NL=3, AL=2, max_iter=4, threshold=0.0, weight_scale=None)
    base.update(kw)
    case = random_case(**base)
    monkeypatch.delenv('GNN_B200_BWD', raising=False)
    got = run_cuda(case, training=True)
    assert _native.last_backward_kernel().startswith('state_bwd_node_l1_kernel'), _native.last_backward_kernel()
    monkeypatch.setenv('GNN_B200_BWD', 'phased')
    old = run_cuda(case, training=True)
    assert _native.last_backward_kernel().startswith('state_bwd_node_kernel'), _native.last_backward_kernel()
    want = run_oracle(case, training=True)
    assert_parity(got, want, want64=lambda: run_oracle(case, training=True, float64=True))
    assert got['k'] == old['k'] and got['loss'] == old['loss']          # same forward
    for a, b in zip(got['gs'], old['gs']): assert rel_err(a, b) < 2e-5


def test_pipelined_backward_kernel_label_gradients(monkeypatch):
    """ gradients wrt node labels, arc labels and x0 (LGNN parallel / residual) through the pipelined kernel == phased kernel """
    _require_gpu()
    from gnn_b200 import _native
    from gnn_b200.state_loop import state_loop, sparse_dense
    from tests.parity import random_case, build_product
    case = random_case(seed=901, n_nodes=11000, n_arcs=50000, NL=3, AL=2, DS=12, act='tanh', max_iter=3, threshold=0.0, masks=False)
    g, gt, gnn = build_product(case)
    probe = None
    res = {}
    for mode in ('pipelined', 'phased'):
        if mode == 'phased': monkeypatch.setenv('GNN_B200_BWD', 'phased')
        else: monkeypatch.delenv('GNN_B200_BWD', raising=False)
        nodes = gt.nodes.clone().requires_grad_()
        labels = gt.arcs[:, 2:].clone().requires_grad_()
        x0 = torch.as_tensor(case['x0'], device='cuda').requires_grad_()
        k, x = state_loop(gt.Adjacency, gnn.net_state, x0, nodes, sparse_dense(gt.Adjacency, nodes), sparse_dense(gt.ArcNode, labels),
                          max_iteration=3, threshold=0.0, training=True)
        if probe is None: probe = torch.as_tensor(np.random.default_rng(0).standard_normal(tuple(x.shape)).astype(np.float32), device='cuda')
        res[mode] = [t.cpu().numpy() for t in torch.autograd.grad((x * probe).sum(), [nodes, labels, x0] + gnn.net_state.trainable_variables)]
        assert _native.last_backward_kernel().startswith('state_bwd_node_l1_kernel' if mode == 'pipelined' else 'state_bwd_node_kernel')
    for a, b in zip(res['pipelined'], res['phased']): assert rel_err(a, b) < 2e-5
