"""GPU tests of the callers of the hot path: LGNN in parallel / residual / serial
modes against the oracle, the training driver (train / test / LKO, early stopping, Keras-Adam) and model save / load."""
import numpy as np
import pytest
import torch

from tests.parity import random_case, rel_err, TOL, _build_sequential

pytestmark = pytest.mark.gpu


def _require_gpu():
    if not torch.cuda.is_available(): pytest.fail('GPU test selected but no CUDA device is visible')


def _lgnn_setup(problem, get_state, get_output, DS, layers=3, seed=0, n_nodes=160, n_arcs=900, n_graphs=1):
    """ per-layer widths by the reference's rule (MLP.get_inout_dims) + seeded weights, for product and oracle """
    from gnn_b200.MLP import get_inout_dims
    from gnn_b200.graph_class import GraphObject, GraphTensor
    from gnn_b200.GNN import GNNnodeBased, GNNgraphBased
    from gnn_b200.LGNN import LGNN
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    from oracle import gnn_oracle as O
    NL, AL, T = 3, 2, 2
    base = random_case(seed=seed, n_nodes=n_nodes, n_arcs=n_arcs, NL=NL, AL=AL, DS=DS, problem=problem, n_graphs=n_graphs, T=T,
                       masks=(problem == 'n'))
    rng = np.random.default_rng(100 + seed)
    g = GraphObject(arcs=base['arcs'], nodes=base['nodes'], targets=base['targets'], problem_based=problem, set_mask=base['set_mask'],
                    output_mask=base['output_mask'], sample_weights=base['sample_weights'], NodeGraph=base['nodegraph'])
    gt = GraphTensor.fromGraphObject(g)
    cls = {'n': GNNnodeBased, 'g': GNNgraphBased}[problem]
    gnns, onets, x0s = [], [], []
    for layer in range(layers):
        f_s, l_s = get_inout_dims('state', NL, AL, T, problem, DS, None, layer=layer, get_state=get_state, get_output=get_output)
        f_o, l_o = get_inout_dims('output', NL, AL, T, problem, DS, None, layer=layer, get_state=get_state, get_output=get_output)
        ws = [(rng.standard_normal((f_s, l_s[0])) / np.sqrt(f_s)).astype(np.float32), (rng.standard_normal(l_s[0]) * 0.1).astype(np.float32)]
        wo = [(rng.standard_normal((f_o, T)) / np.sqrt(f_o)).astype(np.float32), (rng.standard_normal(T) * 0.1).astype(np.float32)]
        net_s = _build_sequential([f_s] + l_s, ['tanh'], [0.0, 0.0], False, ws, 'cuda')
        net_o = _build_sequential([f_o] + l_o, ['softmax'], [0.0, 0.0], False, wo, 'cuda')
        gnn = cls(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=DS, max_iteration=5,
                  threshold=0.01, addressed_problem='c', path_writer=f'/tmp/gnn_b200_lgnn/{layer}/')
        x0 = (0.1 * rng.standard_normal((n_nodes, DS))).astype(np.float32) if DS else None
        if DS: gnn.initial_state = torch.as_tensor(x0, device='cuda')
        gnns.append(gnn)
        onets.append((O.OracleMLP.from_weights(ws, ['tanh']), O.OracleMLP.from_weights(wo, ['softmax'])))
        x0s.append(None if x0 is None else torch.tensor(x0))
    lgnn = LGNN(gnns, get_state, get_output, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, 'c', path_writer='/tmp/gnn_b200_lgnn/w/')
    src, dst = base['arcs'][:, 0].astype(int), base['arcs'][:, 1].astype(int)
    og = O.OracleGraph.build(base['arcs'], base['nodes'], base['targets'], problem, base['set_mask'], base['output_mask'],
                             base['sample_weights'], base['nodegraph'], 'average', endpoints=(src, dst))
    okw = dict(get_state=get_state, get_output=get_output, state_vect_dim=DS, max_iteration=5, threshold=0.01, x0s=x0s, problem_based=problem)
    return lgnn, gt, og, onets, okw


@pytest.mark.parametrize('problem,get_state,get_output,DS', [('n', False, True, 0), ('n', True, True, 4), ('n', True, False, 0),
                                                           ('g', False, True, 0)])
def test_lgnn_forward_matches_oracle(problem, get_state, get_output, DS):
    _require_gpu()
    from oracle import gnn_oracle as O
    lgnn, gt, og, onets, okw = _lgnn_setup(problem, get_state, get_output, DS, n_graphs=7 if problem == 'g' else 1)
    with torch.no_grad():
        K, state, outs = lgnn.Loop(gt, training=False)
        K2, state2, outs2 = O.lgnn_loop(og, onets, training=False, **okw)
    assert [float(k) for k in K] == [float(k) for k in K2]
    assert rel_err(state.cpu().numpy(), state2.numpy()) < TOL
    for a, b in zip(outs, outs2): assert rel_err(a.cpu().numpy(), b.numpy()) < TOL
    assert np.array_equal(lgnn.predict(gt, idx=-1), outs[-1].cpu().numpy())


@pytest.mark.parametrize('mode', ['parallel', 'residual'])
@pytest.mark.parametrize('get_state,DS', [(False, 0), (True, 4)])
def test_lgnn_training_gradients_match_oracle(mode, get_state, DS):
    """ cross-layer gradients: layer i+1's labels depend on layer i's state / output (LGNN.py:258-259) """
    _require_gpu()
    from oracle import gnn_oracle as O
    lgnn, gt, og, onets, okw = _lgnn_setup('n', get_state, True, DS, seed=3)
    lgnn.training_mode = mode
    iters, loss, targs, out = lgnn.evaluate_single_graph(gt, training=True)
    wS, wO = lgnn.trainable_variables()
    flat = [v for layer in wS + wO for v in layer]
    grads = torch.autograd.grad(loss, flat, allow_unused=True)
    grads = [torch.zeros_like(v) if gr is None else gr for v, gr in zip(flat, grads)]
    K2, loss2, gs2, go2, outs2 = O.lgnn_training_gradients(og, onets, O.categorical_crossentropy, training_mode=mode, mean=True, **okw)
    # the same oracle in float64: rounding-free values, consulted when a float32-vs-float32 difference exceeds the tolerance
    O.set_dtype(torch.float64)
    try:
        _, _, _, onets64, okw64 = _lgnn_setup('n', get_state, True, DS, seed=3)
        from oracle.gnn_oracle import OracleGraph
        og64 = OracleGraph(**{k: (v.double() if isinstance(v, torch.Tensor) and v.dtype == torch.float32 else v) for k, v in og.__dict__.items()})
        _, _, gs64, go64, _ = O.lgnn_training_gradients(og64, onets64, O.categorical_crossentropy, training_mode=mode, mean=True, **okw64)
    finally:
        O.set_dtype(torch.float32)

    def check(got, want32, want64, where):
        e32 = rel_err(got, want32)
        if e32 < TOL: return
        # cancellation-dominated sums (e.g. the 2-class softmax bias gradient: +x and -x terms of size ~1 summing to ~1e-3)
        # carry fp32 summation noise of ~1e-7 * |terms|: accept an absolute error of 1e-6 against the float64 value
        e_got = rel_err(got, want64)
        assert e_got < TOL or float(np.max(np.abs(got - want64))) < 1e-6, (where, e32, e_got, rel_err(want32, want64))

    assert [float(k) for k in iters] == [float(k) for k in K2]
    assert abs(float(loss.detach()) - float(loss2)) <= TOL * abs(float(loss2))
    pos = 0
    for li, layer in enumerate(wS):
        for j, _ in enumerate(layer):
            check((grads[pos] / iters[li]).cpu().numpy(), gs2[li][j].numpy(), gs64[li][j].numpy(), ('state', li, j))
            pos += 1
    for li, layer in enumerate(wO):
        for j, _ in enumerate(layer):
            check(grads[pos].cpu().numpy(), go2[li][j].numpy(), go64[li][j].numpy(), ('output', li, j))
            pos += 1


def _toy_dataset(problem='n', n_graphs=24, seed=0):
    from gnn_b200.graph_class import GraphObject
    rng = np.random.default_rng(seed)
    graphs = []
    for _ in range(n_graphs):
        n = int(rng.integers(10, 25))
        e = 4 * n
        src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
        nodes = rng.uniform(-1, 1, (n, 3))
        arcs = np.concatenate([np.stack([src, dst], 1).astype(float), rng.uniform(-1, 1, (e, 1))], axis=1)
        if problem == 'g':
            label = int(nodes[:, 0].mean() > 0)
            targets = np.eye(2)[[label]]
        else:
            targets = np.eye(2)[(nodes[:, 0] + nodes[:, 1] > 0).astype(int)]
        graphs.append(GraphObject(arcs, nodes, targets, problem_based=problem))
    return graphs


def _make_gnn(problem='n', seed=0, path='/tmp/gnn_b200_train/'):
    from gnn_b200.MLP import MLP, get_inout_dims
    from gnn_b200.GNN import GNNnodeBased, GNNgraphBased
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    from gnn_b200 import GNN_metrics as mt
    f_s, l_s = get_inout_dims('state', 3, 1, 2, problem, 0, None)
    f_o, l_o = get_inout_dims('output', 3, 1, 2, problem, 0, None)
    net_s = MLP(f_s, l_s, 'selu', 'lecun_normal', 'lecun_normal', batch_normalization=False, seed=seed)
    net_o = MLP(f_o, l_o, 'softmax', 'glorot_normal', 'glorot_normal', batch_normalization=False, seed=seed + 1)
    cls = {'n': GNNnodeBased, 'g': GNNgraphBased}[problem]
    return cls(net_s, net_o, Adam(learning_rate=0.01), categorical_crossentropy, {'from_logits': False}, state_vect_dim=0, max_iteration=5,
               threshold=0.01, addressed_problem='c', extra_metrics={'Acc': mt.Metrics['Acc']}, path_writer=path)


def test_train_reduces_loss_and_keeps_history_keys():
    _require_gpu()
    from gnn_b200 import GNN_utils as utils
    graphs = _toy_dataset('n')
    gTr = utils.getbatches(graphs[:16], 'n', 'average', batch_size=8)
    gVa = utils.getbatches(graphs[16:], 'n', 'average', batch_size=8)
    gnn = _make_gnn('n')
    before = gnn.test(gTr)
    gnn.train(gTr, epochs=30, gVa=gVa, update_freq=5, max_fails=50, verbose=0)
    after = gnn.test(gTr)
    assert after['Loss'] < 0.8 * before['Loss'] and after['Acc'] > 0.8
    assert set(gnn.history) == {'Epoch', 'It Tr', 'It Va', 'Loss Tr', 'Loss Va', 'Acc Tr', 'Acc Va', 'Fail', 'Best Loss Va'}
    assert gnn.history['Epoch'] == [0, 5, 10, 15, 20, 25]
    gnn.train(gTr, epochs=5, gVa=gVa, update_freq=5, verbose=0)           # resumes at epoch 30 (GNN_BaseClass.py:278-279)
    assert gnn.history['Epoch'][-1] == 30


@pytest.mark.parametrize('problem', ['n', 'g', 'lgnn'])
def test_cuda_graph_training_step_equals_eager(problem):
    """ use_cuda_graph: the whole training step captured once per batch graph and replayed (while_loop launches, BPTT sweep,
    output net, Adam).  Dropout masks and the Adam step counter must advance at every replay exactly as in eager mode:
    after 4 epochs over 3 batches the weights of a graphed model equal those of an eager twin """
    _require_gpu()
    from gnn_b200 import GNN_utils as utils
    from gnn_b200.MLP import MLP, get_inout_dims
    from gnn_b200.GNN import GNNnodeBased, GNNgraphBased
    from gnn_b200.graph_class import GraphTensor
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    lgnn, problem = problem == 'lgnn', 'n' if problem == 'lgnn' else problem
    graphs = _toy_dataset(problem, n_graphs=18, seed=9)
    batches = [GraphTensor.fromGraphObject(b) for b in utils.getbatches(graphs, problem, 'average', batch_size=6)]

    def one(index, seed, **kw):
        f_s, l_s = get_inout_dims('state', 3, 1, 2, problem, 0, None, **kw)
        f_o, l_o = get_inout_dims('output', 3, 1, 2, problem, 0, None, **kw)
        net_s = MLP(f_s, l_s, 'selu', 'lecun_normal', 'lecun_normal', dropout_rate=0.1, dropout_pos=0, batch_normalization=True, seed=seed)
        net_o = MLP(f_o, l_o, 'softmax', 'glorot_normal', 'glorot_normal', dropout_rate=0.1, dropout_pos=0, batch_normalization=False, seed=seed + 1)
        cls = {'n': GNNnodeBased, 'g': GNNgraphBased}[problem]
        return cls(net_s, net_o, Adam(learning_rate=0.01), categorical_crossentropy, {'from_logits': False}, state_vect_dim=0, max_iteration=5,
                   threshold=0.01, addressed_problem='c', path_writer=f'/tmp/gnn_b200_graphed/{index}/')

    def make():
        if not lgnn: return one(0, 3)
        # 3-layer LGNN (get_output): every layer's loop, the graph updates between the layers and one Adam over all layers in ONE graph;
        # the per-layer GraphTensor copies must find their index tensors cached (torch.nonzero cannot run inside a capture)
        from gnn_b200.LGNN import LGNN
        layers = [one(l, 3 + 2 * l, layer=l, get_state=False, get_output=True) for l in range(3)]
        return LGNN(layers, False, True, Adam(learning_rate=0.01), categorical_crossentropy, {'from_logits': False}, 'c', path_writer='/tmp/gnn_b200_graphed/lgnn/')

    eager, graphed = make(), make()
    graphed.use_cuda_graph = True
    losses = {id(eager): [], id(graphed): []}
    for epoch in range(4):
        for b in batches:
            for m in (eager, graphed):
                iters, loss = m.training_step(b)
                losses[id(m)].append((float(iters[0]), float(loss)))
    assert any(entry[0] == 'graph' for b in batches for entry in b.__dict__['_step_graphs'].values())
    for (k1, l1), (k2, l2) in zip(losses[id(eager)], losses[id(graphed)]):
        assert k1 == k2 and abs(l1 - l2) <= 1e-5 * max(1.0, abs(l1)), (k1, l1, k2, l2)
    weights = lambda m: [w for x in (m.gnns if lgnn else [m]) for w in x.net_state.get_weights() + x.net_output.get_weights()]
    for w1, w2 in zip(weights(eager), weights(graphed)):
        np.testing.assert_allclose(w2, w1, rtol=2e-5, atol=2e-6)


def test_training_step_matches_keras_adam_on_oracle_gradients():
    """ one optimizer step == Keras-Adam formula applied to the oracle's (gradient / k) """
    _require_gpu()
    from gnn_b200.graph_class import GraphTensor
    from oracle import gnn_oracle as O
    graphs = _toy_dataset('n', n_graphs=4, seed=5)
    from gnn_b200.graph_class import GraphObject
    g = GraphObject.merge(graphs, 'n', 'average')
    gnn = _make_gnn('n', seed=7)
    w_before = [w.copy() for w in gnn.net_state.get_weights() + gnn.net_output.get_weights()]
    og = O.OracleGraph.build(g.arcs, g.nodes, g.targets, 'n', g.set_mask, g.output_mask, g.sample_weights, None, 'average', endpoints=(g._src, g._dst))
    net_s = O.OracleMLP.from_weights(gnn.net_state.get_weights(), ['selu'])
    net_o = O.OracleMLP.from_weights(gnn.net_output.get_weights(), ['softmax'])
    k, loss, gs, go, _, _ = O.training_gradients(og, net_s, net_o, O.categorical_crossentropy, state_vect_dim=0, max_iteration=5, threshold=0.01)
    gnn.training_step(GraphTensor.fromGraphObject(g))
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-7
    lr_t = lr * np.sqrt(1 - b2) / (1 - b1)
    for w0, grad, w1 in zip(w_before, [t.numpy() for t in gs + go], gnn.net_state.get_weights() + gnn.net_output.get_weights()):
        m, v = (1 - b1) * grad, (1 - b2) * grad ** 2
        want = w0 - lr_t * m / (np.sqrt(v) + eps)
        assert np.max(np.abs(w1 - want)) < 2e-5


def test_lko_and_graph_based_training():
    _require_gpu()
    from gnn_b200 import GNN_utils as utils
    graphs = _toy_dataset('g', n_graphs=40, seed=2)
    batches = utils.prepare_LKO_data(graphs, 'g', number_of_batches=4, useVa=True, seed=1, normalize_method='gTr')
    gnn = _make_gnn('g', path='/tmp/gnn_b200_lko/')
    res = gnn.LKO(batches, epochs=6, update_freq=3, verbose=0)
    assert set(res) == {'Acc', 'It', 'Loss'} and all(len(v) == 4 for v in res.values())
    assert all(np.isfinite(res['Loss']))


def test_lgnn_serial_training_and_save_load(tmp_path):
    _require_gpu()
    from gnn_b200 import GNN_utils as utils
    from gnn_b200.LGNN import LGNN
    from gnn_b200.GNN import GNNnodeBased
    from gnn_b200.MLP import MLP, get_inout_dims
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    graphs = _toy_dataset('n', n_graphs=12, seed=4)
    gTr = utils.getbatches(graphs, 'n', 'average', batch_size=6)
    gnns = []
    for layer in range(2):
        f_s, l_s = get_inout_dims('state', 3, 1, 2, 'n', 0, None, layer=layer, get_state=False, get_output=True)
        f_o, l_o = get_inout_dims('output', 3, 1, 2, 'n', 0, None, layer=layer, get_state=False, get_output=True)
        gnns.append(GNNnodeBased(MLP(f_s, l_s, 'selu', 'lecun_normal', 'zeros', batch_normalization=False, seed=layer),
                                 MLP(f_o, l_o, 'softmax', 'glorot_normal', 'zeros', batch_normalization=False, seed=10 + layer),
                                 Adam(0.01), categorical_crossentropy, {'from_logits': False}, 0, 5, 0.01, 'c', path_writer=str(tmp_path / f'w{layer}')))
    lgnn = LGNN(gnns, False, True, Adam(0.01), categorical_crossentropy, {'from_logits': False}, 'c', path_writer=str(tmp_path / 'lw'))
    lgnn.train(gTr, epochs=4, update_freq=2, training_mode='serial', verbose=0)
    with pytest.raises(ValueError): lgnn.train(gTr, epochs=1, training_mode='parallel', verbose=0)
    out = lgnn(gTr[0]).cpu().numpy()
    lgnn.save(str(tmp_path / 'model'))
    again = LGNN.load(str(tmp_path / 'model'), path_writer=str(tmp_path / 'lw2'))
    np.testing.assert_allclose(again(gTr[0]).cpu().numpy(), out, rtol=1e-6, atol=1e-7)
    single = gnns[0]
    single.save(str(tmp_path / 'gnn'))
    loaded = GNNnodeBased.load(str(tmp_path / 'gnn'), path_writer=str(tmp_path / 'gw'))
    np.testing.assert_allclose(loaded(gTr[0]).cpu().numpy(), single(gTr[0]).cpu().numpy(), rtol=1e-6, atol=1e-7)
    assert loaded.max_iteration == 5 and loaded.state_threshold == 0.01 and loaded.state_vect_dim == 0


def test_error_behaviour_of_models():
    _require_gpu()
    from gnn_b200.GNN import GNNnodeBased, GNNgraphBased
    from gnn_b200.LGNN import LGNN
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    from gnn_b200 import GNN_utils as utils
    gnn = _make_gnn('n')
    with pytest.raises(TypeError): GNNnodeBased(gnn.net_state, gnn.net_output, Adam(), categorical_crossentropy, None, -1, 5, 0.01, 'c')
    with pytest.raises(ValueError): GNNnodeBased(gnn.net_state, gnn.net_output, Adam(), categorical_crossentropy, None, 0, 5, 0.01, 'x')
    with pytest.raises(TypeError): GNNnodeBased(gnn.net_state, gnn.net_output, Adam(), categorical_crossentropy, None, 0, 5, 0.01, 'c', extra_metrics=[1])
    with pytest.raises(ValueError): gnn.train(utils.simple_graph('n'), 1, verbose=7)
    with pytest.raises(TypeError): gnn.checktype(3)
    ggnn = _make_gnn('g')
    with pytest.raises(ValueError): ggnn.Loop(utils.simple_graph('n'))
    with pytest.raises(TypeError): LGNN([gnn, ggnn], True, True, Adam(), categorical_crossentropy, None, 'c')
