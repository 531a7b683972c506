"""CPU tests of the oracle itself: the hand-worked loop example of SURVEY 8c, Keras layer restatements, the dropout
generator (oracle restatement == product generator), and host-side Keras stand-ins."""
import numpy as np
import pytest
import torch

import gnn_b200
from gnn_b200 import GNN_utils as utils
from gnn_b200 import keras_compat as K
from oracle import gnn_oracle as O


def _simple_case(DS, threshold, max_iter):
    g = utils.simple_graph('n', 'average')
    arcs = g.arcs.astype(np.float64).copy()
    arcs[:, 2:] /= 100
    nodes = g.nodes.astype(np.float64) / 100
    og = O.OracleGraph.build(arcs, nodes, g.targets, 'n', aggregation_mode='average')
    D = DS if DS else 2
    F = 1 + 2 * (2 + DS)
    W = np.array([[((i * D + j) % 7 - 3) * 0.1 for j in range(D)] for i in range(F)], dtype=np.float32)
    b = np.linspace(-0.1, 0.1, D).astype(np.float32)
    net_s = O.OracleMLP.from_weights([W, b], ['tanh'])
    net_o = O.OracleMLP.from_weights([np.zeros((2 + DS, 2), np.float32), np.zeros(2, np.float32)], ['softmax'])
    x0 = None
    if DS: x0 = torch.tensor([[((3 * n + j) % 5 - 2) * 0.05 for j in range(DS)] for n in range(4)], dtype=torch.float32)
    with torch.no_grad():
        return O.loop(og, net_s, net_o, state_vect_dim=DS, max_iteration=max_iter, threshold=threshold, x0=x0)


def test_known_answer_loop_counts():
    """ SURVEY 8c worked example (hand restatement of GNN.py:202-280) """
    k, _, _ = _simple_case(0, 0.01, 5)
    assert k == 5
    k, x, _ = _simple_case(0, 0.01, 50)
    assert k == 6
    np.testing.assert_allclose(x[0].numpy(), [-0.1145415, 0.0577931], atol=2e-6)
    k, _, _ = _simple_case(3, 0.01, 50)
    assert k == 7
    k, x, _ = _simple_case(3, 0.001, 50)
    assert k == 10
    np.testing.assert_allclose(x[0].numpy(), [-0.0169562, -0.0075927, 0.0408705], atol=2e-6)


def test_first_condition_compares_with_ones():
    """ state_old starts as ones (GNN.py:266): a state equal to ones stops the loop before the first iteration """
    g = utils.simple_graph('n', 'sum')
    og = O.OracleGraph.build(g.arcs, np.ones((4, 2)), g.targets, 'n', aggregation_mode='sum')
    net_s = O.OracleMLP.from_weights([np.ones((5, 2), np.float32), np.zeros(2, np.float32)], ['linear'])
    net_o = O.OracleMLP.from_weights([np.ones((2, 2), np.float32), np.zeros(2, np.float32)], ['softmax'])
    with torch.no_grad():
        k, x, _ = O.loop(og, net_s, net_o, state_vect_dim=0, max_iteration=5, threshold=0.01)
    assert k == 0 and torch.equal(x, torch.ones(4, 2))


def test_dropout_generator_restatements_agree():
    for seed, stream, step, rows, cols, rate in [(1, 0, 0, 7, 5, 0.1), (99, 3, 4, 33, 71, 0.5), (2 ** 31 + 5, 17, 49, 4, 3, 0.9)]:
        a = O.keep_mask(seed, stream, step, rows, cols, rate)
        b = K.dropout_keep_mask(seed, stream, step, rows, cols, rate).numpy()
        np.testing.assert_array_equal(a, b)
        assert abs(a.mean() - (1 - rate)) < 0.2


def test_keras_layers_against_oracle():
    """ product Sequential (torch evaluation) == oracle MLP, inference and training (dropout + batch-norm statistics) """
    rng = np.random.default_rng(0)
    net = K.Sequential([K.Dropout(0.3), K.Dense(6, activation='selu'), K.Dropout(0.2), K.Dense(4, activation='softmax'),
                        K.BatchNormalization()], input_dim=5, seed=3)
    w = net.get_weights()
    w[-2][:] = rng.uniform(-0.1, 0.1, 4)
    w[-1][:] = rng.uniform(0.5, 1.5, 4)
    net.set_weights(w)
    omlp = O.OracleMLP.from_weights(w, ['selu', 'softmax'], [0.3, 0.2, 0.0], batchnorm=True)
    x = torch.tensor(rng.standard_normal((40, 5)), dtype=torch.float32)
    np.testing.assert_allclose(net(x, training=False).detach().numpy(), omlp(x, False).detach().numpy(), rtol=1e-6, atol=1e-6)
    got = net(x, training=True, dropout_seed=11, stream_base=16, step=2)
    want = omlp(x, True, seed=11, stream_base=16, step=2)
    np.testing.assert_allclose(got.detach().numpy(), want.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(net.layers[-1].moving_mean.numpy(), omlp.moving_mean.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(net.layers[-1].moving_variance.numpy(), omlp.moving_var.numpy(), rtol=1e-6, atol=1e-7)


def test_keras_semantics_spot_values():
    """ documented constants: selu, categorical cross-entropy clipping, Adam update """
    x = torch.tensor([-1.0, 0.0, 2.0])
    np.testing.assert_allclose(K.apply_activation('selu', x).numpy(),
                               [1.0507009873554805 * 1.6732632423543772 * (np.exp(-1) - 1), 0.0, 2 * 1.0507009873554805], rtol=1e-6)
    t = torch.tensor([[1.0, 0.0]])
    p = torch.tensor([[0.0, 2.0]])           # renormalised to [0, 1] then clipped to 1e-7
    np.testing.assert_allclose(K.categorical_crossentropy(t, p).numpy(), [-np.log(1e-7)], rtol=1e-5)
    np.testing.assert_allclose(O.categorical_crossentropy(t, p).numpy(), [-np.log(1e-7)], rtol=1e-5)
    w = torch.tensor([1.0, -2.0])
    opt = K.Adam(learning_rate=0.1)
    opt.apply_gradients([(torch.tensor([0.5, -0.25]), w)])
    # first Adam step: m = 0.1 g, v = 0.001 g^2, lr_t = lr*sqrt(1-b2)/(1-b1) -> step ~ lr * sign(g)
    np.testing.assert_allclose(w.numpy(), [1.0 - 0.1, -2.0 + 0.1], rtol=1e-5)


def test_mlp_factory_layout():
    from gnn_b200.MLP import MLP
    net = MLP(7, [5, 3], 'selu', 'lecun_normal', 'lecun_normal', dropout_rate=0.1, dropout_pos=0, seed=0)
    kinds = [type(l).__name__ for l in net.layers]
    assert kinds == ['Dropout', 'Dense', 'Dense', 'BatchNormalization']
    spec = net.spec()
    assert spec.dims == [7, 5, 3] and spec.dropout_rates == [0.1, 0.0, 0.0] and spec.batchnorm is not None
    assert [tuple(w.shape) for w in net.get_weights()] == [(7, 5), (5,), (5, 3), (3,), (3,), (3,), (3,), (3,)]
    net2 = MLP(4, [2], 'tanh', 'glorot_normal', 'zeros', dropout_rate=[0.2, 0.3], dropout_pos=[0, 1], batch_normalization=False, seed=0)
    assert [type(l).__name__ for l in net2.layers] == ['Dropout', 'Dense', 'Dropout']
    assert net2.spec().dropout_rates == [pytest.approx(0.2), pytest.approx(0.3)]
    with pytest.raises(ValueError): MLP(4, [2, 3], ['tanh'], 'zeros', 'zeros')
    with pytest.raises(ValueError): MLP(4, [2], 'tanh', 'zeros', 'zeros', dropout_rate=[0.1], dropout_pos=[0, 1])


def test_native_library_exports_every_declared_symbol():
    """ the C-ABI library loads on a CPU-only box and exports what include/gnn_b200.h declares (no compute calls) """
    import re, os
    from gnn_b200 import _native
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'gnn_b200.h')) as f: header = f.read()
    declared = set(re.findall(r'\b(gnn_[a-z_0-9]+)\s*\(', header))
    assert declared == set(_native.EXPORTED_SYMBOLS)
    lib = _native.lib()
    for name in declared: assert hasattr(lib, name), name
    assert lib.gnn_abi_version() == 6


def test_no_cpu_fallback():
    """ the product path refuses to run without a CUDA device """
    if torch.cuda.is_available(): pytest.skip('CUDA present')
    from gnn_b200.graph_class import GraphTensor
    with pytest.raises(RuntimeError): GraphTensor.fromGraphObject(utils.simple_graph('n'))
