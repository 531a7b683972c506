"""CPU tests of the host logic added in round 2: device-side dropout seeds, the capture-safe Adam step, the strong-scaling shard
assignment of the graph-batch bench and the small-config datasets of bench.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path: sys.path.insert(0, ROOT)

import gnn_b200
from gnn_b200 import keras_compat as K


def test_dropout_key_on_tensor_equals_host_formula():
    """ the dropout key computed with torch ops from a seed TENSOR (captured training steps) == the python-int formula """
    for seed in (0, 1, 0x5EED, 0xFFFFFFFF, 123456789):
        for stream, step in ((0, 0), (1, 3), (16, 0), (2, 49)):
            want = K.dropout_key(seed, stream, step)
            got = int(K.dropout_key_t(torch.tensor(seed, dtype=torch.int64), stream, step))
            assert got == want
    a = K.dropout_keep_mask(0xABCDEF, 1, 2, 37, 5, 0.3)
    b = K.dropout_keep_mask(torch.tensor(0xABCDEF, dtype=torch.int64), 1, 2, 37, 5, 0.3)
    assert torch.equal(a, b) and 0.55 < float(a.float().mean()) < 0.85


def test_device_seed_sequence_equals_host_sequence():
    """ GNN._next_seed with the counter on the 'device' (here: a CPU tensor) walks the same sequence as the host counter """
    from gnn_b200.GNN import GNNnodeBased
    seq = lambda calls, base=0x5EED: (base * 0x9E3779B1 + calls * 0x85EBCA77) & 0xFFFFFFFF

    class Probe:                     # only what _next_seed / _device_seed touch
        dropout_seed, _calls = 0x5EED, 3
        _next_seed, _device_seed = GNNnodeBased._next_seed, GNNnodeBased._device_seed

    p = Probe()
    assert p._next_seed() == seq(4)
    p._device_seed(True, 'cpu')
    got = [int(p._next_seed()) for _ in range(3)]
    assert got == [seq(5), seq(6), seq(7)]
    p._device_seed(False)
    assert p._next_seed() == seq(8)


def test_capturable_adam_equals_host_adam():
    """ same update from the device-side step counter (float64) as from the python counter (the capturable branch needs CUDA
    tensors; its arithmetic is replayed here on CPU tensors) """
    torch.manual_seed(1)
    p1 = [torch.randn(7, 3), torch.randn(3)]
    p2 = [t.clone() for t in p1]
    a = K.Adam(0.01)
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-7
    m = [torch.zeros_like(t) for t in p2]; v = [torch.zeros_like(t) for t in p2]
    t_dev = torch.zeros((), dtype=torch.float64)
    for _ in range(5):
        grads = [torch.randn_like(t) for t in p1]
        a.apply_gradients(zip(grads, p1))
        t_dev += 1.0
        lr_t = (lr * torch.sqrt(1.0 - b2 ** t_dev) / (1.0 - b1 ** t_dev)).to(torch.float32)
        for g, p, mm, vv in zip(grads, p2, m, v):
            mm.mul_(b1).add_(g, alpha=1 - b1); vv.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.sub_(mm / (vv.sqrt() + eps) * lr_t)
    for x, y in zip(p1, p2): assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)


def test_graph_batch_shards_cover_the_dataset_once():
    """ bench.py c5: batch b always comes from seed 1000 + b and every batch belongs to exactly one rank, for every world size """
    for world in (1, 2, 4, 8):
        n_batches = max(world, 200_000 // 25_000)
        owner = [[b for b in range(n_batches) if b * world // n_batches == r] for r in range(world)]
        assert sorted(sum(owner, [])) == list(range(n_batches))
        assert max(map(len, owner)) - min(map(len, owner)) <= 1


def test_small_config_datasets():
    import bench
    gTr, problem, NL, AL, T = bench.make_small_dataset('c1')
    assert (problem, NL, AL, T) == ('n', 3, 1, 2) and len(gTr) == 3                  # 70 training graphs in batches of 32
    assert sum(int(g.nodes.shape[0]) for g in gTr) > 70 * 15
    b = bench.make_graph_batches(40, 40, seed=1)[0]
    gid, coeff, G = b.nodegraph_segments()
    assert G == 40 and b.targets.shape == (40, 2) and b.nodes.shape[1] == 14 and b.arcs.shape[1] == 5
    assert np.all(b._dst[1:] >= 0) and abs(float(coeff.sum()) - 40.0) < 1e-3      # one unit of pooling weight per graph


def test_save_refuses_python_callables(tmp_path):
    """ a regularizer given as a Python callable has no serialisable config: save() raises instead of writing a model that would
    load without it (the reference's Keras save keeps regularizers, GNN.py:93-111) """
    from gnn_b200.keras_compat import Dense, Sequential
    from gnn_b200.GNN import GNNnodeBased
    net = Sequential([Dense(4, activation='tanh', kernel_regularizer=lambda w: (w * w).sum())], input_dim=6, device='cpu')
    with pytest.raises(ValueError, match='callables'):
        GNNnodeBased._save_net(net, str(tmp_path / 'net'))
    plain = Sequential([Dense(4, activation='tanh')], input_dim=6, device='cpu')
    GNNnodeBased._save_net(plain, str(tmp_path / 'plain'))
    again = GNNnodeBased._load_net(str(tmp_path / 'plain'))
    for a, b in zip(plain.get_weights(), again.get_weights()): np.testing.assert_array_equal(a, b)


def test_tall_dense_gradients_equal_plain_autograd():
    """ the block-wise weight gradient of the output net on tall inputs == torch's own (to rounding), same forward bit for bit """
    torch.manual_seed(0)
    x = torch.randn(20011, 35, requires_grad=True)
    k, b = torch.randn(35, 2, requires_grad=True), torch.randn(2, requires_grad=True)
    probe = torch.randn(20011, 2)
    y1 = K.dense_affine(x, k, b)
    assert type(y1.grad_fn).__name__ == '_TallDenseBackward'
    y2 = x @ k + b
    assert torch.equal(y1, y2)
    g1 = torch.autograd.grad((y1 * probe).sum(), [x, k, b])
    g2 = torch.autograd.grad((y2 * probe).sum(), [x, k, b])
    for a, c in zip(g1, g2): assert torch.allclose(a, c, rtol=2e-5, atol=2e-4), float((a - c).abs().max())
    small = K.dense_affine(torch.randn(100, 35, requires_grad=True), k, b)       # short inputs: the plain expression
    assert type(small.grad_fn).__name__ != '_TallDenseBackward'
