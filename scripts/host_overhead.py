"""GPU helper: host-side cost of one training_step / one inference Loop on small batches (C1 / C2 / C5 shapes)"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gnn_b200
from gnn_b200 import _native
from gnn_b200.graph_class import GraphTensor
from bench import make_graph_batches
from gnn_b200.GNN import GNNgraphBased
from gnn_b200.keras_compat import Dense, Sequential, Adam, categorical_crossentropy

dev = torch.device('cuda')
for graphs in (32, 5000):
    b = make_graph_batches(graphs, graphs)[0]
    gt = GraphTensor.fromGraphObject(b)
    net_s = Sequential([Dense(14, activation='selu')], input_dim=31, device=dev, seed=0)
    net_o = Sequential([Dense(2, activation='softmax')], input_dim=14, device=dev, seed=1)
    gnn = GNNgraphBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, 0, 5, 0.01, 'c', path_writer='/tmp/ho/')
    for _ in range(20): gnn.training_step(gt)
    torch.cuda.synchronize()
    _native.launch_count(reset=True)
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): gnn.training_step(gt)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f'graphs {graphs}: nodes {b.nodes.shape[0]} arcs {b.arcs.shape[0]}: training_step host {1e3*t_host/n:.3f} ms, with sync {1e3*t_all/n:.3f} ms, '
          f'library launches/step {_native.launch_count()/n:.0f}')
    with torch.no_grad():
        for _ in range(20): gnn.Loop(gt)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): gnn.Loop(gt)
        t_host = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f'   inference Loop host {1e3*t_host/n:.3f} ms, with sync {1e3*t_all/n:.3f} ms')
    if graphs == 32:
        pr = cProfile.Profile(); pr.enable()
        for _ in range(100): gnn.training_step(gt)
        pr.disable(); torch.cuda.synchronize()
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
