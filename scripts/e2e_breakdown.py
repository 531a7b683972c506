"""e2e step of bench.py (host buffers -> gnn(GraphObject) -> host output) split into phases (each synchronised)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gnn_b200
import bench
from gnn_b200.graph_class import GraphObject, GraphTensor
from gnn_b200.GNN import GNNnodeBased
from gnn_b200.keras_compat import Sequential, Dense, BatchNormalization, Adam, categorical_crossentropy

wl = bench.make_workload('c4u', 1_000_000, 10_000_000)
device = torch.device('cuda', 0)
net_s = Sequential([Dense(wl['DS'], activation='selu'), BatchNormalization()], input_dim=wl['AL'] + 2 * (wl['NL'] + wl['DS']), device=device)
net_o = Sequential([Dense(wl['T'], activation='softmax')], input_dim=wl['NL'] + wl['DS'], device=device)
net_s.set_weights(wl['ws']); net_o.set_weights(wl['wo'])
gnn = GNNnodeBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=wl['DS'],
                   max_iteration=50, threshold=0.0, addressed_problem='c', path_writer='/tmp/gnn_b200_e2e/')
g_host = GraphObject(arcs=wl['arcs'], nodes=wl['nodes'], targets=wl['targets'], problem_based='n', aggregation_mode='average',
                     _endpoints=(wl['src'], wl['dst']))
g_host.pin_host_buffers()
x0_host = torch.from_numpy(wl['x0']).pin_memory()

def sync(): torch.cuda.synchronize()
for rep in range(4):
    t = [time.perf_counter()]
    gnn.initial_state = x0_host.to(device, non_blocking=True); sync(); t.append(time.perf_counter())
    gt = GraphTensor.fromGraphObject(g_host, device=device); sync(); t.append(time.perf_counter())
    with torch.no_grad(): k, state, out = gnn.Loop(gt, training=False)
    sync(); t.append(time.perf_counter())
    o = out.cpu(); sync(); t.append(time.perf_counter())
    names = ['x0 h2d', 'fromGraphObject (h2d + csr build)', 'Loop', 'd2h']
    print(rep, ' | '.join(f'{n} {1e3 * (b - a):.2f} ms' for n, a, b in zip(names, t[:-1], t[1:])), f'| total {1e3 * (t[-1] - t[0]):.2f} ms', flush=True)

# inside fromGraphObject (default structure path)
from gnn_b200 import _native
print('default structure:', g_host.has_default_structure(), 'host bytes', g_host.host_bytes())
for rep in range(3):
    t0 = time.perf_counter()
    as_dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=device)
    dst = as_dev(g_host._dst32, np.int32); src = as_dev(g_host._src32, np.int32); val = as_dev(g_host.Adjacency.data, np.float32)
    sync(); t1 = time.perf_counter()
    csr = _native.csr_build(dst, src, val, 1_000_000, 1_000_000, with_transpose=True, assume_uniform=True)
    t1b = time.perf_counter(); sync(); t2 = time.perf_counter()
    val2 = as_dev(g_host.ArcNode.data, np.float32)
    an = _native.csr_build(dst, torch.arange(10_000_000, dtype=torch.int32, device=device), val2, 1_000_000, 10_000_000, assume_uniform=True)
    sync(); t3 = time.perf_counter()
    gt = GraphTensor(nodes=g_host.nodes, arcs=g_host.arcs, targets=g_host.targets, set_mask=g_host.set_mask, output_mask=g_host.output_mask,
                     sample_weights=g_host.sample_weights, NodeGraph=g_host._nodegraph_payload(), Adjacency=csr, ArcNode=an,
                     aggregation_mode=g_host.aggregation_mode, device=device)
    sync(); t4 = time.perf_counter()
    print(f'endpoints+val h2d {1e3*(t1-t0):.2f} | adjacency csr_build(+T) enqueue {1e3*(t1b-t1):.2f} total {1e3*(t2-t1):.2f} | arcnode h2d+build {1e3*(t3-t2):.2f} | dense tensors h2d {1e3*(t4-t3):.2f} ms', flush=True)
