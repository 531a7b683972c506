"""debug helper (GPU): training-mode loop on the bench workload at several sizes; prints k, NaN counts, grad norms"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gnn_b200
from bench import make_workload
from gnn_b200.graph_class import GraphObject, GraphTensor
from gnn_b200.GNN import GNNnodeBased
from gnn_b200.keras_compat import Dense, BatchNormalization, Sequential, Adam, categorical_crossentropy

for N in [int(a) for a in sys.argv[1:]] or [20000, 200000, 1000000]:
    wl = make_workload('c4u', N, 10 * N)
    dev = torch.device('cuda')
    net_s = Sequential([Dense(32, activation='selu'), BatchNormalization()], input_dim=71, device=dev)
    net_o = Sequential([Dense(2, activation='softmax')], input_dim=35, device=dev)
    net_s.set_weights(wl['ws']); net_o.set_weights(wl['wo'])
    gnn = GNNnodeBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=32, max_iteration=50,
                       threshold=0.0, addressed_problem='c', path_writer='/tmp/dbg/')
    gnn.initial_state = torch.as_tensor(wl['x0'], device=dev)
    g = GraphObject(arcs=wl['arcs'], nodes=wl['nodes'], targets=wl['targets'], _endpoints=(wl['src'], wl['dst']))
    gt = GraphTensor.fromGraphObject(g)
    with torch.no_grad():
        k, s, o = gnn.Loop(gt, training=False)
    print(N, 'inference k', float(k), 'nan', int(torch.isnan(s).sum()), 'absmax', float(s.abs().max()))
    k, s, o = gnn.Loop(gt, training=True)
    print(N, 'training  k', float(k), 'nan', int(torch.isnan(s).sum()), 'absmax', float(s.abs().max()))
    targs, w = gt.targets, gt.sample_weights
    loss = (categorical_crossentropy(targs, o) * w).sum()
    grads = torch.autograd.grad(loss, net_s.trainable_variables + net_o.trainable_variables)
    print(N, 'loss', float(loss), 'grad absmax', [float(x.abs().max()) for x in grads])
    for step in range(3):
        it, l = gnn.training_step(gt)
        print(N, 'step', step, 'k', float(it[0]), 'loss', float(l), 'W absmax', float(net_s.layers[0].kernel.abs().max()))
