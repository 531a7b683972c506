"""torchrun --nproc-per-node N scripts/dist_e2e_breakdown.py [c4u|c4l] : where the host-to-host time of one partitioned step goes
(every wrapped call is bracketed by a device synchronise, so the parts add up to MORE than the pipelined step)"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import gnn_b200
from gnn_b200 import dist_graph, _native, state_loop as SL
from gnn_b200.graph_class import GraphObject
from gnn_b200.GNN import GNNnodeBased
from gnn_b200.keras_compat import Dense, BatchNormalization, Sequential, Adam, categorical_crossentropy
import bench

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
device = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=device)
name = sys.argv[1] if len(sys.argv) > 1 else 'c4u'
wl = bench.make_workload(name, 1_000_000, 10_000_000)
net_s = Sequential([Dense(wl['DS'], activation='selu'), BatchNormalization()], input_dim=wl['AL'] + 2 * (wl['NL'] + wl['DS']), device=device)
net_o = Sequential([Dense(wl['T'], activation='softmax')], input_dim=wl['NL'] + wl['DS'], device=device)
net_s.set_weights(wl['ws']); net_o.set_weights(wl['wo'])
gnn = GNNnodeBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=wl['DS'],
                   max_iteration=50, threshold=0.0, addressed_problem='c', path_writer=f'/tmp/gnn_b200_prof_{rank}/')
g = GraphObject(arcs=wl['arcs'], nodes=wl['nodes'], targets=wl['targets'], problem_based='n', aggregation_mode='average', _endpoints=(wl['src'], wl['dst']))
g.pin_host_buffers()
x0_host = torch.from_numpy(wl['x0']).pin_memory()

acc = collections.OrderedDict()
def wrap(obj, attr, label):
    fn = getattr(obj, attr)
    def timed(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return out
    setattr(obj, attr, timed)
wrap(_native, 'csr_build', 'csr_build x2')
wrap(torch, 'as_tensor', 'torch.as_tensor (uploads)')
wrap(np, 'ascontiguousarray', 'np.ascontiguousarray')
wrap(np, 'searchsorted', 'np.searchsorted')
wrap(dist_graph, 'partition_bounds', 'bounds')
wrap(dist_graph.HaloPlan, '__init__', 'HaloPlan')
wrap(dist_graph.GraphPartition, '_build_peer_mask', 'peer mask')
wrap(dist_graph.GraphPartition, 'alloc_workspace', 'alloc_workspace')
wrap(dist_graph.GraphPartition, 'peer_setup', 'peer_setup')
wrap(dist_graph, 'partitioned_loop', 'partitioned_loop (all)')
wrap(dist_graph.GraphPartition, '__init__', 'GraphPartition (all)')

def step():
    gnn.initial_state = x0_host.to(device, non_blocking=True)
    p = dist_graph.GraphPartition(g, rank, world, device=device)
    k, state, out = dist_graph.partitioned_loop(gnn, p)
    return out.cpu()

for _ in range(2): step()
for rep in range(2):
    acc.clear()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    step()
    torch.cuda.synchronize(); total = (time.perf_counter() - t0) * 1e3
    print(f'[rank {rank}] {name} x{world}: step {total:.2f} ms | ' + ' | '.join(f'{k} {v:.2f}' for k, v in acc.items()), flush=True)
dist.destroy_process_group()
