"""torchrun --nproc-per-node N scripts/dist_profile.py [c4u|c4l] : where the time of one partitioned loop goes (per rank):
whole call, the iteration launches alone (events inside the library), and the host time spent enqueueing the call"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import gnn_b200
from gnn_b200 import dist_graph, _native
from gnn_b200.graph_class import GraphObject
from gnn_b200.GNN import GNNnodeBased
from gnn_b200.keras_compat import Dense, BatchNormalization, Sequential, Adam, categorical_crossentropy
import bench

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
device = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=device)
name = sys.argv[1] if len(sys.argv) > 1 else 'c4u'
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 50
wl = bench.make_workload(name, 1_000_000, 10_000_000)
net_s = Sequential([Dense(wl['DS'], activation='selu'), BatchNormalization()], input_dim=wl['AL'] + 2 * (wl['NL'] + wl['DS']), device=device)
net_o = Sequential([Dense(wl['T'], activation='softmax')], input_dim=wl['NL'] + wl['DS'], device=device)
net_s.set_weights(wl['ws']); net_o.set_weights(wl['wo'])
gnn = GNNnodeBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=wl['DS'],
                   max_iteration=max_iter, threshold=0.0, addressed_problem='c', path_writer=f'/tmp/gnn_b200_prof_{rank}/')
gnn.initial_state = torch.as_tensor(wl['x0'], device=device)
g = GraphObject(arcs=wl['arcs'], nodes=wl['nodes'], targets=wl['targets'], problem_based='n', aggregation_mode='average', _endpoints=(wl['src'], wl['dst']))
part = dist_graph.GraphPartition(g, rank, world, device=device, fused=os.environ.get('GNN_B200_FUSED', '1') != '0')
for _ in range(3): dist_graph.partitioned_loop(gnn, part)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    _native.profile_iterations(True)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); a.record()
    k, x, out = dist_graph.partitioned_loop(gnn, part)
    b.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms_iter, n = _native.profile_last_iterations()
    _native.profile_iterations(False)
    print(f'[rank {rank}] {name} x{world} max_iter {max_iter} fused={part.fused} signals={part.in_kernel_signals} dbg={os.environ.get('GNN_B200_WS_DEBUG', '')} kernel={_native.last_forward_kernel()}: call {a.elapsed_time(b):.3f} ms, '
          f'iteration launches {ms_iter:.3f} ms / {n} = {ms_iter / max(n, 1):.4f} ms each, host enqueue {1e3 * (t1 - t0):.3f} ms, k {float(k)}', flush=True)
dist.destroy_process_group()
