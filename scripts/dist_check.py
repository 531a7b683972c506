"""torchrun --nproc-per-node N scripts/dist_check.py : partitioned loop over N GPUs (NCCL) == single-GPU loop"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import gnn_b200
from gnn_b200 import dist_graph
from gnn_b200.graph_class import GraphObject, GraphTensor
from tests.parity import random_case, build_product, rel_err

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok = True
for name, kw in {'uniform': dict(n_nodes=40000, n_arcs=320000), 'converging': dict(n_nodes=30000, n_arcs=150000, max_iter=40, threshold=0.01, weight_scale=0.05),
                 'local': dict(n_nodes=40000, n_arcs=0)}.items():
    base = dict(seed=600, NL=3, AL=2, DS=16, act='tanh', max_iter=6, threshold=0.0, masks=False)
    base.update(kw)
    case = random_case(**base)
    if name == 'local':   # sources within +-50 of the destination: only boundary rows travel
        rng = np.random.default_rng(1)
        dst = np.repeat(np.arange(40000), 6)
        src = (dst + rng.integers(-50, 51, dst.shape[0])) % 40000
        case['arcs'] = np.concatenate([np.stack([src, dst], 1).astype(float), rng.uniform(-1, 1, (dst.shape[0], 2))], axis=1)
    g, gt, gnn = build_product(case, device=f'cuda:{local}')
    with torch.no_grad():
        k_ref, x_ref, out_ref = gnn.Loop(gt, training=False)
    for fused in (True, False):
        part = dist_graph.GraphPartition(g, rank, world, device=f'cuda:{local}', fused=fused)
        for rep in range(2):      # twice: the second call re-uses the symmetric workspace
            k, x, out = dist_graph.partitioned_loop(gnn, part)
        lo, hi = part.row_offset, part.row_offset + part.n_local
        signals = getattr(part, 'in_kernel_signals', False)
        # every rank holds its own rows + the rows it gathers from; other remote rows are only valid when all rows travel
        e_state = rel_err(x.cpu().numpy(), x_ref.cpu().numpy()) if part.halo.use_allgather else rel_err(x[lo:hi].cpu().numpy(), x_ref[lo:hi].cpu().numpy())
        e_out = rel_err(out.cpu().numpy(), out_ref[lo:hi].cpu().numpy())
        good = float(k) == float(k_ref) and e_state < 2e-5 and e_out < 2e-5    # (the partition may run another kernel than the whole graph: rounding of the Dense layer)
        ok &= good
        print(f'[rank {rank}] {name}: requested fused={fused} used fused={part.fused} in-kernel signals={signals} {getattr(part, "_fused_error", "")} k {float(k)} vs {float(k_ref)}, '
              f'state err {e_state:.2e}, out err {e_out:.2e}, rows {"all" if part.halo.use_allgather else "boundary"} -> {"OK" if good else "FAIL"}', flush=True)
    # another net (wider constant row, other max_iteration) over the SAME partition object: the state buffers sit at another offset of
    # the symmetric workspace (the cached peer offsets must follow) and the signal area is re-used with a new epoch
    case2 = random_case(**dict(base, seed=601, AL=1, max_iter=base['max_iter'] + 3))
    case2['arcs'] = np.concatenate([case['arcs'][:, :2], case['arcs'][:, 2:3]], axis=1)
    case2['nodes'], case2['targets'], case2['sample_weights'] = case['nodes'], case['targets'], case['sample_weights']
    g2, gt2, gnn2 = build_product(case2, device=f'cuda:{local}')
    with torch.no_grad():
        k_ref, x_ref, out_ref = gnn2.Loop(gt2, training=False)
    part2 = dist_graph.GraphPartition(g2, rank, world, device=f'cuda:{local}', fused=True)
    for which, gg, pp, ref in (('first net again', gnn, part, None), ('second net', gnn2, part2, (k_ref, x_ref, out_ref))):
        k, x, out = dist_graph.partitioned_loop(gg, pp)
        if ref is None: continue
        lo, hi = pp.row_offset, pp.row_offset + pp.n_local
        e_state = rel_err(x[lo:hi].cpu().numpy(), ref[1][lo:hi].cpu().numpy())
        good = float(k) == float(ref[0]) and e_state < 2e-5
        ok &= good
        print(f'[rank {rank}] {name} / {which}: k {float(k)} vs {float(ref[0])}, state err {e_state:.2e} -> {"OK" if good else "FAIL"}', flush=True)
flag = torch.tensor([0 if ok else 1], device='cuda')
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item()))
