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
        # every rank holds its own rows + the rows it gathers from; other remote rows are only valid when all rows travel
        e_state = rel_err(x.cpu().numpy(), x_ref.cpu().numpy()) if part.halo.use_allgather else rel_err(x[lo:hi].cpu().numpy(), x_ref[lo:hi].cpu().numpy())
        e_out = rel_err(out.cpu().numpy(), out_ref[lo:hi].cpu().numpy())
        good = float(k) == float(k_ref) and e_state < 1e-6 and e_out < 1e-6
        ok &= good
        print(f'[rank {rank}] {name}: requested fused={fused} used fused={part.fused} {getattr(part, "_fused_error", "")} k {float(k)} vs {float(k_ref)}, '
              f'state err {e_state:.2e}, out err {e_out:.2e}, rows {"all" if part.halo.use_allgather else "boundary"} -> {"OK" if good else "FAIL"}', flush=True)
flag = torch.tensor([0 if ok else 1], device='cuda')
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item()))
