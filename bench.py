#!/usr/bin/env python
"""bench.py -- the state-convergence loop on BASELINE.json's headline workload (C4: one synthetic graph, 1M nodes /
10M arcs, node classification, state_dim 32, max_iteration 50), metric arc-updates/sec = arcs x iterations / time.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4u|c4l] [--nodes N --arcs E]

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  value        forward Loop (training=False, threshold 0 so that k = max_iteration), inputs resident in HBM
  e2e          same metric through the public call gnn(GraphObject) with HOST buffers: host->device copy of the graph
               arrays (pinned), CSR build, loop, device->host read of the outputs, every step
  train        forward + BPTT backward + optimizer (BaseClass.training_step) arc-updates/sec and epoch time
  roofline     fused iteration kernel: algorithmic bytes per launch / mean launch time (CUDA events on the launching
               stream inside libgnn_b200.so) against MEASURED_PEAKS.json
  cpu_baseline the oracle port (torch-CPU restatement of the reference's TF path, NOT TensorFlow) on the host cores,
               bounded sample of the same workload
`--impl reference` times that CPU port alone (the reference's TensorFlow cannot be installed: see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path: sys.path.insert(0, ROOT)

# algorithmic bytes per arc-update of the fused forward iteration (SURVEY.md 8d / DESIGN.md): every array touched
# once, gathered state re-used from L2:  4 E (source index) + N (4 rowptr + 4 D read + 4 D write + 4 (2 NL + AL))
def algorithmic_bytes_per_iteration(N, E, D, NL, AL, per_arc_weights=False):
    return 4 * E * (2 if per_arc_weights else 1) + N * (4 + 8 * D + 4 * (2 * NL + AL))


def progress(msg: str) -> None:
    """ one stderr line per leg (with the rank and the wall clock): a multi-GPU run that stalls shows where """
    print(f'[bench rank {os.environ.get("RANK", "0")} +{time.perf_counter() - _T0:7.1f}s] {msg}', file=sys.stderr, flush=True)


_T0 = time.perf_counter()


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
def make_workload(name: str, N: int, E: int, seed: int = 0):
    """ in-degree exactly E/N per node; sources uniform ('c4u') or within +-2048 of the destination
    ('c4l'); NL=3, AL=1, T=2, labels U(-1,1); x0 = 0.1 randn; net_state Dense(71->32, selu)+BN, randn*sqrt(1/71) """
    rng = np.random.default_rng(seed)
    deg = E // N
    dst = np.repeat(np.arange(N, dtype=np.int64), deg)
    if name == 'c4l':
        delta = rng.integers(1, 2049, dst.shape[0]) * rng.choice([-1, 1], dst.shape[0])
        src = (dst + delta) % N
    else:
        src = rng.integers(0, N, dst.shape[0])
    NL, AL, T, DS = 3, 1, 2, 32
    arcs = np.empty((dst.shape[0], 2 + AL), dtype=np.float32)
    arcs[:, 0], arcs[:, 1] = src, dst
    arcs[:, 2:] = rng.uniform(-1, 1, (dst.shape[0], AL))
    nodes = rng.uniform(-1, 1, (N, NL)).astype(np.float32)
    targets = np.eye(T, dtype=np.float32)[np.argmax(nodes[:, :2], axis=1)]
    x0 = (0.1 * rng.standard_normal((N, DS))).astype(np.float32)
    F = AL + 2 * (NL + DS)
    ws = [(rng.standard_normal((F, DS)) * np.sqrt(1.0 / F)).astype(np.float32), np.zeros(DS, np.float32),
          np.ones(DS, np.float32), np.zeros(DS, np.float32), np.zeros(DS, np.float32), np.ones(DS, np.float32)]
    Fo = NL + DS
    # output net: Dense(softmax) only. The reference default appends BatchNormalization after the softmax; with T = 2
    # the two normalised columns are exact opposites, the loss divides by their sum (= 0) and every gradient is NaN
    wo = [(rng.standard_normal((Fo, T)) * np.sqrt(2.0 / (Fo + T))).astype(np.float32), np.zeros(T, np.float32)]
    return dict(arcs=arcs, nodes=nodes, targets=targets, x0=x0, ws=ws, wo=wo, src=src, dst=dst, NL=NL, AL=AL, T=T, DS=DS, N=N,
                E=int(dst.shape[0]))


def make_graph_batches(n_graphs: int, graphs_per_step: int, seed: int = 0):
    """ MUTAG-shaped molecule-like graphs (N_g ~ clip(round(lognormal), 4, 417), mean ~30; symmetric sparse
    arcs ~2.03 N_g: a chain plus a few chords), NL=14 one-hot, AL=3 one-hot, T=2, graph targets Bernoulli(.5); returned as
    merged batches (disjoint unions) of `graphs_per_step` graphs, built directly as arrays """
    from gnn_b200.graph_class import GraphObject
    rng = np.random.default_rng(seed)
    batches = []
    for start in range(0, n_graphs, graphs_per_step):
        G = min(graphs_per_step, n_graphs - start)
        sizes = np.clip(np.rint(rng.lognormal(np.log(26.0), 0.5, G)), 4, 417).astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        N = int(offs[-1])
        gid = np.repeat(np.arange(G), sizes)
        local = np.arange(N) - offs[gid]
        chain = np.nonzero(local < sizes[gid] - 1)[0]                     # i -> i+1 inside a graph
        n_chords = np.maximum(1, np.rint(0.015 * sizes)).astype(np.int64)
        cg = np.repeat(np.arange(G), n_chords)
        ca = offs[cg] + (rng.random(len(cg)) * sizes[cg]).astype(np.int64)
        cb = offs[cg] + (rng.random(len(cg)) * sizes[cg]).astype(np.int64)
        keep = ca != cb
        a = np.concatenate([chain, ca[keep]]); b = np.concatenate([chain + 1, cb[keep]])
        lab = rng.integers(0, 3, len(a))
        src, dst = np.concatenate([a, b]), np.concatenate([b, a])
        order = np.lexsort((dst, src))
        src, dst, lab2 = src[order], dst[order], np.concatenate([lab, lab])[order]
        arcs = np.zeros((len(src), 5), dtype=np.float32)
        arcs[:, 0], arcs[:, 1] = src, dst
        arcs[np.arange(len(src)), 2 + lab2] = 1
        nodes = np.eye(14, dtype=np.float32)[rng.integers(0, 14, N)]
        targets = np.eye(2, dtype=np.float32)[rng.integers(0, 2, G)]
        nodegraph = ('segments', gid, (1.0 / sizes[gid]).astype(np.float32), G)
        batches.append(GraphObject(arcs=arcs, nodes=nodes, targets=targets, problem_based='g', NodeGraph=nodegraph,
                                   aggregation_mode='average', _endpoints=(src, dst)))
    return batches


def bench_graph_batches(args, device, rank, world):
    """ C5 leg: 200 000 MUTAG-shaped graphs as merged batches of 25 000, sharded by whole batch over the ranks (STRONG scaling: the
    same dataset at every N; N = 1 trains all of it).  One training_step (forward + BPTT + flat gradient all-reduce + Adam) per
    batch, each captured once as a CUDA graph and replayed; step = one epoch of the rank's shard """
    import torch
    import torch.distributed as dist
    from gnn_b200 import _native
    from gnn_b200.graph_class import GraphTensor
    from gnn_b200.GNN import GNNgraphBased
    from gnn_b200.keras_compat import Dense, Sequential, Adam, categorical_crossentropy
    total, per_step = args.graphs_total, args.graphs_per_step
    progress(f'c5: generating the graph batches of rank {rank}')
    n_batches = max(world, total // per_step)
    mine = [b for b in range(n_batches) if b * world // n_batches == rank]          # contiguous blocks of batches per rank
    gts, arcs = [], 0
    for b in mine:                                   # the dataset does not depend on N: batch b is always drawn from seed 1000 + b
        g = make_graph_batches(per_step, per_step, seed=1000 + b)[0]
        arcs += int(g.arcs.shape[0])
        gts.append(GraphTensor.fromGraphObject(g, device=device))
        del g
    rng = np.random.default_rng(0)     # identical (replicated) initial weights on every rank
    net_s = Sequential([Dense(14, activation='selu')], input_dim=31, device=device)
    net_o = Sequential([Dense(2, activation='softmax')], input_dim=14, device=device)
    net_s.set_weights([(rng.standard_normal((31, 14)) / np.sqrt(31)).astype(np.float32), np.zeros(14, np.float32)])
    net_o.set_weights([(rng.standard_normal((14, 2)) / np.sqrt(8)).astype(np.float32), np.zeros(2, np.float32)])
    gnn = GNNgraphBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=0, max_iteration=5,
                        threshold=0.01, addressed_problem='c', path_writer=f'/tmp/gnn_b200_bench_c5_{rank}/')
    gnn.distributed = world > 1
    gnn.use_cuda_graph = not args.no_cuda_graph
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks = []

    def epoch():
        for gt in gts:
            iters, _ = gnn.training_step(gt)
            ks.append(iters[0])

    progress(f'c5: {len(gts)} batches on the device, warm-up / capture')
    for _ in range(max(3, args.warmup)): epoch()     # (CUDA graph: the eager first step and the capture of every batch happen here)
    progress('c5: timed epochs')
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    _native.launch_count(reset=True)
    start.record()
    for _ in range(args.steps): epoch()
    stop.record()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    ms = torch.tensor([start.elapsed_time(stop) / args.steps], device=device)
    upd = torch.tensor([float(sum(float(k) for k in ks[-len(gts):]) / max(len(gts), 1)) * arcs], device=device)   # arcs x iterations of one epoch
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(upd, op=dist.ReduceOp.SUM)
    # fwd + bwd algorithmic bytes per arc-update of this shape (67 B forward; backward 4 E + N (4 + 16 D + 12)): ~190 B
    value = float(upd.item()) / (float(ms.item()) * 1e-3)
    peak = 6553.9
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f: peak = float(json.load(f).get('hbm_gbs', peak))
    except Exception:
        pass
    return {'value': value, 'ms_per_step': float(ms.item()), 'epoch_time_s': float(ms.item()) * 1e-3,
            'gpu_launches': int(_native.launch_count() // args.steps), 'roofline_frac': value * 190.0 / 1e9 / (peak * world),
            'config': {'workload': f'c5: {n_batches * per_step} MUTAG-shaped graphs in merged batches of {per_step}, sharded by whole batch over {world} '
                                   f'rank(s) ({len(mine)} batch(es) per rank), graph classification, NL 14, AL 3, state = labels (D 14), max_iteration 5, '
                                   f'threshold 0.01; one training_step (forward + BPTT + gradient all-reduce + Adam) per batch, '
                                   f'{"replayed as a CUDA graph" if gnn.use_cuda_graph else "eager launches"}; step = one epoch of the dataset',
                       'arcs_per_rank': arcs, 'l2': 'inputs larger than L2 (25 000 graphs = 760k nodes x 56-byte rows x 11 saved iterates)'}}


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = 'clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int = 0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout: self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try: self.proc.wait(timeout=2)
            except Exception: self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith('active'): reasons.add(name)
        if not sm: return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU port (oracle) -- cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_run(wl, iterations: int, repeats: int = 1):
    """ forward loop of the oracle restatement (torch-CPU, all host threads) for `iterations` iterations of the workload.
    :return: (arc-updates per second, seconds per run, threads) """
    import torch
    from oracle import gnn_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = O.OracleGraph.build(wl['arcs'], wl['nodes'], wl['targets'], 'n', None, None, 1, None, 'average', endpoints=(wl['src'], wl['dst']))
    net_s = O.OracleMLP.from_weights(wl['ws'], ['selu'], batchnorm=True, requires_grad=False)
    net_o = O.OracleMLP.from_weights(wl['wo'], ['softmax'], batchnorm=False, requires_grad=False)
    x0 = torch.from_numpy(wl['x0'])
    best = float('inf')
    k = 0
    for rep in range(repeats + 1):   # first run warms the CSR cache / allocator and is not timed
        t0 = time.perf_counter()
        with torch.no_grad():
            k = O.loop(g, net_s, net_o, state_vect_dim=wl['DS'], max_iteration=iterations, threshold=0.0, x0=x0, fast_spmm=True)[0]
        dt = time.perf_counter() - t0
        if rep > 0: best = min(best, dt)
    return wl['E'] * k / best, best, threads



# ---------------------------------------------------------------------------------------------------------------------
# parity of the TIMED configuration against the CPU oracle (the checker leg; never inside a timed region)
# ---------------------------------------------------------------------------------------------------------------------
def _tensor_errors(got, want):
    """ (max-norm relative error of the tensor, 99th percentile and maximum of the ELEMENT-WISE relative error
    |got - want| / (|want| + 1e-3 max|want|): the floor keeps elements near zero from dividing by nothing) """
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = float(np.max(np.abs(want))) if want.size else 0.0
    diff = np.abs(got - want)
    elem = diff / (np.abs(want) + 1e-3 * scale + 1e-30)
    return float(diff.max() / (scale + 1e-6)), float(np.percentile(elem, 99)), float(elem.max())


def parity_check(wl, gnn, gt, device, iterations: int, training: bool, tol: float = 1e-4):
    """ `iterations` iterations of the loop on the full-size workload through the GPU path and through the oracle port:
    iteration count exact, state / outputs (and, training=True, loss + BPTT gradients + moving statistics) within tol.
    :return: dict for the JSON line; 'ok' False makes bench.py refuse to print a value """
    import torch
    from oracle import gnn_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = O.OracleGraph.build(wl['arcs'], wl['nodes'], wl['targets'], 'n', None, None, 1, None, 'average', endpoints=(wl['src'], wl['dst']))
    net_s = O.OracleMLP.from_weights(wl['ws'], ['selu'], batchnorm=True, requires_grad=training)
    net_o = O.OracleMLP.from_weights(wl['wo'], ['softmax'], batchnorm=False, requires_grad=training)
    x0 = torch.from_numpy(wl['x0'])
    kw = dict(state_vect_dim=wl['DS'], max_iteration=iterations, threshold=0.0, x0=x0, fast_spmm=True)
    saved_iter, saved_w = gnn.max_iteration, [w.copy() for w in gnn.net_state.get_weights()]
    gnn.max_iteration = iterations
    tensors = {}
    try:
        if not training:
            with torch.no_grad():
                k_ref, state_ref, out_ref = O.loop(g, net_s, net_o, **kw)
                k, state, out = gnn.Loop(gt, training=False)
            tensors = {'state': (state.cpu().numpy(), state_ref.numpy()), 'out': (out.cpu().numpy(), out_ref.numpy())}
        else:
            k_ref, loss_ref, gs_ref, go_ref, out_ref, state_ref = O.training_gradients(g, net_s, net_o, O.categorical_crossentropy, mean=True, **kw)
            targs = gnn.get_filtered_tensor(gt, gt.targets)
            weights = gnn.get_filtered_tensor(gt, gt.sample_weights)
            k, state, out = gnn.Loop(gt, training=True)
            loss = (gnn.loss_function(targs, out, **gnn.loss_args) * weights).sum()
            ws, wo = gnn.net_state.trainable_variables, gnn.net_output.trainable_variables
            grads = torch.autograd.grad(loss, ws + wo, allow_unused=True)
            grads = [torch.zeros_like(v) if gr is None else gr for v, gr in zip(ws + wo, grads)]
            tensors = {'state': (state.detach().cpu().numpy(), state_ref.numpy()), 'out': (out.detach().cpu().numpy(), out_ref.numpy()),
                       'loss': (np.array([float(loss)]), np.array([float(loss_ref)]))}
            for i, (gr, ref) in enumerate(zip(grads[:len(ws)], gs_ref)): tensors[f'grad_state[{i}]'] = ((gr / k).cpu().numpy(), ref.numpy())
            for i, (gr, ref) in enumerate(zip(grads[len(ws):], go_ref)): tensors[f'grad_output[{i}]'] = (gr.cpu().numpy(), ref.numpy())
            bn = gnn.net_state.layers[-1]
            tensors['moving_mean'] = (bn.moving_mean.cpu().numpy(), net_s.moving_mean.numpy())
            tensors['moving_var'] = (bn.moving_variance.cpu().numpy(), net_s.moving_var.numpy())
    finally:
        gnn.max_iteration = saved_iter
        gnn.net_state.set_weights(saved_w)          # the training-mode loop moved the BatchNormalization statistics
        torch.set_num_threads(1)
    errs = {name: _tensor_errors(a, b) for name, (a, b) in tensors.items()}
    worst = max(errs, key=lambda n: errs[n][0])
    k_equal = float(k) == float(k_ref)
    return {'ok': bool(k_equal and errs[worst][0] <= tol), 'k_equal': k_equal, 'k': float(k), 'tol': tol, 'iterations': iterations,
            'training': training, 'max_rel': errs[worst][0], 'worst_tensor': worst,
            'elementwise_rel_p99': max(e[1] for e in errs.values()), 'elementwise_rel_max': max(e[2] for e in errs.values()),
            'per_tensor_max_rel': {n: e[0] for n, e in errs.items()}, 'kernel': None}


# ---------------------------------------------------------------------------------------------------------------------
# configs 1-3 of BASELINE.json: small-graph training through the reference's own API (epoch time + steps/s)
# ---------------------------------------------------------------------------------------------------------------------
def make_small_dataset(kind: str, seed: int = 0):
    """ c1 / c3: starter.py:31-40 -- 100 random graphs, N ~ U{15..39}, NL 3, AL 1, T 2, density .7, node-focused, split
    .7 / .1 / .2 by getindices, training batches of 32 merged graphs.  c2: 4337 MUTAG-shaped graphs (the bundled dataset's size;
    NL 14 / AL 3 one-hot, T 2, graph-focused), one of the ten LKO folds: 9/10 of the graphs train, batches of 32. """
    from gnn_b200 import GNN_utils as utils
    np.random.seed(seed)
    if kind in ('c1', 'c3'):
        graphs = [utils.randomGraph(int(np.random.choice(range(15, 40))), 3, 1, 2, 0.7, problem_based='n') for _ in range(100)]
        iTr, iTe, iVa = utils.getindices(len(graphs), 0.7, 0.1, seed=seed)
        gTr = utils.getbatches([graphs[i] for i in iTr], 'n', 'average', batch_size=32)
        return gTr, 'n', 3, 1, 2
    graphs = []
    for b in make_graph_batches(4337, 1, seed=seed): graphs.append(b)
    fold = len(graphs) // 10
    gTr = utils.getbatches(graphs[fold:], 'g', 'average', batch_size=32)
    return gTr, 'g', 14, 3, 2


def build_small_model(kind: str, problem: str, NL: int, AL: int, T: int, device, path: str):
    """ nets per starter.py:52-69 (selu + lecun_normal state net with BatchNormalization and Dropout(0.1) at position 0; softmax +
    glorot_normal output net -- WITHOUT the trailing BatchNormalization: with T = 2 it makes the two outputs exact opposites and
    every loss NaN, in the reference too), state_vect_dim 0, max_iteration 5, threshold 0.01, Adam(1e-3); c3: 5 layers,
    get_state False / get_output True (starter.py:77-79) """
    from gnn_b200.MLP import MLP, get_inout_dims
    from gnn_b200.GNN import GNNnodeBased, GNNgraphBased
    from gnn_b200.LGNN import LGNN
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    cls = {'n': GNNnodeBased, 'g': GNNgraphBased}[problem]

    def one(layer, seed):
        kw = dict(layer=layer, get_state=False, get_output=True) if kind == 'c3' else dict()
        f_s, l_s = get_inout_dims('state', NL, AL, T, problem, 0, None, **kw)
        f_o, l_o = get_inout_dims('output', NL, AL, T, problem, 0, None, **kw)
        net_s = MLP(f_s, l_s, 'selu', 'lecun_normal', 'lecun_normal', dropout_rate=0.1, dropout_pos=0, batch_normalization=True, device=device, seed=seed)
        net_o = MLP(f_o, l_o, 'softmax', 'glorot_normal', 'glorot_normal', dropout_rate=0.1, dropout_pos=0, batch_normalization=False, device=device, seed=seed + 1)
        return cls(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=0, max_iteration=5, threshold=0.01,
                   addressed_problem='c', path_writer=f'{path}{layer}/')

    if kind != 'c3': return one(0, 10)
    return LGNN([one(l, 10 + 2 * l) for l in range(5)], False, True, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, 'c', path_writer=f'{path}lgnn/')


def cpu_port_small(kind, gTr, problem, weights):
    """ one epoch of forward + BPTT gradients (no optimizer) of the oracle port on the same batches, all host threads """
    import torch
    from oracle import gnn_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    graphs = []
    for g in gTr:
        ng = None
        if problem == 'g':
            gid, coeff, G = g.nodegraph_segments()
            ng = ('segments', gid, coeff, G)
        graphs.append(O.OracleGraph.build(g.arcs, g.nodes, g.targets, problem, g.set_mask, g.output_mask, g.sample_weights, ng, 'average', endpoints=(g._src, g._dst)))
    nets = [(O.OracleMLP.from_weights(ws, ['selu'], drop=[0.1, 0.0], batchnorm=True), O.OracleMLP.from_weights(wo, ['softmax'], drop=[0.1, 0.0])) for ws, wo in weights]
    t0 = time.perf_counter()
    for og in graphs:
        if kind == 'c3':
            O.lgnn_training_gradients(og, nets, O.categorical_crossentropy, training_mode='parallel', mean=True, get_state=False, get_output=True,
                                      state_vect_dim=0, max_iteration=5, threshold=0.01, x0s=[None] * 5, problem_based=problem)
        else:
            O.training_gradients(og, nets[0][0], nets[0][1], O.categorical_crossentropy, mean=True, state_vect_dim=0, max_iteration=5, threshold=0.01,
                                 problem_based=problem)
    return time.perf_counter() - t0, os.cpu_count() or 1


def bench_small(args, device):
    """ --workload c1 | c2 | c3: epoch time and training steps per second of gnn.train's inner loop (one training_step per merged
    batch of 32 graphs: forward loop, BPTT, Adam), eager launches and CUDA-graph replay, CPU port beside """
    import torch
    from gnn_b200 import _native
    from gnn_b200.graph_class import GraphTensor
    kind = args.workload
    gTr, problem, NL, AL, T = make_small_dataset(kind)
    gts = [GraphTensor.fromGraphObject(g, device=device) for g in gTr]
    arcs = sum(int(g.arcs.shape[0]) for g in gTr)
    modes = ['parallel', 'residual'] if kind == 'c3' else [None]     # serial = the single-GNN numbers, layer after layer
    out = {}
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    weights = None
    for mode in modes:
        for graphed in (False, True):
            model = build_small_model(kind, problem, NL, AL, T, device, f'/tmp/gnn_b200_bench_{kind}/')
            if mode: model.training_mode = mode
            if weights is None:
                gl = model.gnns if kind == 'c3' else [model]
                weights = [(m.net_state.get_weights(), m.net_output.get_weights()) for m in gl]
            model.use_cuda_graph = graphed
            ks = []

            def epoch():
                for gt in gts:
                    iters, _ = model.training_step(gt)
                    ks.append(iters)

            for _ in range(max(3, args.warmup)): epoch()        # (graphed: the eager first step and the capture happen here)
            torch.cuda.synchronize()
            _native.launch_count(reset=True)
            start.record()
            for _ in range(args.steps): epoch()
            stop.record()
            torch.cuda.synchronize()
            ms = start.elapsed_time(stop) / args.steps
            k_mean = float(np.mean([float(k) for it in ks[-len(gts):] for k in it]))
            out[(mode or 'gnn') + ('+cuda_graph' if graphed else '+eager')] = {
                'epoch_time_s': ms * 1e-3, 'steps_per_s': len(gts) / (ms * 1e-3), 'arc_updates_per_s': arcs * k_mean / (ms * 1e-3), 'mean_iterations': k_mean}
    cpu = None
    if not args.skip_cpu:
        sec, threads = cpu_port_small(kind, gTr, problem, weights)
        cpu = {'value': len(gTr) / sec, 'unit': 'training steps/s', 'epoch_time_s': sec, 'cores': threads, 'kind': 'port',
               'sample': 'one epoch of forward + BPTT gradients (no optimizer step) on the same batches',
               'note': 'torch-CPU restatement of the reference TF2 path (oracle/), not TensorFlow'}
    best = max(out, key=lambda name: out[name]['steps_per_s'])
    names = {'c1': 'config 1: starter.py default GNN, node-focused, 100 random graphs (70 train), batches of 32',
             'c2': 'config 2: MUTAG-shaped graph-focused GNN with NodeGraph pooling, one LKO fold of 4337 graphs (3904 train), batches of 32',
             'c3': 'config 3: 5-layer LGNN (get_output) on the config-1 data, parallel / residual modes'}
    print(json.dumps({'metric': 'training steps/s (one training_step per merged batch of 32 graphs); epoch time', 'value': out[best]['steps_per_s'],
                      'unit': 'training steps/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': out[best]['epoch_time_s'] * 1e3,
                      'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': names[kind], 'batches_per_epoch': len(gts), 'arcs_per_epoch': arcs, 'best': best,
                                 'l2': 'working set smaller than L2 by nature (a few thousand arcs per batch)'},
                      'variants': out, 'cpu_baseline': cpu}))

# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c4u', choices=['c4u', 'c4l', 'c5', 'c1', 'c2', 'c3'])
    ap.add_argument('--graphs-total', type=int, default=200_000)
    ap.add_argument('--graphs-per-step', type=int, default=25_000)
    ap.add_argument('--no-cuda-graph', action='store_true', help='graph batches: eager launches instead of one CUDA graph per batch')
    ap.add_argument('--nodes', type=int, default=1_000_000)
    ap.add_argument('--arcs', type=int, default=10_000_000)
    ap.add_argument('--max-iter', type=int, default=50)
    ap.add_argument('--cpu-iterations', type=int, default=40, help='iterations of the workload timed on the CPU port')
    ap.add_argument('--skip-train', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-parity', action='store_true', help='skip the oracle check of the timed configuration (3 iterations on the full graph)')
    ap.add_argument('--parity-iterations', type=int, default=3)
    ap.add_argument('--skip-variant', action='store_true', help='skip the side-by-side forward run on the other source distribution')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    metric = 'arc-updates/sec (arcs x iterations), state-convergence loop forward, 10M-arc graph'
    config = {'workload': f'{args.workload}: single synthetic graph, {args.nodes} nodes / {args.arcs} arcs, in-degree {args.arcs // args.nodes}, '
                          f'state_dim 32, NL 3, AL 1, max_iteration {args.max_iter}, threshold 0 (k = max_iteration), '
                          f'net_state Dense(71->32, selu)+BatchNormalization, aggregation average',
              'l2': 'inputs larger than L2 (state 2 x 128 MB + 40 MB arc indices per iteration)', 'seed': 0}

    # ---------------- reference arm: the CPU port of the reference path on the host cores -----------------------------
    if args.impl == 'reference':
        if rank != 0: return
        wl = make_workload(args.workload, args.nodes, args.arcs)
        vals = []
        for _ in range(max(1, args.warmup > 0) + args.steps):
            v, sec, threads = cpu_port_run(wl, args.cpu_iterations, repeats=1)
            vals.append((v, sec))
        vals = vals[1:] if len(vals) > 1 else vals
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([s for _, s in vals])) * 1e3
        sample = f'{args.cpu_iterations} iterations of the forward loop on the full {args.workload} graph per step (of {args.max_iter})'
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': value, 'unit': 'arc-updates/s', 'n_gpus': args.gpus,
                          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                          'cpu_baseline': {'value': value, 'unit': 'arc-updates/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                                           'note': 'torch-CPU restatement of the reference TF2 path (oracle/), not TensorFlow'},
                          'e2e': {'value': value, 'unit': 'arc-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    # ---------------- our arm -----------------------------------------------------------------------------------------
    import torch
    import gnn_b200
    from gnn_b200 import _native
    from gnn_b200.graph_class import GraphObject, GraphTensor
    from gnn_b200.GNN import GNNnodeBased
    from gnn_b200.keras_compat import Dense, BatchNormalization, Sequential, Adam, categorical_crossentropy
    if not torch.cuda.is_available(): raise SystemExit('bench.py needs a CUDA device')
    torch.set_num_threads(1)     # the host side only enqueues: intra-op CPU threads just add wake-up latency (the CPU port sets its own)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)

    from gnn_b200 import dist_graph
    if args.workload in ('c1', 'c2', 'c3'):      # small-graph training through the reference API (single GPU)
        if rank == 0: bench_small(args, device)
        if world > 1: dist.destroy_process_group()
        return
    if args.workload == 'c5':   # graph-batch sharding (weak scaling); not the headline workload
        with ClockSampler(local_rank) as clocks:
            res = bench_graph_batches(args, device, rank, world)
        if rank == 0:
            res.update({'metric': 'arc-updates/sec (arcs x iterations), forward+backward+optimizer, sharded graph batches', 'unit': 'arc-updates/s',
                        'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'higher_is_better': True, 'scaling': 'strong',
                        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'clocks': clocks.summary()})
            print(json.dumps(res))
        if world > 1: dist.destroy_process_group()
        return
    wl = make_workload(args.workload, args.nodes, args.arcs)
    N, E = wl['N'], wl['E']

    def build_gnn():
        net_s = Sequential([Dense(wl['DS'], activation='selu'), BatchNormalization()], input_dim=wl['AL'] + 2 * (wl['NL'] + wl['DS']), device=device)
        net_o = Sequential([Dense(wl['T'], activation='softmax')], input_dim=wl['NL'] + wl['DS'], device=device)
        net_s.set_weights(wl['ws']); net_o.set_weights(wl['wo'])
        gnn = GNNnodeBased(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=wl['DS'],
                           max_iteration=args.max_iter, threshold=0.0, addressed_problem='c', path_writer=f'/tmp/gnn_b200_bench_{rank}/')
        gnn.initial_state = torch.as_tensor(wl['x0'], device=device)
        return gnn

    g_host = GraphObject(arcs=wl['arcs'], nodes=wl['nodes'], targets=wl['targets'], problem_based='n', aggregation_mode='average',
                         _endpoints=(wl['src'], wl['dst']))
    if world > 1:
        progress(f'{args.workload}: partitioned forward loop + e2e')
        with ClockSampler(local_rank) as clocks:
            result = dist_graph.bench_partitioned(g_host, wl, build_gnn, args, device, rank, world)
        progress(f'{args.workload}: done, {result["ms_per_step"]:.2f} ms per loop')
        variant = None
        if not args.skip_variant and args.workload in ('c4u', 'c4l'):
            other = 'c4l' if args.workload == 'c4u' else 'c4u'
            del g_host
            wl2 = make_workload(other, args.nodes, args.arcs)
            g2 = GraphObject(arcs=wl2['arcs'], nodes=wl2['nodes'], targets=wl2['targets'], problem_based='n', aggregation_mode='average',
                             _endpoints=(wl2['src'], wl2['dst']))
            wl.update(x0=wl2['x0'], ws=wl2['ws'], wo=wl2['wo'])
            progress(f'{other}: partitioned forward loop')
            r2 = dist_graph.bench_partitioned(g2, wl2, build_gnn, args, device, rank, world, with_e2e=False)
            progress(f'{other}: done, {r2["ms_per_step"]:.2f} ms per loop')
            variant = {'workload': other + (': sources within +-2048 of the destination (boundary rows travel)' if other == 'c4l' else ': uniform sources'),
                       'value': r2['value'], 'ms_per_step': r2['ms_per_step'], 'iterations': r2['iterations'], 'partition': r2['partition']}
        batches = None
        if not args.skip_variant:
            r5 = bench_graph_batches(args, device, rank, world)
            batches = {'workload': r5['config']['workload'], 'scaling': 'strong', 'value': r5['value'], 'unit': 'arc-updates/s (forward+backward+Adam)',
                       'ms_per_step': r5['ms_per_step'], 'arcs_per_rank': r5['config']['arcs_per_rank'], 'roofline_frac': r5['roofline_frac'],
                       'gpu_launches': r5['gpu_launches']}
        if rank == 0:
            result['source_distribution_variant'] = variant
            result['graph_batches'] = batches
            config['partition'] = result.pop('partition')
            result.update({'metric': metric, 'unit': 'arc-updates/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
                           'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                           'config': config, 'cpu_baseline': None, 'clocks': clocks.summary(),
                           'roofline': {'bound': 'nvlink+hbm', 'note': 'per-iteration exchange of the state rows other ranks gather from: see '
                                        'config.partition.halo_bytes_received_per_iteration_per_rank; single-GPU kernel roofline in the N=1 line'}})
            print(json.dumps(result))
        dist.destroy_process_group()
        return

    gnn = build_gnn()
    gt = GraphTensor.fromGraphObject(g_host, device=device)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup):
        for _ in range(warmup): fn()
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps): fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / steps

    # --- parity of the timed configuration: same graph, same nets, same kernel plan, `parity-iterations` iterations ------
    parity = None
    if not args.skip_parity:
        parity = {args.workload: parity_check(wl, gnn, gt, device, args.parity_iterations, training=False)}
        parity[args.workload]['kernel'] = _native.last_forward_kernel()
        if not args.skip_train:
            parity[args.workload + '_train'] = parity_check(wl, gnn, gt, device, args.parity_iterations, training=True)
            parity[args.workload + '_train']['kernel'] = _native.last_forward_kernel()

    # --- value: forward loop, inputs resident ---------------------------------------------------------------------
    ks = []

    def fwd():
        with torch.no_grad():
            k, state, out = gnn.Loop(gt, training=False)
        ks.append(k)

    with ClockSampler(local_rank) as clocks:
        _native.launch_count(reset=True)
        ms_fwd = timed(fwd, args.steps, max(3, args.warmup))
        launches_fwd = _native.launch_count() // (args.steps + max(3, args.warmup))      # library kernels per step (one forward loop)
        k_fwd = float(ks[-1])
        value = E * k_fwd / (ms_fwd * 1e-3)

        # --- roofline: the fused iteration kernel alone (events inside the library, launching stream) --------------
        _native.profile_iterations(True)
        iter_ms = []
        for _ in range(args.steps + 1):
            fwd()
            ms, n_launch = _native.profile_last_iterations()
            iter_ms.append(ms / max(n_launch, 1))
        _native.profile_iterations(False)
    kernel_ms = float(np.mean(iter_ms[1:]))     # (the first profiled run also pays for the creation of the events)
    alg_bytes = algorithmic_bytes_per_iteration(N, E, wl['DS'], wl['NL'], wl['AL'])
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f: peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get('hbm_gbs', 6650.0))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': _native.last_forward_kernel(), 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                'frac': achieved / peak_gbs, 'traffic': None, 'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback 6650 GB/s',
                'algorithmic_bytes_per_launch': alg_bytes, 'bytes_per_arc_update': alg_bytes / E, 'ms_per_launch': kernel_ms}
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic_bytes_per_launch.json')
    if os.path.exists(traffic_file):     # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel (not measured in this run)
        with open(traffic_file) as f: roofline['traffic'] = json.load(f).get(args.workload)
        roofline['traffic_source'] = 'static: dram__bytes_read.sum + dram__bytes_write.sum of the ncu capture summarised in profiles/'

    # --- e2e: host buffers -> gnn(GraphObject) -> host outputs, every step -----------------------------------------
    g_host.pin_host_buffers()
    x0_host = torch.from_numpy(wl['x0']).pin_memory()
    h2d = g_host.host_bytes() + x0_host.numel() * 4
    d2h_holder = []

    def e2e_step():
        gnn.initial_state = x0_host.to(device, non_blocking=True)
        out = gnn(g_host)
        if not d2h_holder: d2h_holder.append(torch.empty(out.shape, dtype=out.dtype).pin_memory())   # page-locked result buffer
        d2h_holder[0].copy_(out, non_blocking=True)      # the timed region ends with a synchronize: the result is on the host

    ms_e2e = timed(e2e_step, args.steps, 1)
    e2e = {'value': E * k_fwd / (ms_e2e * 1e-3), 'unit': 'arc-updates/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': int(h2d),
           'd2h_bytes_per_step': int(d2h_holder[-1].numel() * 4)}
    gnn.initial_state = torch.as_tensor(wl['x0'], device=device)

    # --- train: forward + BPTT backward + optimizer (= one epoch of this one-graph dataset) -------------------------
    train = None
    if not args.skip_train:
        kt = []

        def train_step():
            iters, _ = gnn.training_step(gt, mean=True)
            kt.append(iters[0])

        _native.launch_count(reset=True)
        ms_train = timed(train_step, max(1, args.steps // 2), 1)
        launches_train = _native.launch_count() // (max(1, args.steps // 2) + 1)
        train = {'value': E * float(kt[-1]) / (ms_train * 1e-3), 'unit': 'arc-updates/s (forward+backward+Adam)', 'ms_per_step': ms_train,
                 'epoch_time_s': ms_train * 1e-3, 'k': float(kt[-1]), 'gpu_launches': int(launches_train)}

    # --- the other source distribution, forward loop only (U and L side by side) --------------------------
    variant = None
    if not args.skip_variant and args.workload in ('c4u', 'c4l'):
        other = 'c4l' if args.workload == 'c4u' else 'c4u'
        del gt, g_host
        torch.cuda.empty_cache()
        wl2 = make_workload(other, args.nodes, args.arcs)
        g2 = GraphObject(arcs=wl2['arcs'], nodes=wl2['nodes'], targets=wl2['targets'], problem_based='n', aggregation_mode='average',
                         _endpoints=(wl2['src'], wl2['dst']))
        gt2 = GraphTensor.fromGraphObject(g2, device=device)
        gnn.initial_state = torch.as_tensor(wl2['x0'], device=device)
        gnn.net_state.set_weights(wl2['ws']); gnn.net_output.set_weights(wl2['wo'])     # this workload's own nets (the train leg moved the others)
        if parity is not None:
            parity[other] = parity_check(wl2, gnn, gt2, device, args.parity_iterations, training=False)
            parity[other]['kernel'] = _native.last_forward_kernel()
            if not args.skip_train:
                parity[other + '_train'] = parity_check(wl2, gnn, gt2, device, args.parity_iterations, training=True)
                parity[other + '_train']['kernel'] = _native.last_forward_kernel()
        ks2 = []

        def fwd2():
            with torch.no_grad():
                k, state, out = gnn.Loop(gt2, training=False)
            ks2.append(k)

        ms2 = timed(fwd2, args.steps, 10)      # long warm-up: the GPU idled while the oracle ran the parity check on the host
        _native.profile_iterations(True)
        per_launch = []
        for _ in range(max(3, args.steps)):     # (the first profiled run also pays for the creation of the events)
            fwd2()
            ms_k2, n2 = _native.profile_last_iterations()
            per_launch.append(ms_k2 / max(n2, 1))
        _native.profile_iterations(False)
        ms_k2, n2 = float(np.mean(per_launch[1:])), 1
        variant = {'workload': other + (': sources within +-2048 of the destination' if other == 'c4l' else ': uniform sources'),
                   'value': wl2['E'] * float(ks2[-1]) / (ms2 * 1e-3), 'ms_per_step': ms2, 'iterations': float(ks2[-1]),
                   'ms_per_launch': ms_k2 / max(n2, 1), 'roofline_frac': alg_bytes / (ms_k2 / max(n2, 1) * 1e-3) / 1e9 / peak_gbs}
        del gt2, g2

    # --- sharded graph batches (C5, weak scaling: the same per-rank work at every N), training steps -----------------
    batches = None
    if not args.skip_variant:
        r5 = bench_graph_batches(args, device, rank, world)
        batches = {'workload': r5['config']['workload'], 'scaling': 'strong', 'value': r5['value'], 'unit': 'arc-updates/s (forward+backward+Adam)',
                   'ms_per_step': r5['ms_per_step'], 'arcs_per_rank': r5['config']['arcs_per_rank'], 'roofline_frac': r5['roofline_frac'],
                   'gpu_launches': r5['gpu_launches']}

    # --- CPU baseline: oracle port on the host cores, bounded sample ----------------------------------------------
    cpu = None
    if not args.skip_cpu:
        v, sec, threads = cpu_port_run(wl, args.cpu_iterations, repeats=1)
        cpu = {'value': v, 'unit': 'arc-updates/s', 'cores': threads, 'kind': 'port',
               'sample': f'{args.cpu_iterations} iterations of the forward loop on the full {args.workload} graph ({sec:.2f} s)',
               'note': 'torch-CPU restatement of the reference TF2 path (oracle/), not TensorFlow'}

    if parity is not None:
        failed = {name: p for name, p in parity.items() if not p['ok']}
        if failed:     # a fast kernel whose results differ from the reference's is not done: no value is printed
            print(json.dumps({'metric': metric, 'value': None, 'error': 'parity check of the timed configuration failed', 'parity': failed}))
            raise SystemExit(2)
    print(json.dumps({'metric': metric, 'parity': parity, 'value': value, 'unit': 'arc-updates/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': max(3, args.warmup),
                      'ms_per_step': ms_fwd, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
                      'data': 'synthetic', 'config': config, 'iterations': k_fwd, 'e2e': e2e, 'gpu_launches': int(launches_fwd),
                      'roofline': roofline, 'train': train, 'source_distribution_variant': variant, 'graph_batches': batches, 'cpu_baseline': cpu,
                      'clocks': clocks.summary()}))


if __name__ == '__main__':
    main()
