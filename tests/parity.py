"""Shared helpers of the parity tests (and of __graft_entry__.smoke): build one seeded case, run it through the CUDA
product path (gnn_b200, C ABI) and through the CPU oracle (oracle/), compare.

Tolerance (north_star): iteration count EXACT; states, outputs, loss and gradients within 1e-4 relative in fp32,
measured per tensor as max|got - want| <= tol * max(1e-6 + max|want|).
"""
from __future__ import annotations

import numpy as np
import torch

TOL = 1e-4


def _rng_weights(rng, dims, scale=None):
    ws = []
    for i, o in zip(dims[:-1], dims[1:]):
        s = (1.0 / np.sqrt(i)) if scale is None else scale
        ws += [(rng.standard_normal((i, o)) * s).astype(np.float32), (rng.standard_normal(o) * 0.1).astype(np.float32)]
    return ws


def random_case(seed=0, n_nodes=200, n_arcs=1500, NL=3, AL=2, DS=0, hidden=(), act='tanh', max_iter=10, threshold=0.01,
                aggregation='average', bn=False, drop=None, problem='n', T=2, out_hidden=(), out_act='softmax', out_bn=False,
                out_drop=None, n_graphs=1, isolated=True, duplicates=True, masks=True, custom_arcnode=False, weight_scale=None):
    """ one seeded problem instance: graph arrays, net weights in Keras get_weights() order, injected initial state """
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_nodes, n_arcs)
    dst = rng.integers(0, n_nodes, n_arcs)
    if isolated and n_nodes > 4:      # nodes 0 and 1 receive no arc
        dst = np.where(dst < 2, 2 + dst % (n_nodes - 2), dst)
    if duplicates and n_arcs > 4:     # duplicate arcs stay duplicate
        src[1], dst[1] = src[0], dst[0]
    if n_graphs > 1:                  # disjoint union: keep arcs inside their graph
        bounds = np.linspace(0, n_nodes, n_graphs + 1).astype(int)
        gid = np.searchsorted(bounds, dst, side='right') - 1
        lo, hi = bounds[gid], bounds[gid + 1]
        src = lo + (src % np.maximum(hi - lo, 1))
    arcs = np.concatenate([np.stack([src, dst], axis=1).astype(np.float64), rng.uniform(-1, 1, (n_arcs, AL))], axis=1)
    nodes = rng.uniform(-1, 1, (n_nodes, NL))
    D = DS if DS else NL
    n_targets = {'n': n_nodes, 'a': n_arcs, 'g': n_graphs}[problem]
    mask_len = n_arcs if problem == 'a' else n_nodes
    set_mask = output_mask = None
    if masks and problem != 'g':
        set_mask = rng.random(mask_len) < 0.8
        output_mask = rng.random(mask_len) < 0.7
        output_mask[:3] = True
        set_mask[:3] = True
        n_targets = int(output_mask.sum())
    targets = np.eye(T)[rng.integers(0, T, n_targets)]
    sample_weights = rng.uniform(0.5, 2.0, n_targets)
    nodegraph = None
    if problem == 'g':
        bounds = np.linspace(0, n_nodes, n_graphs + 1).astype(int)
        nodegraph = np.zeros((n_nodes, n_graphs), dtype=np.float32)
        for gidx in range(n_graphs): nodegraph[bounds[gidx]:bounds[gidx + 1], gidx] = 1.0 / max(bounds[gidx + 1] - bounds[gidx], 1)
    arcnode = None
    if custom_arcnode:                # arbitrary per-arc weights: exercises the per-arc value path of the kernels
        from scipy.sparse import coo_matrix
        arcnode = coo_matrix((rng.uniform(0.1, 1.0, n_arcs), (np.arange(n_arcs), dst)), shape=(n_arcs, n_nodes))

    F_state = AL + 2 * (NL + DS)
    F_out = (NL + DS) * (2 if problem == 'a' else 1) + (AL if problem == 'a' else 0)
    state_dims = [F_state] + list(hidden) + [D]
    out_dims = [F_out] + list(out_hidden) + [T]
    ws = _rng_weights(rng, state_dims, weight_scale)
    wo = _rng_weights(rng, out_dims)
    if bn:
        ws += [rng.uniform(0.5, 1.5, D).astype(np.float32), rng.uniform(-0.2, 0.2, D).astype(np.float32),
               rng.uniform(-0.1, 0.1, D).astype(np.float32), rng.uniform(0.5, 1.5, D).astype(np.float32)]
    if out_bn:
        wo += [rng.uniform(0.5, 1.5, T).astype(np.float32), rng.uniform(-0.2, 0.2, T).astype(np.float32),
               rng.uniform(-0.1, 0.1, T).astype(np.float32), rng.uniform(0.5, 1.5, T).astype(np.float32)]
    acts_state = [act] * len(state_dims[1:])
    acts_out = ['tanh'] * len(out_hidden) + [out_act]
    x0 = (0.1 * rng.standard_normal((n_nodes, DS))).astype(np.float32) if DS else None
    return dict(arcs=arcs, nodes=nodes, targets=targets, set_mask=set_mask, output_mask=output_mask, sample_weights=sample_weights,
                nodegraph=nodegraph, arcnode=arcnode, aggregation=aggregation, problem=problem, DS=DS, D=D, NL=NL, AL=AL, T=T,
                state_dims=state_dims, out_dims=out_dims, ws=ws, wo=wo, acts_state=acts_state, acts_out=acts_out, bn=bn, out_bn=out_bn,
                drop=list(drop) if drop is not None else [0.0] * (len(acts_state) + 1),
                out_drop=list(out_drop) if out_drop is not None else [0.0] * (len(acts_out) + 1),
                x0=x0, max_iter=max_iter, threshold=threshold, seed=1234 + seed)


# ---------------------------------------------------------------------------------------------------------------------
def _build_sequential(dims, acts, drop, bn, weights, device):
    from gnn_b200.keras_compat import Dense, Dropout, BatchNormalization, Sequential
    layers = []
    for l, (units, a) in enumerate(zip(dims[1:], acts)):
        if drop[l] > 0: layers.append(Dropout(drop[l]))
        layers.append(Dense(units, activation=a))
    if drop[len(acts)] > 0: layers.append(Dropout(drop[len(acts)]))
    if bn: layers.append(BatchNormalization())
    net = Sequential(layers, input_dim=dims[0], device=device)
    net.set_weights(weights)
    return net


def build_product(case, device='cuda'):
    """ GraphTensor + GNN of the product package for a case """
    import gnn_b200
    from gnn_b200.graph_class import GraphObject, GraphTensor
    from gnn_b200.GNN import GNNnodeBased, GNNedgeBased, GNNgraphBased
    from gnn_b200.keras_compat import Adam, categorical_crossentropy
    g = GraphObject(arcs=case['arcs'], nodes=case['nodes'], targets=case['targets'], problem_based=case['problem'],
                    set_mask=case['set_mask'], output_mask=case['output_mask'], sample_weights=case['sample_weights'],
                    NodeGraph=case['nodegraph'], ArcNode=case['arcnode'], aggregation_mode=case['aggregation'])
    gt = GraphTensor.fromGraphObject(g, device=device)
    net_s = _build_sequential(case['state_dims'], case['acts_state'], case['drop'], case['bn'], case['ws'], device)
    net_o = _build_sequential(case['out_dims'], case['acts_out'], case['out_drop'], case['out_bn'], case['wo'], device)
    cls = {'n': GNNnodeBased, 'a': GNNedgeBased, 'g': GNNgraphBased}[case['problem']]
    gnn = cls(net_s, net_o, Adam(1e-3), categorical_crossentropy, {'from_logits': False}, state_vect_dim=case['DS'],
              max_iteration=case['max_iter'], threshold=case['threshold'], addressed_problem='c', path_writer='/tmp/gnn_b200_writer/')
    if case['x0'] is not None: gnn.initial_state = torch.as_tensor(case['x0'], device=device)
    return g, gt, gnn


def run_cuda(case, training=False, mean=True):
    """ product path: Loop (+ BPTT gradients when training) through the C ABI on the current CUDA device """
    g, gt, gnn = build_product(case)
    seed = case['seed']
    res = dict()
    if not training:
        with torch.no_grad():
            k, state, out = gnn.Loop(gt, training=False, seed=seed)
        res.update(k=float(k), state=state.cpu().numpy(), out=out.cpu().numpy())
        return res
    targs = gnn.get_filtered_tensor(gt, gt.targets)
    weights = gnn.get_filtered_tensor(gt, gt.sample_weights)
    k, state, out = gnn.Loop(gt, training=True, seed=seed)
    loss = (gnn.loss_function(targs, out, **gnn.loss_args) * weights).sum()
    ws, wo = gnn.net_state.trainable_variables, gnn.net_output.trainable_variables
    grads = torch.autograd.grad(loss, ws + wo, allow_unused=True)
    grads = [torch.zeros_like(v) if gr is None else gr for v, gr in zip(ws + wo, grads)]
    gs = [(gr / k if mean else gr).cpu().numpy() for gr in grads[:len(ws)]]
    go = [gr.cpu().numpy() for gr in grads[len(ws):]]
    res.update(k=float(k), state=state.detach().cpu().numpy(), out=out.detach().cpu().numpy(), loss=float(loss), gs=gs, go=go)
    if case['bn']:
        bn = gnn.net_state.layers[-1]
        res.update(moving_mean=bn.moving_mean.cpu().numpy(), moving_var=bn.moving_variance.cpu().numpy())
    return res


def run_oracle(case, training=False, mean=True, float64=False):
    """ CPU oracle on the same inputs (float64=True: same restatement evaluated in double precision) """
    from oracle import gnn_oracle as O
    O.set_dtype(torch.float64 if float64 else torch.float32)
    try:
        return _run_oracle(O, case, training, mean)
    finally:
        O.set_dtype(torch.float32)


def _run_oracle(O, case, training, mean):
    src, dst = case['arcs'][:, 0].astype(int), case['arcs'][:, 1].astype(int)
    g = O.OracleGraph.build(case['arcs'], case['nodes'], case['targets'], case['problem'], case['set_mask'], case['output_mask'],
                            case['sample_weights'], case['nodegraph'], case['aggregation'], endpoints=(src, dst))
    if case['arcnode'] is not None:   # custom ArcNode values
        from oracle import graph_oracle as G
        data = np.asarray(case['arcnode'].data, dtype=np.float32)
        n_nodes, n_arcs = case['nodes'].shape[0], case['arcs'].shape[0]
        g.arcnode = G.transposed_row_major(np.arange(n_arcs), dst, data, (n_arcs, n_nodes))
        g.adj = G.transposed_row_major(src, dst, data, (n_nodes, n_nodes))
    net_s = O.OracleMLP.from_weights(case['ws'], case['acts_state'], case['drop'], case['bn'])
    net_o = O.OracleMLP.from_weights(case['wo'], case['acts_out'], case['out_drop'], case['out_bn'])
    x0 = None if case['x0'] is None else torch.tensor(case['x0'])
    kw = dict(state_vect_dim=case['DS'], max_iteration=case['max_iter'], threshold=case['threshold'], x0=x0, seed=case['seed'],
              problem_based=case['problem'])
    if not training:
        with torch.no_grad():
            k, state, out = O.loop(g, net_s, net_o, training=False, **kw)
        return dict(k=float(k), state=state.numpy(), out=out.numpy())
    k, loss, gs, go, out, state = O.training_gradients(g, net_s, net_o, O.categorical_crossentropy, mean=mean, **kw)
    res = dict(k=float(k), state=state.numpy(), out=out.numpy(), loss=float(loss), gs=[t.numpy() for t in gs],
               go=[torch.zeros_like(w).numpy() if t is None else t.numpy() for w, t in zip(net_o.trainable(), go)])
    if case['bn']: res.update(moving_mean=net_s.moving_mean.numpy(), moving_var=net_s.moving_var.numpy())
    return res


def rel_err(got, want) -> float:
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if got.shape != want.shape: return float('inf')
    if want.size == 0: return 0.0
    return float(np.max(np.abs(got - want)) / (1e-6 + np.max(np.abs(want))))


def _errors(got: dict, want: dict) -> dict:
    errs = dict()
    for key in want:
        if key == 'k': continue
        if isinstance(want[key], list):
            for i, (a, b) in enumerate(zip(got[key], want[key])): errs[f'{key}[{i}]'] = rel_err(a, b)
        else:
            errs[key] = rel_err(got[key], want[key])
    return errs


def assert_parity(got: dict, want: dict, tol: float = TOL, want64=None):
    """ every tensor within tol of the float32 oracle.  want64 (callable returning the SAME oracle evaluated in float64):
    consulted only for tensors that miss the float32 oracle -- a float32 run carries its own rounding noise, so a tensor
    also passes when it is within tol of the rounding-free value AND at least as close to it as the float32 oracle is """
    assert got['k'] == want['k'], f"iteration count differs: got {got['k']} want {want['k']}"
    errs = _errors(got, want)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    if bad and want64 is not None:
        exact = want64()
        assert got['k'] == exact['k']
        e_got, e_ref = _errors(got, exact), _errors(want, exact)
        bad = {k: (v, e_got[k], e_ref[k]) for k, v in bad.items() if not (e_got[k] <= tol and e_got[k] <= 2 * e_ref[k] + 1e-7)}
    assert not bad, f'parity failures (tol {tol}): {bad}; all errors: {errs}'
    return errs
