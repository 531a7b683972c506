# coding=utf-8
"""ORACLE (test infrastructure, NOT product code) -- floating-point part of the hot path.

CPU fp32 restatement (PyTorch-CPU tensors, autograd for the BPTT gradients) of the reference's state-convergence loop,
the Keras layers it calls, the output net, NodeGraph pooling, the loss and the training-step gradient scaling.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.

PARITY UNPINNED at the TensorFlow boundary: TensorFlow/Keras cannot be installed in the build container and the
reference ships no tests or golden vectors, so the Keras/TF op semantics below are restated from their documented
behaviour and cannot be checked against a TF run.  What IS pinned: the hand-worked loop-count
example of SURVEY 8c (tests/test_oracle.py) and every GraphObject-level structure (oracle/graph_oracle.py).

Reference lines followed (paths relative to /root/reference):
  GNN/GNN.py:202-220 condition, :223-242 convergence, :245-248 apply_filters, :251-280 Loop, :286-302 edge-based filters,
  :318-333 graph-based Loop, :180-199 evaluate_single_graph; GNN/GNN_BaseClass.py:231-247 training_step, :405-409
  get_filtered_tensor; GNN/MLP.py:11-64 layer order; GNN/LGNN.py:201-290.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import graph_oracle as G

# arithmetic type of the restatement: float32 as the reference; tests may switch to float64 to obtain a
# rounding-free reference value when a float32-vs-float32 difference has to be attributed
DTYPE = torch.float32


def set_dtype(dtype) -> None:
    global DTYPE
    DTYPE = dtype


SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805
_M32 = 0xFFFFFFFF


# =====================================================================================================================
# dropout generator (independent restatement of the product's counter-based generator; tests check they agree)
# =====================================================================================================================
def _fmix32(x):
    """ murmur3 finaliser on uint32 values held in uint64 arrays (products stay below 2**64) """
    x = x.astype(np.uint64) & np.uint64(_M32)
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(_M32)
    x = x ^ (x >> np.uint64(13))
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(_M32)
    return x ^ (x >> np.uint64(16))


def keep_mask(seed: int, stream: int, step: int, rows: int, cols: int, rate: float) -> np.ndarray:
    """ keep[n, j] = u >= rate, u = (h >> 8) / 2**24, h = fmix32(fmix32(lo ^ key) + hi*0xC2B2AE35 + 0x165667B1),
    idx = n*cols + j, key = fmix32(fmix32(seed + 0x9E3779B9*(stream+1)) ^ (step*0x85EBCA6B + 0x27D4EB2F)) """
    def fmix_int(x: int) -> int:
        x &= _M32
        x ^= x >> 16
        x = (x * 0x85EBCA6B) & _M32
        x ^= x >> 13
        x = (x * 0xC2B2AE35) & _M32
        return x ^ (x >> 16)

    key = fmix_int(fmix_int(seed + 0x9E3779B9 * (stream + 1)) ^ ((step * 0x85EBCA6B + 0x27D4EB2F) & _M32))
    idx = np.arange(rows, dtype=np.int64)[:, None] * cols + np.arange(cols, dtype=np.int64)[None, :]
    lo, hi = (idx & _M32).astype(np.uint64), ((idx >> 32) & _M32).astype(np.uint64)
    v = _fmix32(lo ^ np.uint64(key))
    v = _fmix32((v + hi * np.uint64(0xC2B2AE35) + np.uint64(0x165667B1)) & np.uint64(_M32))
    return ((v >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)) >= np.float32(rate)


# =====================================================================================================================
# Keras layers
# =====================================================================================================================
def activation(name: str, x: torch.Tensor) -> torch.Tensor:
    if name in ('linear', None): return x
    if name == 'relu': return torch.relu(x)
    if name == 'tanh': return torch.tanh(x)
    if name == 'sigmoid': return torch.sigmoid(x)
    if name == 'selu': return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * torch.expm1(x))
    if name == 'elu': return torch.where(x > 0, x, torch.expm1(x))
    if name == 'softmax': return torch.softmax(x, dim=-1)
    if name == 'softplus': return torch.nn.functional.softplus(x)
    raise ValueError(name)


@dataclass
class OracleMLP:
    """ what GNN/MLP.py:11-64 builds: Dense chain, Dropout in front of Dense l (drop[l]) or after the last one
    (drop[L]), optional trailing BatchNormalization (Keras defaults momentum .99, eps 1e-3) """
    W: list                      # kernel[in, out] per Dense
    b: list                      # bias[out]
    act: list                    # activation names
    drop: list = field(default_factory=list)   # L+1 rates (0 = none)
    gamma: Optional[torch.Tensor] = None
    beta: Optional[torch.Tensor] = None
    moving_mean: Optional[torch.Tensor] = None
    moving_var: Optional[torch.Tensor] = None
    eps: float = 1e-3
    momentum: float = 0.99

    @classmethod
    def from_weights(cls, weights: list, acts: list, drop=None, batchnorm: bool = False, requires_grad: bool = True):
        """ weights in Keras get_weights() order: [kernel, bias]*L (+ [gamma, beta, moving_mean, moving_var]) """
        t = lambda a, g: torch.tensor(np.asarray(a), dtype=DTYPE).requires_grad_(g)
        L = len(acts)
        W = [t(weights[2 * i], requires_grad) for i in range(L)]
        b = [t(weights[2 * i + 1], requires_grad) for i in range(L)]
        m = cls(W, b, list(acts), list(drop) if drop is not None else [0.0] * (L + 1))
        if batchnorm:
            m.gamma, m.beta = t(weights[2 * L], requires_grad), t(weights[2 * L + 1], requires_grad)
            m.moving_mean, m.moving_var = t(weights[2 * L + 2], False), t(weights[2 * L + 3], False)
        return m

    def trainable(self) -> list:
        """ Keras trainable_variables order """
        out = [v for pair in zip(self.W, self.b) for v in pair]
        if self.gamma is not None: out += [self.gamma, self.beta]
        return out

    def __call__(self, x, training: bool, seed: int = 0, stream_base: int = 0, step: int = 0):
        L = len(self.W)
        for l in range(L + 1):
            rate = self.drop[l] if self.drop else 0.0
            if training and rate > 0:                      # Dropout: x * keep / (1 - rate)
                keep = torch.from_numpy(keep_mask(seed, stream_base + l, step, x.shape[0], x.shape[1], rate))
                x = torch.where(keep, x * (1.0 / (1.0 - rate)), torch.zeros_like(x))
            if l < L:                                      # Dense: act(x @ kernel + bias)
                x = activation(self.act[l], x @ self.W[l] + self.b[l])
        if self.gamma is not None:                         # BatchNormalization
            if training:
                mean = x.mean(dim=0)
                var = ((x - mean) ** 2).mean(dim=0)        # biased batch variance
                with torch.no_grad():                      # moving <- moving*m + batch*(1-m), every call
                    self.moving_mean.mul_(self.momentum).add_(mean.detach() * (1 - self.momentum))
                    self.moving_var.mul_(self.momentum).add_(var.detach() * (1 - self.momentum))
            else:
                mean, var = self.moving_mean, self.moving_var
            x = (x - mean) * torch.rsqrt(var + self.eps) * self.gamma + self.beta
        return x


# =====================================================================================================================
# graph tensors (GNN/graph_class.py:330-372)
# =====================================================================================================================
@dataclass
class OracleGraph:
    nodes: torch.Tensor                # (N, NL)
    arcs: torch.Tensor                 # (E, 2 + AL), ids in columns 0-1 (float32, as the reference stores them)
    targets: torch.Tensor
    set_mask: torch.Tensor
    output_mask: torch.Tensor
    sample_weights: torch.Tensor
    adj: dict                          # transposed_row_major(Adjacency): rows = dst, cols = src
    arcnode: dict                      # transposed_row_major(ArcNode): rows = dst, cols = arc id
    nodegraph: Optional[torch.Tensor]  # dense (N, G), or sparse COO (N, G) for large merged batches, or None

    @classmethod
    def build(cls, arcs, nodes, targets, problem_based='n', set_mask=None, output_mask=None, sample_weights=1,
              nodegraph=None, aggregation_mode='average', endpoints=None):
        """ GraphObject.__init__ (graph_class.py:16-77) followed by GraphTensor.fromGraphObject (:354-361) """
        arcs, nodes, targets = np.asarray(arcs), np.asarray(nodes), np.asarray(targets)
        n_nodes, n_arcs = nodes.shape[0], arcs.shape[0]
        src, dst = endpoints if endpoints is not None else (arcs[:, 0].astype(int), arcs[:, 1].astype(int))
        mask_len = {'n': n_nodes, 'a': n_arcs, 'g': n_nodes}[problem_based]
        set_mask = np.ones(mask_len, bool) if set_mask is None else np.asarray(set_mask).astype(bool)
        output_mask = np.ones(mask_len, bool) if output_mask is None else np.asarray(output_mask).astype(bool)
        an_row, an_col, an_data = G.arcnode_coo(dst, n_nodes, aggregation_mode)
        ad_row, ad_col, ad_data = G.adjacency_coo(src, dst, an_data)
        if nodegraph is None: nodegraph = G.nodegraph(n_nodes, problem_based)
        f = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), dtype=DTYPE)
        if isinstance(nodegraph, tuple) and nodegraph[0] == 'segments':
            # block-diagonal NodeGraph of a merged batch (graph_class.py:313-315) given as (graph id, coefficient) per node:
            # the same matrix, held sparse -- the dense (N, G) array of 5 000 graphs x 150 000 nodes would take 3 GB
            _, gid, coeff, n_graphs = nodegraph
            idx = torch.stack([torch.arange(n_nodes), torch.as_tensor(np.asarray(gid, dtype=np.int64))])
            ng = torch.sparse_coo_tensor(idx, f(coeff), (n_nodes, int(n_graphs))).coalesce()
            nodegraph = None
        else:
            ng = None if nodegraph is None else f(nodegraph)
        return cls(nodes=f(nodes), arcs=f(arcs), targets=f(targets), set_mask=torch.tensor(set_mask),
                   output_mask=torch.tensor(output_mask), sample_weights=f(sample_weights * np.ones(targets.shape[0])),
                   adj=G.transposed_row_major(ad_row, ad_col, ad_data, (n_nodes, n_nodes)),
                   arcnode=G.transposed_row_major(an_row, an_col, an_data, (n_arcs, n_nodes)),
                   nodegraph=ng)


def spmm(sp: dict, dense: torch.Tensor, fast: bool = False) -> torch.Tensor:
    """ tf.sparse.sparse_dense_matmul(sp, dense): rows accumulate their entries in stored (ascending) order """
    rows = torch.from_numpy(sp['indices'][:, 0])
    cols = torch.from_numpy(sp['indices'][:, 1])
    vals = torch.from_numpy(sp['values']).to(dense.dtype)
    if fast:   # multi-threaded CSR kernel, used only for the timed CPU baseline
        if '_csr' not in sp:
            sp['_csr'] = torch.sparse_csr_tensor(torch.from_numpy(sp['rowptr']), cols, vals, size=sp['dense_shape'])
        return sp['_csr'] @ dense
    out = torch.zeros((sp['dense_shape'][0], dense.shape[1]), dtype=dense.dtype)
    return out.index_add(0, rows, vals[:, None] * dense[cols])


# =====================================================================================================================
# the loop
# =====================================================================================================================
def condition(k: float, state, state_old, threshold: float, max_iteration: int) -> bool:
    """ GNN/GNN.py:202-220 """
    out_distance = torch.sqrt(torch.sum(torch.square(state - state_old), dim=1))       # :206
    state_norm = torch.sqrt(torch.sum(torch.square(state_old), dim=1))                 # :209
    check = out_distance > threshold * state_norm                                       # :212-215
    return bool(torch.any(check)) and k < max_iteration                                 # :218-220


def loop(g: OracleGraph, net_state: OracleMLP, net_output: OracleMLP, *, state_vect_dim: int, max_iteration: int,
         threshold: float, training: bool = False, x0: Optional[torch.Tensor] = None, seed: int = 0,
         problem_based: str = 'n', own_state_gradient: bool = True, fast_spmm: bool = False, return_iterates: bool = False):
    """ GNN/GNN.py:251-280 (+ :318-333 for 'g', :286-302 for 'a').
    :param x0: injected initial state when state_vect_dim > 0 (the reference draws N(0, 0.1^2) unseeded, :262)
    :param own_state_gradient: tf.constant(eager_tensor) at :228 is an identity -> gradient flows (True)
    :return: (k float, state, out) """
    threshold = float(np.float32(threshold))
    with torch.no_grad():
        labels = g.arcs[:, 2:]
    aggregated_arcs = spmm(g.arcnode, labels, fast_spmm)                                   # :259
    n_nodes = g.nodes.shape[0]
    aggregated_nodes = torch.zeros((n_nodes, 0), dtype=DTYPE)                                         # :260
    if state_vect_dim > 0:
        state = x0.to(DTYPE) if x0 is not None else 0.1 * torch.randn(n_nodes, state_vect_dim, dtype=DTYPE)    # :262
        aggregated_nodes = spmm(g.adj, g.nodes, fast_spmm)                               # :263
    else:
        state = g.nodes                                                                  # :265
    state_old = torch.ones_like(state)                                                   # :266
    k = 0.0                                                                              # :267
    iterates = [state]

    while condition(k, state.detach(), state_old.detach(), threshold, max_iteration):    # :271 (eager while_loop)
        node_components = state if own_state_gradient else state.detach()               # :228
        if state_vect_dim: node_components = torch.cat([node_components, g.nodes], dim=1)   # :229-230
        aggregated_states = spmm(g.adj, state, fast_spmm)                                # :234
        inp_state = torch.cat([node_components, aggregated_states, aggregated_nodes, aggregated_arcs], dim=1)  # :237
        state_new = net_state(inp_state, training, seed=seed, stream_base=0, step=int(k))  # :240
        k, state, state_old = k + 1, state_new, state                                    # :242
        iterates.append(state)

    if problem_based == 'a':                                                             # :289-302
        conv = torch.cat([state, g.nodes], dim=1) if state_vect_dim else state
        idx = torch.from_numpy(g.adj['indices'])                                          # reordered (dst, src) pairs
        states = conv[idx].reshape(labels.shape[0], 2 * conv.shape[1])
        net_in = torch.cat([states, labels], dim=1)
    else:                                                                                # :245-248
        net_in = torch.cat([state, g.nodes], dim=1) if state_vect_dim else state
    mask = g.set_mask & g.output_mask                                                    # :275
    out = net_output(net_in[mask], training, seed=seed, stream_base=16, step=0)          # :279
    if problem_based == 'g':                                                             # :331-332
        out = g.nodegraph.t() @ out
    if return_iterates: return k, state, out, iterates
    return k, state, out


def categorical_crossentropy(y_true, y_pred):
    """ tf.keras.losses.categorical_crossentropy(from_logits=False): renormalise, clip [1e-7, 1-1e-7], -sum t log p """
    y_pred = y_pred / y_pred.sum(dim=-1, keepdim=True)
    y_pred = torch.clamp(y_pred, 1e-7, 1 - 1e-7)
    return -(y_true * torch.log(y_pred)).sum(dim=-1)


def mean_squared_error(y_true, y_pred):
    return ((y_pred - y_true) ** 2).mean(dim=-1)


def filtered(g: OracleGraph, inp: torch.Tensor, problem_based: str):
    """ GNN_BaseClass.py:405-409 (node/arc) and GNN.py:313-315 (graph) """
    if problem_based == 'g': return inp
    return inp[g.set_mask[g.output_mask]]


def evaluate_single_graph(g, net_state, net_output, loss_fn, *, problem_based='n', **loop_args):
    """ GNN/GNN.py:180-199: loss = sum_i loss_fn(t_i, o_i) * w_i """
    targs = filtered(g, g.targets, problem_based)
    weights = filtered(g, g.sample_weights, problem_based)
    k, state, out = loop(g, net_state, net_output, problem_based=problem_based, **loop_args)
    loss = loss_fn(targs, out) * weights
    return k, loss.sum(), targs, out, state


def training_gradients(g, net_state, net_output, loss_fn, *, mean: bool = True, **kw):
    """ GNN_BaseClass.py:231-247 without the optimizer: BPTT gradients of the summed loss wrt
    [net_state vars], [net_output vars]; net_state gradients divided by k when mean (:241) """
    k, loss, _, out, state = evaluate_single_graph(g, net_state, net_output, loss_fn, training=True, **kw)
    ws, wo = net_state.trainable(), net_output.trainable()
    grads = torch.autograd.grad(loss, ws + wo, allow_unused=True)
    gs, go = list(grads[:len(ws)]), list(grads[len(ws):])
    gs = [torch.zeros_like(w) if gr is None else gr for w, gr in zip(ws, gs)]
    if mean: gs = [gr / k for gr in gs]
    return k, loss.detach(), gs, go, out.detach(), state.detach()


# =====================================================================================================================
# LGNN (GNN/LGNN.py:201-290)
# =====================================================================================================================
def lgnn_update_graph(g: OracleGraph, state, output, get_state: bool, get_output: bool, problem_based: str) -> OracleGraph:
    """ GNN/LGNN.py:227-260: new graph whose node (arc) labels are the ORIGINAL labels ++ state ++ scatter_nd(where(mask),
    output) -- zeros where the mask is off """
    import copy
    new = copy.copy(g)
    extra_nodes, extra_arcs = [], []
    if get_state: extra_nodes.append(state)
    if get_output:
        mask = g.set_mask & g.output_mask                                                  # :247
        rows = g.arcs.shape[0] if problem_based == 'a' else g.nodes.shape[0]
        out = torch.zeros((rows, output.shape[1]), dtype=output.dtype).index_put((torch.nonzero(mask)[:, 0],), output)  # :251
        (extra_arcs if problem_based == 'a' else extra_nodes).append(out)                   # :253-256
    if extra_nodes: new.nodes = torch.cat([g.nodes] + extra_nodes, dim=1)                   # :258
    if extra_arcs: new.arcs = torch.cat([g.arcs] + extra_arcs, dim=1)                       # :259
    return new


def lgnn_loop(g: OracleGraph, nets: list, *, get_state: bool, get_output: bool, state_vect_dim: int, max_iteration: int,
              threshold: float, training: bool, x0s=None, seeds=None, problem_based: str = 'n'):
    """ GNN/LGNN.py:263-290: every layer runs the node-level Loop on the updated graph; for graph-based problems the
    pooled output of the non-final layers only enters the output list (:276-278).
    :param nets: list of (net_state, net_output) per layer; x0s / seeds: per-layer injected initial states / dropout seeds
    :return: (list of k, state of the last layer, list of outputs) """
    L = len(nets)
    gtmp, K, outs = g, [], []
    state = out = None
    for idx, (net_s, net_o) in enumerate(nets):
        last = idx == L - 1
        node_level = 'n' if (problem_based == 'g' and not last) else problem_based
        x0 = None if x0s is None else x0s[idx]
        seed = 0 if seeds is None else seeds[idx]
        k, state, out = loop(gtmp, net_s, net_o, state_vect_dim=state_vect_dim, max_iteration=max_iteration, threshold=threshold,
                             training=training, x0=x0, seed=seed, problem_based=node_level)
        K.append(k)
        if problem_based == 'g' and not last:
            outs.append(g.nodegraph.t() @ out)                                               # :278
            gtmp = lgnn_update_graph(g, state, out, get_state, get_output, 'n')
        else:
            outs.append(out)
            if not last: gtmp = lgnn_update_graph(g, state, out, get_state, get_output, problem_based)
    return K, state, outs


def lgnn_training_gradients(g, nets, loss_fn, *, training_mode: str, mean: bool = True, problem_based='n', **kw):
    """ GNN/LGNN.py:201-224 + GNN_BaseClass.py:231-247: parallel -> sum_targets mean_i(loss(t, o_i) * w);
    residual -> loss(t, mean_i o_i) * w; net_state gradients of layer i divided by its own k_i """
    targs = filtered(g, g.targets, problem_based)
    weights = filtered(g, g.sample_weights, problem_based)
    K, state, outs = lgnn_loop(g, nets, training=True, problem_based=problem_based, **kw)
    if training_mode == 'residual':
        loss = (loss_fn(targs, torch.stack(outs, 0).mean(0)) * weights).sum()
    else:
        loss = torch.stack([loss_fn(targs, o) * weights for o in outs], 0).mean(0).sum()
    gs, go = [], []
    params = [p for net_s, net_o in nets for p in net_s.trainable() + net_o.trainable()]
    grads = list(torch.autograd.grad(loss, params, allow_unused=True))
    pos = 0
    for (net_s, net_o), k in zip(nets, K):
        ns, no = len(net_s.trainable()), len(net_o.trainable())
        g_s = [torch.zeros_like(w) if gr is None else gr for w, gr in zip(net_s.trainable(), grads[pos:pos + ns])]
        g_o = [torch.zeros_like(w) if gr is None else gr for w, gr in zip(net_o.trainable(), grads[pos + ns:pos + ns + no])]
        gs.append([gr / k if mean else gr for gr in g_s])
        go.append(g_o)
        pos += ns + no
    return K, loss.detach(), gs, go, [o.detach() for o in outs]
