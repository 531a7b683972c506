# coding=utf-8
"""GNN models behind the reference's API (``GNN/GNN.py``): node-, edge- and graph-based Scarselli GNNs.

``Loop`` keeps the reference's structure (GNN/GNN.py:251-280) but the whole ``tf.while_loop`` -- aggregation of the
neighbour states, ``net_state``, the convergence test -- is one call into the CUDA library (``state_loop``), forward
and backward.  ``net_output`` (a few rows x a few units) is evaluated with torch ops on the same device.
"""
from __future__ import annotations

import json
import os
from typing import Optional, Union

import numpy as np
import torch

from .GNN_BaseClass import BaseClass
from .graph_class import GraphObject, GraphTensor
from . import _native
from .keras_compat import Sequential, Dense, Dropout, losses as _losses, optimizers as _optimizers
from .state_loop import state_loop, sparse_dense


#######################################################################################################################
### CLASS GNN - NODE BASED ############################################################################################
#######################################################################################################################
class GNNnodeBased(BaseClass):
    """ GNN for node-based problem """

    ## CONSTRUCTORS METHODS ###########################################################################################
    def __init__(self,
                 net_state: Sequential,
                 net_output: Sequential,
                 optimizer,
                 loss_function,
                 loss_arguments: Optional[dict],
                 state_vect_dim: int,
                 max_iteration: int,
                 threshold: float,
                 addressed_problem: str,
                 extra_metrics: Optional[dict] = None,
                 extra_metrics_arguments: Optional[dict[str, dict]] = None,
                 path_writer: str = 'writer/',
                 namespace: str = 'GNN') -> None:
        """ CONSTRUCTOR (GNN.py:22-64)

        :param net_state: (Sequential) MLP for the state network, built by MLP().
        :param net_output: (Sequential) MLP for the output network, built by MLP().
        :param optimizer: (keras_compat.optimizers) for gradient application.
        :param loss_function: (keras_compat.losses) loss, called as loss(targets, outputs, **loss_arguments).
        :param loss_arguments: (dict) extra arguments of the loss.
        :param state_vect_dim: (int)>=0, state width for a GNN which does not initialize states with node labels.
        :param max_iteration: (int) max number of iteration for the unfolding procedure (to reach convergence).
        :param threshold: threshold for specifying if convergence is reached or not.
        :param addressed_problem: (str) in ['r','c'], 'r':regression, 'c':classification.
        :param extra_metrics: None or dict {'name':function} for metrics watched during training/validation/test.
        :param extra_metrics_arguments: None or dict {'name': {'argument': value}}.
        :param path_writer: (str) path for TensorBoard files. If the folder is not empty, all files are removed.
        :param namespace: (str) namespace for tensorboard visualization.
        """
        if not isinstance(state_vect_dim, int) or state_vect_dim < 0: raise TypeError('param <state_vect_dim> must be int>=0')
        super().__init__(optimizer, loss_function, loss_arguments, addressed_problem, extra_metrics, extra_metrics_arguments,
                         path_writer, namespace)
        self.net_state = net_state
        self.net_output = net_output
        self.max_iteration = max_iteration
        self.state_threshold = threshold
        self.state_vect_dim = state_vect_dim
        # reproducibility hooks (the reference draws unseeded): initial state override and dropout seed
        self.initial_state: Optional[torch.Tensor] = None
        self.dropout_seed: int = 0x5EED
        self._calls = 0

    # -----------------------------------------------------------------------------------------------------------------
    def copy(self, *, path_writer: str = '', namespace: str = '', copy_weights: bool = True):
        """ deep copy of the GNN (GNN.py:67-90); copy_weights False re-initialises both networks """
        if not path_writer: path_writer = self.path_writer + '_copied/'
        if not namespace: namespace = 'GNN'
        netS, netO = self.net_state.clone(), self.net_output.clone()
        if copy_weights:
            netS.set_weights(self.net_state.get_weights())
            netO.set_weights(self.net_output.get_weights())
        return self.__class__(net_state=netS, net_output=netO, optimizer=self.optimizer.__class__(**self.optimizer.get_config()),
                              loss_function=self.loss_function, loss_arguments=self.loss_args, max_iteration=self.max_iteration,
                              threshold=self.state_threshold, addressed_problem=self.addressed_problem,
                              extra_metrics=self.extra_metrics, extra_metrics_arguments=self.mt_args,
                              state_vect_dim=self.state_vect_dim, path_writer=path_writer, namespace=namespace)

    ## SAVE AND LOAD METHODs ##########################################################################################
    @staticmethod
    def _save_net(net: Sequential, folder: str) -> None:
        os.makedirs(folder, exist_ok=True)
        arch = []
        for l in net.layers:
            cfg = l.config()
            unsaved = [k for k, v in cfg.items() if callable(v)]
            # the reference's Keras save keeps regularizers / initializers by config; a Python callable has none: refuse rather than
            # silently write a model that loads without its regularisation
            if unsaved: raise ValueError(f'cannot save layer {type(l).__name__}: {unsaved} are Python callables (pass them by name, or None)')
            arch.append({'class': type(l).__name__, 'config': cfg})
        with open(os.path.join(folder, 'architecture.json'), 'w') as f:
            json.dump({'input_dim': net.input_dim, 'layers': arch}, f)
        np.savez(os.path.join(folder, 'weights.npz'), *net.get_weights())

    @staticmethod
    def _load_net(folder: str) -> Sequential:
        from . import keras_compat as K
        with open(os.path.join(folder, 'architecture.json')) as f: arch = json.load(f)
        net = Sequential([getattr(K, l['class'])(**l['config']) for l in arch['layers']], input_dim=arch['input_dim'])
        data = np.load(os.path.join(folder, 'weights.npz'))
        net.set_weights([data[f'arr_{i}'] for i in range(len(data.files))])
        return net

    def save(self, path: str):
        """ save model to folder <path> (GNN.py:93-111): net_state/, net_output/, config.json with the reference's keys """
        if path[-1] != '/': path += '/'
        self._save_net(self.net_state, f'{path}net_state/')
        self._save_net(self.net_output, f'{path}net_output/')
        config = {'loss_function': _losses.serialize(self.loss_function), 'loss_arguments': self.loss_args,
                  'optimizer': _optimizers.serialize(self.optimizer),
                  'max_iteration': self.max_iteration, 'threshold': self.state_threshold,
                  'addressed_problem': self.addressed_problem, 'state_vect_dim': self.state_vect_dim}
        with open(f'{path}config.json', 'w') as json_file:
            json.dump(config, json_file)

    @classmethod
    def load(cls, path: str, path_writer: Optional[str] = None, namespace: str = 'GNN',
             extra_metrics: Optional[dict] = None, extra_metrics_arguments: Optional[dict[str, dict]] = None):
        """ load model from folder <path> (GNN.py:115-149); no eval() of stored strings """
        if path[-1] != '/': path += '/'
        if path_writer is None: path_writer = f'{path}writer'
        with open(f'{path}config.json', 'r') as read_file:
            config = json.loads(read_file.read())
        optz = _optimizers.deserialize(config.pop('optimizer'))
        loss = _losses.deserialize(config.pop('loss_function'))
        netS, netO = cls._load_net(f'{path}net_state/'), cls._load_net(f'{path}net_output/')
        return cls(net_state=netS, net_output=netO, optimizer=optz, loss_function=loss, extra_metrics=extra_metrics,
                   extra_metrics_arguments=extra_metrics_arguments, path_writer=path_writer, namespace=namespace, **config)

    ## GETTERS AND SETTERS METHODs ####################################################################################
    def get_dense_layers(self) -> list:
        """ Dense layers of both nets, for the regularizers applied at training time (GNN.py:152-156) """
        return [l for net in (self.net_state, self.net_output) for l in net.layers if isinstance(l, Dense)]

    def trainable_variables(self) -> tuple[list[list[torch.Tensor]], list[list[torch.Tensor]]]:
        return [self.net_state.trainable_variables], [self.net_output.trainable_variables]

    def get_weights(self) -> tuple[list[list[np.ndarray]], list[list[np.ndarray]]]:
        return [self.net_state.get_weights()], [self.net_output.get_weights()]

    def set_weights(self, weights_state: list[list[np.ndarray]], weights_output: list[list[np.ndarray]]) -> None:
        assert len(weights_state) == len(weights_output) == 1
        self.net_state.set_weights(weights_state[0])
        self.net_output.set_weights(weights_output[0])

    def to(self, device):
        self.net_state.to(device)
        self.net_output.to(device)
        return self

    ## CALL/PREDICT METHOD ############################################################################################
    def __call__(self, g: Union[GraphObject, GraphTensor]) -> torch.Tensor:
        """ return ONLY the GNN output in test mode (training == False) for graph g """
        with torch.no_grad():
            return self.Loop(g, training=False)[-1]

    ## EVALUATE METHODS ###############################################################################################
    def evaluate_single_graph(self, g: Union[GraphObject, GraphTensor], training: bool) -> tuple:
        """ (iterations, summed loss, targets, outputs) of one graph (GNN.py:180-199); loss = SUM_i loss_i * weight_i """
        if isinstance(g, GraphObject): g = GraphTensor.fromGraphObject(g)
        targs = self.get_filtered_tensor(g, g.targets)
        loss_weights = self.get_filtered_tensor(g, g.sample_weights)
        it, _, out = self.Loop(g, training=training)
        loss = self.loss_function(targs, out, **self.loss_args) * loss_weights
        return it, loss.sum(), targs, out

    ## LOOP METHODS ###################################################################################################
    def condition(self, k, state, state_old, *args) -> torch.Tensor:
        """ the loop condition (GNN.py:202-220) as torch ops; the CUDA loop evaluates the same predicate on the device """
        out_distance = torch.sqrt(torch.sum(torch.square(state - state_old), dim=1))
        state_norm = torch.sqrt(torch.sum(torch.square(state_old), dim=1))
        moving = torch.any(out_distance > self.state_threshold * state_norm)
        return torch.logical_and(moving, torch.as_tensor(k < self.max_iteration, device=state.device))

    def convergence(self, k, state, state_old, nodes, adjacency, aggregated_nodes, aggregated_arcs, training) -> tuple:
        """ ONE iteration of the loop body (GNN.py:223-242) through the same CUDA kernel (max_iteration = 1, threshold 0) """
        node_self = nodes if self.state_vect_dim else None
        _, state_new = state_loop(adjacency, self.net_state, state, node_self, aggregated_nodes, aggregated_arcs, max_iteration=1,
                                  threshold=0.0, training=bool(training), seed=self._next_seed())
        return k + 1, state_new, state, nodes, adjacency, aggregated_nodes, aggregated_arcs, training

    def apply_filters(self, state_converged, nodes, adjacency, arcs_label, mask_index) -> torch.Tensor:
        """ [states] or [states|labels] of the nodes with output_mask AND set_mask (GNN.py:245-248) """
        if self.state_vect_dim: state_converged = torch.cat([state_converged, nodes], dim=1)
        # every node selected (the index holds sorted, distinct positions: as many entries as rows = the identity): no gather, and no
        # index_add in the backward (0.09 + 0.25 ms on the 1M-node graph)
        if int(mask_index.shape[0]) == int(state_converged.shape[0]): return state_converged
        return state_converged.index_select(0, mask_index)

    def _next_seed(self):
        """ dropout seed of the next call.  Inside a captured training step (BaseClass.training_step with use_cuda_graph) the
        call counter lives on the device: the value is advanced by a captured op and handed to the kernels as a tensor """
        self._calls += 1
        state = getattr(self, '_seed_state', None)
        if state is not None:
            state.add_(0x85EBCA77).bitwise_and_(0xFFFFFFFF)
            return state.clone()              # this call's own copy: forward and backward kernels of the call read the same value
        return (self.dropout_seed * 0x9E3779B1 + self._calls * 0x85EBCA77) & 0xFFFFFFFF

    def _device_seed(self, enable: bool, device=None) -> None:
        """ switch the dropout call counter between host (python int) and device (int64 tensor) """
        if enable:
            start = (self.dropout_seed * 0x9E3779B1 + self._calls * 0x85EBCA77) & 0xFFFFFFFF
            self._seed_state = torch.full((), start, dtype=torch.int64, device=device)
        else:
            self._seed_state = None

    def _state_and_inputs(self, g: GraphTensor):
        """ prologue of Loop (GNN.py:257-268) """
        labels = g.arc_labels                     # = g.arcs[:, 2:] (does not materialise the id columns of a lazy GraphTensor)
        aggregated_arcs = sparse_dense(g.ArcNode, labels)
        n_nodes = g.nodes.shape[0]
        if self.state_vect_dim > 0:
            if self.initial_state is not None:
                state = self.initial_state.to(g.device, torch.float32)
                if tuple(state.shape) != (n_nodes, self.state_vect_dim): raise ValueError('initial_state has the wrong shape')
            else:
                state = 0.1 * torch.randn((n_nodes, self.state_vect_dim), dtype=torch.float32, device=g.device)
            aggregated_nodes = sparse_dense(g.Adjacency, g.nodes)
            node_self = g.nodes
        else:
            state = g.nodes
            aggregated_nodes = torch.zeros((n_nodes, 0), dtype=torch.float32, device=g.device)
            node_self = None
        return state, node_self, aggregated_nodes, aggregated_arcs, labels

    def Loop(self, g: Union[GraphObject, GraphTensor], *, training: bool = False, seed: Optional[int] = None):
        """ process a single graph, returning (iterations, states, output) (GNN.py:251-280) """
        if isinstance(g, GraphObject): g = GraphTensor.fromGraphObject(g)
        self.to(g.device)
        seed = self._next_seed() if seed is None else seed
        state, node_self, aggregated_nodes, aggregated_arcs, labels = self._state_and_inputs(g)
        k, state = state_loop(g.Adjacency, self.net_state, state, node_self, aggregated_nodes, aggregated_arcs,
                              max_iteration=self.max_iteration, threshold=self.state_threshold, training=training, seed=seed)
        mask_index = g.mask_index()
        out = self._output_one_pass(state, g.nodes, mask_index, training)
        if out is None:
            net_in = self.apply_filters(state, g.nodes, g.Adjacency, labels, mask_index)
            out = self.net_output(net_in, training=training, dropout_seed=seed, stream_base=16)
        return k, state, out

    def _output_one_pass(self, state, nodes, mask_index, training: bool):
        """ inference, every node selected, output net = ONE Dense layer with at most 16 units (the reference's default output MLP
        without its trailing BatchNormalization): act([state | labels] @ W + b) in one kernel of the library (gnn_output_dense)
        instead of concat + GEMM + bias + softmax (GNN.py:245-248, 279).  None when it does not apply. """
        if training or torch.is_grad_enabled() or state.device.type != 'cuda': return None
        if type(self).apply_filters is not GNNnodeBased.apply_filters: return None
        if int(mask_index.shape[0]) != int(state.shape[0]): return None
        layers = [l for l in self.net_output.layers if not isinstance(l, Dropout)]        # Dropout is the identity in inference
        if len(layers) != 1 or not isinstance(layers[0], Dense): return None
        dense = layers[0]
        if dense.units > 16 or dense.activation not in _native.ACT_CODES: return None
        return _native.output_dense(state, nodes if self.state_vect_dim else None, dense.kernel, dense.bias, dense.activation)


#######################################################################################################################
### CLASS GNN - EDGE BASED ############################################################################################
#######################################################################################################################
class GNNedgeBased(GNNnodeBased):
    """ GNN for edge-based problem """

    def apply_filters(self, state_converged, nodes, adjacency, arcs_label, mask_index) -> torch.Tensor:
        """ [x_dst | x_src | arc label] per arc (GNN.py:289-302): states gathered in the row-major (dst, src) order of the
        transposed adjacency, arc labels in the original arc order -- exactly the reference's pairing """
        if self.state_vect_dim: state_converged = torch.cat([state_converged, nodes], dim=1)
        idx = adjacency.indices
        states = state_converged[idx].reshape(arcs_label.shape[0], 2 * state_converged.shape[1])
        arc_state = torch.cat([states, arcs_label], dim=1)
        return arc_state.index_select(0, mask_index)


#######################################################################################################################
### CLASS GNN - GRAPH BASED ###########################################################################################
#######################################################################################################################
class GNNgraphBased(GNNnodeBased):
    """ GNN for graph-based problem """

    @staticmethod
    def get_filtered_tensor(g: GraphTensor, inp: torch.Tensor):
        """ targets / sample weights are per graph: no masking (GNN.py:313-315) """
        return inp.to(torch.float32)

    def Loop(self, g: Union[GraphObject, GraphTensor], *, training: bool = False, seed: Optional[int] = None):
        """ output of a graph-based problem is the NodeGraph-weighted sum of the node outputs (GNN.py:318-333) """
        if not g.has_nodegraph(): raise ValueError('WRONG GNN. NodeGraph is None: GNN is graph-based, while problem is non graph-based.')
        if isinstance(g, GraphObject): g = GraphTensor.fromGraphObject(g)
        it, state_nodes, out_nodes = super().Loop(g, training=training, seed=seed)
        return it, state_nodes, g.pool_nodes(out_nodes)
