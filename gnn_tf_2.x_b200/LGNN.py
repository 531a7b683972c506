# coding=utf-8
"""Layered GNN behind the reference's API (``GNN/LGNN.py``): a stack of GNNs in which layer i+1 sees the original
labels extended with the state and/or the output of layer i; serial, parallel and residual training.

Every layer's loop runs in the CUDA library; cross-layer gradients (parallel / residual) flow through the label
gradients the backward kernel returns (g_nodes / g_agg_nodes / g_agg_arcs and g_x0).
"""
from __future__ import annotations

import json
import os
from typing import Optional, Union

import numpy as np
import torch

from .GNN import GNNnodeBased, GNNgraphBased, GNNedgeBased
from .GNN_BaseClass import BaseClass
from .graph_class import GraphObject, GraphTensor
from .keras_compat import Dense, losses as _losses, optimizers as _optimizers


class LGNN(BaseClass):
    ## CONSTRUCTORS METHODS ###########################################################################################
    def __init__(self, gnns: list, get_state: bool, get_output: bool, optimizer, loss_function,
                 loss_arguments: Optional[dict], addressed_problem: str, extra_metrics: Optional[dict] = None,
                 extra_metrics_arguments: Optional[dict[str, dict]] = None, path_writer: str = 'writer/',
                 namespace: str = 'LGNN') -> None:
        """ CONSTRUCTOR (LGNN.py:15-61)

        :param gnns: (list) GNN instances of the same type, one per layer, initialized externally.
        :param get_state: (bool) node states are propagated through the layers.
        :param get_output: (bool) outputs on nodes/arcs are propagated through the layers.
        other parameters: as in GNNnodeBased.
        """
        kinds = set(type(i) for i in gnns)
        if len(kinds) != 1: raise TypeError('parameter <gnn> must contain gnns of the same type')
        super().__init__(optimizer, loss_function, loss_arguments, addressed_problem, extra_metrics, extra_metrics_arguments,
                         path_writer, namespace)
        self.get_state = get_state
        self.get_output = get_output
        self.gnns = gnns
        self.LAYERS = len(gnns)
        self.GNNS_TYPE = list(kinds)[0]
        self.namespace = [f'{namespace} - GNN{i}' for i in range(self.LAYERS)]
        self.training_mode = None
        for gnn, name in zip(self.gnns, self.namespace):
            gnn.namespace = [name]
            gnn.path_writer = f'{self.path_writer}{name}/'

    # -----------------------------------------------------------------------------------------------------------------
    def copy(self, *, path_writer: str = '', namespace: str = '', copy_weights: bool = True) -> 'LGNN':
        if not path_writer: path_writer = self.path_writer + '_copied/'
        if not namespace: namespace = 'LGNN'
        return self.__class__(gnns=[i.copy(copy_weights=copy_weights) for i in self.gnns], get_state=self.get_state,
                              get_output=self.get_output, optimizer=self.optimizer.__class__(**self.optimizer.get_config()),
                              loss_function=self.loss_function, loss_arguments=self.loss_args,
                              addressed_problem=self.addressed_problem, extra_metrics=self.extra_metrics,
                              extra_metrics_arguments=self.mt_args, path_writer=path_writer, namespace=namespace)

    ## SAVE AND LOAD METHODs ##########################################################################################
    def save(self, path: str):
        """ one sub-folder per layer + config.json (LGNN.py:83-104) """
        if path[-1] != '/': path += '/'
        if os.path.exists(path):
            import shutil
            shutil.rmtree(path)
        os.makedirs(path)
        for i, gnn in enumerate(self.gnns): gnn.save(f'{path}GNN{i}/')
        config = {'gnn_type': self.GNNS_TYPE.__name__, 'get_state': self.get_state, 'get_output': self.get_output,
                  'loss_function': _losses.serialize(self.loss_function), 'loss_arguments': self.loss_args,
                  'optimizer': _optimizers.serialize(self.optimizer), 'addressed_problem': self.addressed_problem}
        with open(f'{path}config.json', 'w') as f: json.dump(config, f)

    @classmethod
    def load(cls, path: str, path_writer: Optional[str] = None, namespace: str = 'LGNN', extra_metrics: Optional[dict] = None,
             extra_metrics_arguments: Optional[dict[str, dict]] = None):
        """ load from folder (LGNN.py:108-141); the GNN class of the layers is read from each layer's config """
        if path[-1] != '/': path += '/'
        if path_writer is None: path_writer = f'{path}writer'
        with open(f'{path}config.json') as f: config = json.load(f)
        kind = config.pop('gnn_type', 'GNNnodeBased')
        gnn_cls = {'GNNnodeBased': GNNnodeBased, 'GNNedgeBased': GNNedgeBased, 'GNNgraphBased': GNNgraphBased}[kind]
        layers = sorted(d for d in os.listdir(path) if d.startswith('GNN') and os.path.isdir(path + d))
        gnns = [gnn_cls.load(f'{path}{d}', path_writer=f'{path_writer}/{d}', namespace='GNN') for d in layers]
        optz = _optimizers.deserialize(config.pop('optimizer'))
        loss = _losses.deserialize(config.pop('loss_function'))
        return cls(gnns=gnns, optimizer=optz, loss_function=loss, extra_metrics=extra_metrics,
                   extra_metrics_arguments=extra_metrics_arguments, path_writer=path_writer, namespace=namespace, **config)

    ## GETTERS AND SETTERS METHODs ####################################################################################
    def get_dense_layers(self) -> list:
        return [l for gnn in self.gnns for l in gnn.get_dense_layers()]

    def trainable_variables(self) -> tuple:
        return [i.net_state.trainable_variables for i in self.gnns], [i.net_output.trainable_variables for i in self.gnns]

    def get_weights(self) -> tuple:
        return [i.net_state.get_weights() for i in self.gnns], [i.net_output.get_weights() for i in self.gnns]

    def set_weights(self, weights_state, weights_output) -> None:
        assert len(weights_state) == len(weights_output) == self.LAYERS
        for gnn, wst, wout in zip(self.gnns, weights_state, weights_output):
            gnn.net_state.set_weights(wst)
            gnn.net_output.set_weights(wout)

    ## CALL/PREDICT METHOD ############################################################################################
    def __call__(self, g) -> torch.Tensor:
        """ ONLY the output of the last layer, test mode """
        with torch.no_grad():
            return self.Loop(g, training=False)[-1][-1]

    def predict(self, g, idx: Union[int, list[int], range, str] = -1):
        """ output(s) of one or more layers in test mode (LGNN.py:172-198) """
        all_layers = range(self.LAYERS)
        if isinstance(idx, int):
            if idx < 0: idx += self.LAYERS
            assert idx in all_layers
        elif isinstance(idx, (list, range)):
            assert all(i in all_layers for i in idx)
            idx = sorted(idx)
        elif idx == 'all':
            idx = all_layers
        else:
            raise ValueError('param <idx> must be 1.int; 2.list of ordered ints in range(self.LAYERS); 3. str "all"')
        if isinstance(g, GraphObject): g = GraphTensor.fromGraphObject(g)
        with torch.no_grad():
            out = self.Loop(g, training=False)[-1]
        return out[idx].cpu().numpy() if isinstance(idx, int) else [out[i].cpu().numpy() for i in idx]

    ## EVALUATE METHODS ###############################################################################################
    def evaluate_single_graph(self, g, training: bool) -> tuple:
        """ (iterations per layer, loss, targets, output of the last layer) (LGNN.py:201-224).
        parallel: sum over targets of mean_i(loss(t, o_i) * w); residual (training only): loss(t, mean_i o_i) * w """
        if isinstance(g, GraphObject): g = GraphTensor.fromGraphObject(g)
        targs = self.GNNS_TYPE.get_filtered_tensor(g, g.targets)
        loss_weights = self.GNNS_TYPE.get_filtered_tensor(g, g.sample_weights)
        it, _, out = self.Loop(g, training=training)
        if training and self.training_mode == 'residual':
            loss = self.loss_function(targs, torch.stack(out, dim=0).mean(dim=0), **self.loss_args) * loss_weights
        else:
            loss = torch.stack([self.loss_function(targs, o, **self.loss_args) * loss_weights for o in out], dim=0).mean(dim=0)
        return it, loss.sum(), targs, out[-1]

    ## LOOP METHODS ###################################################################################################
    def update_graph(self, g: GraphTensor, state, output) -> GraphTensor:
        """ new GraphTensor whose node / arc labels are the ORIGINAL labels extended with the state and/or the output
        (scattered to the masked positions, zeros elsewhere) of the previous layer (LGNN.py:227-260) """
        g = g.copy()
        extra_nodes, extra_arcs = [], []
        if self.get_state: extra_nodes.append(state)
        if self.get_output:
            index = g.mask_index()
            rows = g.arcs.shape[0] if self.GNNS_TYPE == GNNedgeBased else g.nodes.shape[0]
            scattered = torch.zeros((rows, output.shape[1]), dtype=output.dtype, device=output.device).index_copy(0, index, output)
            (extra_arcs if self.GNNS_TYPE == GNNedgeBased else extra_nodes).append(scattered)
        if extra_nodes: g.nodes = torch.cat([g.nodes] + extra_nodes, dim=1)
        if extra_arcs: g.arcs = torch.cat([g.arcs] + extra_arcs, dim=1)
        return g

    def Loop(self, g, *, training: bool = False) -> tuple:
        """ (iterations per layer, state of the last layer, outputs of every layer) (LGNN.py:263-290) """
        if isinstance(g, GraphObject): g = GraphTensor.fromGraphObject(g)
        gtmp = g.copy()
        K, outs = [], []
        for gnn in self.gnns[:-1]:
            if isinstance(gnn, GNNgraphBased):
                # node-level loop; the pooled output only enters the loss list
                k, state, out = GNNnodeBased.Loop(gnn, gtmp, training=training)
                outs.append(gtmp.pool_nodes(out))
            else:
                k, state, out = gnn.Loop(gtmp, training=training)
                outs.append(out)
            K.append(k)
            gtmp = self.update_graph(g, state, out)
        k, state, out = self.gnns[-1].Loop(gtmp, training=training)
        return K + [k], state, outs + [out]

    ## TRAINING METHOD ################################################################################################
    def train(self, gTr, epochs: int, gVa=None, update_freq: int = 10, max_fails: int = 10, observed_metric: str = 'Loss',
              policy='min', *, mean: bool = True, training_mode: str = 'parallel', verbose: int = 3) -> None:
        """ LEARNING PROCEDURE (LGNN.py:293-344)

        :param training_mode: (str) in ['serial','parallel','residual']. Default 'parallel'
            > 'serial' - GNNs are trained separately, from layer 0 to layer N, each with its own optimizer
            > 'parallel' - GNNs are trained together, loss = mean_i(Loss(t, O_i))
            > 'residual' - GNNs are trained together, loss = Loss(t, mean_i(O_i))
        other parameters: as in BaseClass.train
        """
        assert training_mode in ['parallel', 'serial', 'residual']
        if (self.training_mode is not None) and (self.training_mode != training_mode):
            raise ValueError('training_mode cannot change once the LGNN has been trained')
        self.training_mode = training_mode
        gTr = self.checktype(gTr)
        gVa = self.checktype(gVa)

        if training_mode == 'serial':
            gTr1 = [i.copy() for i in gTr]
            gVa1 = [i.copy() for i in gVa] if gVa is not None else None
            node_loop = lambda gnn, graph: GNNnodeBased.Loop(gnn, graph) if isinstance(gnn, GNNgraphBased) else gnn.Loop(graph)
            for idx, gnn in enumerate(self.gnns):
                if verbose in [1, 3]: print(f'\n\n------------------- GNN{idx} -------------------\n')
                gnn.train(gTr1, epochs, gVa1, update_freq, max_fails, observed_metric, policy, mean=mean, verbose=verbose)
                with torch.no_grad():
                    _, sTr, oTr = zip(*[node_loop(gnn, i) for i in gTr1])
                    gTr1 = [self.update_graph(i, s, o) for i, s, o in zip(gTr, sTr, oTr)]
                    if gVa:
                        _, sVa, oVa = zip(*[node_loop(gnn, i) for i in gVa1])
                        gVa1 = [self.update_graph(i, s, o) for i, s, o in zip(gVa, sVa, oVa)]
        else:
            super().train(gTr, epochs, gVa, update_freq, max_fails, observed_metric, policy, mean=mean, verbose=verbose)
