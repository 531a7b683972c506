# coding=utf-8
"""Training driver shared by GNN and LGNN, behind the reference's ``BaseClass`` API (``GNN/GNN_BaseClass.py``):
train / evaluate / test / LKO, history helpers.  One optimizer step per (merged) batch graph; the forward and the
BPTT backward of every step run in the CUDA library through ``state_loop``.
"""
from __future__ import annotations

import os
import shutil
from abc import ABC, abstractmethod
from typing import Optional, Union

import numpy as np
import torch
from pandas import DataFrame

from . import GNN_metrics as mt
from .graph_class import GraphObject, GraphTensor


class BaseClass(ABC):
    ## CONSTRUCTORS METHODS ###########################################################################################
    def __init__(self, optimizer, loss_function, loss_arguments: Optional[dict], addressed_problem: str,
                 extra_metrics: Optional[dict] = None, extra_metrics_arguments: Optional[dict[str, dict]] = None,
                 path_writer: str = 'writer/', namespace='GNN') -> None:
        """ CONSTRUCTOR (GNN_BaseClass.py:19-63) - other attributes are defined in the inheriting class """
        if addressed_problem not in ['c', 'r']: raise ValueError('param <addressed_problem> not in [\'c\',\'r\']')
        if not isinstance(extra_metrics, (dict, type(None))): raise TypeError('type of param <extra_metrics> must be None or dict')
        self.loss_function = loss_function
        self.loss_args = dict() if loss_arguments is None else loss_arguments
        self.optimizer = optimizer
        self.addressed_problem = addressed_problem
        self.extra_metrics = dict() if extra_metrics is None else extra_metrics
        self.mt_args = dict() if extra_metrics_arguments is None else extra_metrics_arguments
        if path_writer[-1] != '/': path_writer += '/'
        if not isinstance(namespace, list): namespace = [namespace]
        if os.path.exists(path_writer): shutil.rmtree(path_writer)   # as the reference (GNN_BaseClass.py:58)
        self.path_writer = path_writer
        self.namespace = namespace
        self.history = dict()

    ## ABSTRACT METHODS ###############################################################################################
    @abstractmethod
    def copy(self, *, path_writer: str = '', namespace: str = '', copy_weights: bool = True): pass

    @abstractmethod
    def save(self, path: str) -> None: pass

    @abstractmethod
    def get_dense_layers(self) -> list: pass

    @abstractmethod
    def trainable_variables(self) -> tuple: pass

    @abstractmethod
    def get_weights(self) -> tuple: pass

    @abstractmethod
    def set_weights(self, weights_state, weights_output) -> None: pass

    @abstractmethod
    def Loop(self, g, *, training: bool = False) -> tuple: pass

    ## HISTORY METHOD #################################################################################################
    def printHistory(self) -> None:
        print('\n', DataFrame(self.history), end='\n\n')

    def saveHistory_csv(self, path) -> None:
        if path[-4:] != '.csv': path += '.csv'
        DataFrame(self.history).to_csv(path, index=False)

    def saveHistory_txt(self, path) -> None:
        if path[-4:] != '.txt': path += '.txt'
        with open(path, 'w') as txt: txt.write(DataFrame(self.history).to_string(index=False))

    ## EVALUATE METHODs ###############################################################################################
    def evaluate_single_graph(self, g, training: bool) -> tuple:
        """ returns iteration, loss, targets and outputs of one graph; defined in the inheriting class """
        pass

    def evaluate(self, g) -> tuple:
        """ metrics in self.extra_metrics + 'It' & 'Loss' for one graph or a list of graphs, test mode
        (GNN_BaseClass.py:165-189). :return: metrics, y_true, y_pred, targets, y_score """
        g = self.checktype(g)
        with torch.no_grad():
            iters, losses, targets, outs = zip(*[self.evaluate_single_graph(i, training=False) for i in g])
        targets = torch.cat(targets, dim=0)
        y_score = torch.cat(outs, dim=0)
        y_true = torch.argmax(targets, dim=1) if self.addressed_problem == 'c' else targets
        y_pred = torch.argmax(y_score, dim=1) if self.addressed_problem == 'c' else y_score
        yt, yp = y_true.cpu().numpy(), y_pred.cpu().numpy()
        metrics = {k: float(np.mean(self.extra_metrics[k](yt, yp, **self.mt_args.get(k, dict())))) for k in self.extra_metrics}
        # iteration counts: each entry is a scalar (GNN) or a list of scalars (LGNN)
        flat_iters = torch.stack([torch.as_tensor(j, dtype=torch.float32, device=targets.device).reshape(())
                                  for i in iters for j in (i if isinstance(i, (list, tuple)) else [i])])
        metrics['It'] = int(flat_iters.mean())
        metrics['Loss'] = float(torch.stack([l.reshape(()) for l in losses]).mean())
        return metrics, y_true, y_pred, targets, y_score

    ## TRAINING METHOD ################################################################################################
    def _regularizer_terms(self):
        extra_loss = 0
        for layer in self.get_dense_layers():
            if layer.kernel_regularizer is not None: extra_loss = extra_loss + layer.kernel_regularizer(layer.kernel)
            if layer.bias_regularizer is not None: extra_loss = extra_loss + layer.bias_regularizer(layer.bias)
        return extra_loss

    def training_step(self, g: GraphTensor, mean: bool = True):
        """ one optimizer step on one (batch) graph (GNN_BaseClass.py:231-247): BPTT gradient of the summed loss
        (+ regularisers); net_state gradients divided by the iteration count when ``mean``.
        With ``self.use_cuda_graph = True`` the whole step (prologue SpMMs, the loop's launches, output net, loss, backward
        sweep, gradient all-reduce, optimizer) is captured ONCE per batch graph into a CUDA graph and replayed afterwards:
        one launch per step instead of a hundred (the iteration count, the dropout seeds and the Adam step counter live on
        the device, so a replay is a genuine new step).
        :return: (iterations, loss) as device tensors -- nothing is synchronised """
        if getattr(self, 'use_cuda_graph', False) and g.device.type == 'cuda':
            return self._graphed_training_step(g, mean)
        return self._training_step_eager(g, mean)

    def _step_gradients(self, g: GraphTensor, mean: bool):
        """ forward + BPTT: (iterations, loss, gradients in the order of `flat`, flat list of the trainable variables) """
        iters, loss, *_ = self.evaluate_single_graph(g, training=True)
        loss = loss + self._regularizer_terms()
        wS, wO = self.trainable_variables()
        flat = [v for layer in wS + wO for v in layer]
        grads = torch.autograd.grad(loss, flat, allow_unused=True)
        grads = [torch.zeros_like(v) if gr is None else gr for v, gr in zip(flat, grads)]
        if not isinstance(iters, list): iters = [iters]
        pos, dW = 0, []
        for layer_idx, layer in enumerate(wS):
            for _ in layer:
                dW.append(grads[pos] / iters[layer_idx] if mean else grads[pos])
                pos += 1
        dW += grads[pos:]
        assert len(dW) == len(flat)
        return iters, loss.detach(), dW, flat

    def _training_step_eager(self, g: GraphTensor, mean: bool = True):
        iters, loss, dW, flat = self._step_gradients(g, mean)
        if getattr(self, 'distributed', False):   # graph batches sharded by whole graph over the ranks: one all-reduce of the flat gradient
            from .dist_graph import allreduce_gradients
            dW = allreduce_gradients(dW, getattr(self, 'process_group', None))
        self.optimizer.apply_gradients(zip(dW, flat))
        return iters, loss

    def _seeded_models(self) -> list:
        """ the GNNs whose dropout call counter must live on the device inside a captured step """
        return list(getattr(self, 'gnns', [self]))

    def _graphed_training_step(self, g: GraphTensor, mean: bool):
        """ capture-and-replay of the training step for one batch graph (cache on the GraphTensor, keyed by model and `mean`).
        Replaces the eager while_loop / GradientTape replay of GNN.py:271-272 + GNN_BaseClass.py:233-247 by one graph launch.
        Single process: the optimizer step is part of the graph.  Sharded graph batches (self.distributed): the graph ends with
        the flat gradient; the NCCL all-reduce and the optimizer follow it on the stream (two more launches, no host sync). """
        cache = g.__dict__.setdefault('_step_graphs', dict())
        key = (id(self), bool(mean))
        entry = cache.get(key)
        distributed = bool(getattr(self, 'distributed', False))
        if entry is None:
            if not hasattr(self.optimizer, 'capturable'):
                raise NotImplementedError('use_cuda_graph needs an optimizer whose step is free of host state (keras_compat.Adam)')
            self.optimizer.capturable = True
            for m in self._seeded_models():
                if getattr(m, '_seed_state', None) is None: m._device_seed(True, g.device)
            # the first step of a batch graph runs eagerly on a side stream (PyTorch's capture protocol: lazy initialisation,
            # allocator warm-up, optimizer slots); the second call captures, every later call replays.  One optimizer step per call.
            side = torch.cuda.Stream(device=g.device)
            side.wait_stream(torch.cuda.current_stream(g.device))
            with torch.cuda.stream(side):
                out = self._training_step_eager(g, mean)
            torch.cuda.current_stream(g.device).wait_stream(side)
            cache[key] = ('warm',)
            return out
        if entry[0] == 'warm':
            graph = torch.cuda.CUDAGraph()
            # thread_local: other threads (the NCCL watchdog of torch.distributed) may call the CUDA API while this thread captures
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                iters, loss, dW, flat = self._step_gradients(g, mean)
                if distributed:
                    packed = torch.cat([d.reshape(-1) for d in dW])
                else:
                    packed = None
                    self.optimizer.apply_gradients(zip(dW, flat))
            entry = cache[key] = ('graph', graph, iters, loss, packed, flat)
        _, graph, iters, loss, packed, flat = entry
        graph.replay()
        if distributed:
            import torch.distributed as dist
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=getattr(self, 'process_group', None))
            views, pos = [], 0
            for v in flat:
                views.append(packed[pos:pos + v.numel()].view_as(v))
                pos += v.numel()
            self.optimizer.apply_gradients(zip(views, flat))
        return iters, loss           # static tensors of the captured step: valid until the next replay

    def train(self, gTr, epochs: int, gVa=None, update_freq: int = 10, max_fails: int = 10, observed_metric='Loss',
              policy='min', *, mean: bool = True, verbose: int = 3) -> None:
        """ TRAINING PROCEDURE (GNN_BaseClass.py:192-335)

        :param gTr: element/list of GraphObjects/GraphTensors used for the learning procedure.
        :param epochs: (int) the max number of epochs for the learning procedure.
        :param gVa: element/list of GraphsObjects/GraphTensors for early stopping. Default None, no early stopping.
        :param update_freq: (int) epochs between two evaluations of gTr/gVa. Default 10.
        :param max_fails: (int) max number of failures in gVa improvement before early stopping. Default 10.
        :param observed_metric: (str) key of the metric observed for early stopping
        :param policy: (str) 'min' | 'max': minimise / maximise observed_metric
        :param mean: (bool) net_state gradients are averaged over the iterations (True) or summed. Default True.
        :param verbose: (int) 0: silent; 1: history; 2: epochs/batches; 3: history + epochs/batches. Default 3.
        """
        if verbose not in range(4): raise ValueError('param <verbose> not in [0,1,2,3]')
        gTr = self.checktype(gTr)
        gVa = self.checktype(gVa)

        if not self.history:
            keys = ['Epoch'] + [i + j for i in ['It', 'Loss'] + list(self.extra_metrics) for j in ([' Tr', ' Va'] if gVa else [' Tr'])]
            if gVa: keys += ['Fail', f'Best {observed_metric} Va']
            self.history.update({i: list() for i in keys})
            os.makedirs(self.path_writer, exist_ok=True)
        writers = _Writers(self.path_writer, bool(gVa))

        def update_history(name: str, val: dict) -> None:
            if name not in ['Tr', 'Va']: raise TypeError('param <name> must be \'Tr\' or \'Va\'')
            for key in val: self.history[f'{key} {name}'].append(val[key])

        def checkpoint(value: float):
            wst, wout = self.get_weights()
            return value, 0, wst, wout

        if gVa:
            assert policy in ['min', 'max']
            best_key = f'Best {observed_metric} Va'
            better, start = (np.less, float(1e30)) if policy == 'min' else (np.greater, float(-1e30))
            start = self.history[best_key][-1] if self.history[best_key] else start
            best_value, fails, ws, wo = checkpoint(start)

        initial_epoch = self.history['Epoch'][-1] + 1 if self.history['Epoch'] else 0
        epochs += initial_epoch
        e = initial_epoch
        for e in range(initial_epoch, epochs):
            for i, elem in enumerate(gTr):
                self.training_step(elem, mean=mean)
                if verbose > 2: print(f' > Epoch {e:4d}/{epochs} \t\t> Batch {i + 1:4d}/{len(gTr)}', end='\r')

            if e % update_freq == 0:
                metricsTr, *_ = self.evaluate(gTr)
                self.history['Epoch'].append(e)
                update_history('Tr', metricsTr)
                writers.scalars('Training', metricsTr, e)
                for wst, wout, namespace in zip(*self.get_weights(), self.namespace):
                    writers.weights('Net - State', namespace, 'N1', wst, e)
                    writers.weights('Net - Output', namespace, 'N2', wout, e)

            if (e % update_freq == 0) and gVa:
                metricsVa, *_ = self.evaluate(gVa)
                new_value = metricsVa[observed_metric]
                if better(new_value, best_value): best_value, fails, ws, wo = checkpoint(new_value)
                else: fails += 1
                self.history[best_key].append(best_value)
                self.history['Fail'].append(fails)
                update_history('Va', metricsVa)
                writers.scalars('Validation', metricsVa, e)
                if fails >= max_fails:
                    if verbose in [1, 3]: self.printHistory()
                    if verbose: print('\r Validation Stop')
                    break

            if (e % update_freq == 0) and verbose in [1, 3]: self.printHistory()
        else:
            if verbose: print('\r End of Epochs Stop')

        if gVa: self.set_weights(ws, wo)   # best weights seen on the validation set
        for wst, wout, namespace in zip(*self.get_weights(), self.namespace):
            writers.weights('Net - State', namespace, 'N1', wst, e)
            writers.weights('Net - Output', namespace, 'N2', wout, e)
        writers.close()

    ## TEST METHOD ####################################################################################################
    def test(self, gTe, *, rocdir: str = '', micro_and_macro: bool = False, prisofsdir: str = '', pos_label=0) -> dict:
        """ TEST PROCEDURE (GNN_BaseClass.py:338-359): metrics of gTe (+ optional ROC / precision-recall output) """
        gTe = self.checktype(gTe)
        metricsTe, y_true, y_pred, targets, y_score = self.evaluate(gTe)
        if rocdir: mt.ROC(targets, y_score, rocdir, micro_and_macro, pos_label=pos_label)
        if prisofsdir: mt.PRISOFS(targets, y_score, prisofsdir, pos_label=pos_label)
        return metricsTe

    ## K-FOLD CROSS VALIDATION METHOD #################################################################################
    def LKO(self, batches: tuple, epochs: int = 500, training_mode=None, update_freq: int = 10, max_fails: int = 10,
            observed_metric: str = 'Loss', policy='min', mean: bool = True, verbose: int = 3) -> dict:
        """ LEAVE K OUT CROSS VALIDATION (GNN_BaseClass.py:362-402). ``batches`` is the output of prepare_LKO_data:
        (training sets, test sets, validation sets). A fresh re-initialised copy of the model is trained per fold. """
        metrics = {i: list() for i in list(self.extra_metrics) + ['It', 'Loss']}
        kwargs = dict()
        if training_mode: kwargs['training_mode'] = training_mode
        number_of_batches = len(batches[0])
        for i, (gTr, gTe, gVa) in enumerate(zip(*batches)):
            if verbose: print(f'\nBATCH K-OUT {i + 1}/{number_of_batches}')
            temp = self.copy(copy_weights=False, path_writer=f'{self.path_writer}{i}', namespace=f'Batch {i + 1}-{number_of_batches}')
            temp.train(gTr, epochs, gVa, update_freq, max_fails, observed_metric, policy, mean=mean, verbose=verbose, **kwargs)
            res = temp.test(gTe)
            for m in res: metrics[m].append(res[m])
            if verbose > 1: print(f'\nRESULTS BATCH {i + 1}/{number_of_batches}\n', DataFrame(res, index=['res']).transpose())
        return metrics

    ## STATIC METHODs #################################################################################################
    @staticmethod
    def get_filtered_tensor(g: GraphTensor, inp: torch.Tensor):
        """ rows of inp [targets or sample_weights, one per output_mask row] that belong to set_mask (GNN_BaseClass.py:405-409) """
        return inp.index_select(0, g.filtered_index())

    @staticmethod
    def checktype(elem) -> Optional[list[GraphTensor]]:
        """ None, or a list of GraphTensors from GraphObject(s)/GraphTensor(s) (GNN_BaseClass.py:412-425) """
        if elem is None:
            pass
        elif isinstance(elem, GraphTensor):
            elem = [elem]
        elif isinstance(elem, GraphObject):
            elem = [GraphTensor.fromGraphObject(elem)]
        elif isinstance(elem, (list, tuple)) and all(isinstance(g, (GraphObject, GraphTensor)) for g in elem):
            elem = [GraphTensor.fromGraphObject(g) if isinstance(g, GraphObject) else g for g in elem]
        else:
            raise TypeError('Error - <gTr> and/or <gVa> are not GraphObject/GraphTensor or LIST/TUPLE of GraphObjects/GraphTensors')
        return elem


class _Writers:
    """ TensorBoard scalars / weight histograms (GNN_BaseClass.py:428-459) when torch.utils.tensorboard is usable,
    silently nothing otherwise: logging is not part of the hot path """
    _names = {'Acc': 'Accuracy', 'Bacc': 'Balanced Accuracy', 'Ck': 'Cohen\'s Kappa', 'Js': 'Jaccard Score', 'Fs': 'F1-Score',
              'Prec': 'Precision Score', 'Rec': 'Recall Score', 'Tpr': 'TPR', 'Tnr': 'TNR', 'Fpr': 'FPR', 'Fnr': 'FNR',
              'Loss': 'Loss', 'It': 'Iteration @ Convergence'}

    def __init__(self, path: str, validation: bool):
        self._w = dict()
        if os.environ.get('GNN_B200_TENSORBOARD', '0') != '1': return
        try:
            from torch.utils.tensorboard import SummaryWriter
            for name in ['Net - State', 'Net - Output', 'Training'] + (['Validation'] if validation else []):
                self._w[name] = SummaryWriter(f'{path}{name}')
        except Exception:
            self._w = dict()

    def scalars(self, writer: str, metrics: dict, epoch: int) -> None:
        if not isinstance(metrics, dict): raise TypeError('type of param <metrics> must be dict')
        w = self._w.get(writer)
        if w is None: return
        for key, value in metrics.items(): w.add_scalar(self._names.get(key, key), value, epoch)

    def weights(self, writer: str, namespace: str, net_name: str, val_list: list, epoch: int) -> None:
        w = self._w.get(writer)
        if w is None: return
        for i, arr in enumerate(val_list): w.add_histogram(f'{namespace}/{net_name} V{i}', arr, epoch)

    def close(self):
        for w in self._w.values(): w.close()
