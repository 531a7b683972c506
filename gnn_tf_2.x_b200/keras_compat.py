# coding=utf-8
"""Host-side stand-ins for the Keras objects the reference API hands around.

The reference passes ``tf.keras`` objects through its constructors (``GNN/GNN.py:22-35``,
``GNN/MLP.py:62-64``, ``starter.py:81-83``): ``Sequential`` MLPs, an optimizer, a loss function.
TensorFlow is not allowed on this path, so these small classes keep the same names, constructor
arguments and numerics (Keras defaults restated from SURVEY.md section 8c) on top of torch tensors:

  * ``Dense`` / ``Dropout`` / ``BatchNormalization`` / ``Sequential``  -- parameter holders + a
    *description* (``Sequential.spec()``) that the CUDA kernels evaluate. ``Sequential.__call__``
    is a plain torch evaluation with the same semantics; the state loop never uses it.
  * ``Adam`` / ``SGD`` with the Keras update formula and ``apply_gradients(zip(grads, vars))``.
  * ``categorical_crossentropy`` / ``mean_squared_error`` / ... with Keras semantics.
  * the counter-based dropout generator shared bit-for-bit with the CUDA kernels and the oracle.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Union

import numpy as np
import torch

SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805

# activation codes shared with include/gnn_b200.h (GNN_ACT_*)
ACTIVATION_CODES = {'linear': 0, None: 0, 'relu': 1, 'tanh': 2, 'sigmoid': 3, 'selu': 4, 'elu': 5, 'softmax': 6,
                    'softplus': 7}


def _act_name(act) -> str:
    if act is None: return 'linear'
    if isinstance(act, str): return act.lower()
    name = getattr(act, '__name__', None)
    if name in ACTIVATION_CODES: return name
    raise ValueError(f'unsupported activation {act!r}: supported {sorted(k for k in ACTIVATION_CODES if k)}')


def apply_activation(name: str, x: torch.Tensor) -> torch.Tensor:
    """ torch evaluation of a Keras activation string """
    if name == 'linear': return x
    if name == 'relu': return torch.relu(x)
    if name == 'tanh': return torch.tanh(x)
    if name == 'sigmoid': return torch.sigmoid(x)
    if name == 'selu': return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * torch.expm1(x))
    if name == 'elu': return torch.where(x > 0, x, torch.expm1(x))
    if name == 'softmax': return torch.softmax(x, dim=-1)
    if name == 'softplus': return torch.nn.functional.softplus(x)
    raise ValueError(f'unsupported activation {name!r}')


#######################################################################################################################
## COUNTER-BASED DROPOUT GENERATOR ####################################################################################
#######################################################################################################################
_M32 = 0xFFFFFFFF


def _fmix32_int(x: int) -> int:
    x &= _M32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & _M32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & _M32
    x ^= x >> 16
    return x


def dropout_key(seed: int, stream: int, step: int) -> int:
    """ 32-bit key of one (call seed, dropout layer, loop iteration) triple; same formula in csrc/rng.cuh """
    a = _fmix32_int(seed + 0x9E3779B9 * (stream + 1))
    return _fmix32_int(a ^ ((step * 0x85EBCA6B + 0x27D4EB2F) & _M32))


def _fmix32_t(x: torch.Tensor) -> torch.Tensor:
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & _M32
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & _M32
    x = x ^ (x >> 16)
    return x


def dropout_key_t(seed: torch.Tensor, stream: int, step: int) -> torch.Tensor:
    """ dropout_key for a seed held on the device (int64 tensor with a value in [0, 2^32)): same arithmetic, torch ops only,
    so that a CUDA-graph replay draws the mask of the CURRENT seed value """
    a = _fmix32_t(seed + ((0x9E3779B9 * (stream + 1)) & _M32))
    return _fmix32_t(a ^ ((step * 0x85EBCA6B + 0x27D4EB2F) & _M32))


def dropout_keep_mask(seed, stream: int, step: int, rows: int, cols: int, rate: float, device=None,
                      row_offset: int = 0) -> torch.Tensor:
    """ boolean keep-mask [rows, cols]: element (n, j) is kept iff u(n, j) >= rate, with
    u = (hash >> 8) * 2**-24 and hash = fmix32(fmix32(lo ^ key) + hi * 0xC2B2AE35 + 0x165667B1), idx = n * cols + j.
    seed: python int, or an int64 device tensor (device-side seed of a captured training step) """
    key = dropout_key_t(seed, stream, step) if isinstance(seed, torch.Tensor) else dropout_key(seed, stream, step)
    n = torch.arange(row_offset, row_offset + rows, dtype=torch.int64, device=device)[:, None]
    j = torch.arange(cols, dtype=torch.int64, device=device)[None, :]
    idx = n * cols + j
    lo, hi = idx & _M32, (idx >> 32) & _M32
    v = _fmix32_t(lo ^ key)
    v = _fmix32_t((v + hi * 0xC2B2AE35 + 0x165667B1) & _M32)
    u = (v >> 8).to(torch.float32) * (1.0 / 16777216.0)
    return u >= rate


#######################################################################################################################
## INITIALIZERS #######################################################################################################
#######################################################################################################################
_TRUNC_STD = 0.87962566103423978


def _fans(shape) -> tuple[int, int]:
    if len(shape) == 1: return shape[0], shape[0]
    return shape[0], shape[1]


def initialize(name: Union[str, Callable, None], shape, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """ Keras initializer strings (variance-scaling family uses the truncated normal of Keras) """
    if callable(name): return torch.as_tensor(name(shape), dtype=torch.float32)
    name = 'glorot_uniform' if name is None else name.lower()
    fan_in, fan_out = _fans(shape)
    out = torch.empty(shape, dtype=torch.float32)

    def trunc_normal(std):
        torch.nn.init.trunc_normal_(out, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=generator)
        return out

    def uniform(limit):
        return out.uniform_(-limit, limit, generator=generator)

    if name == 'zeros': return out.zero_()
    if name == 'ones': return out.fill_(1.0)
    if name == 'lecun_normal': return trunc_normal(math.sqrt(1.0 / fan_in) / _TRUNC_STD)
    if name == 'glorot_normal': return trunc_normal(math.sqrt(2.0 / (fan_in + fan_out)) / _TRUNC_STD)
    if name == 'he_normal': return trunc_normal(math.sqrt(2.0 / fan_in) / _TRUNC_STD)
    if name == 'lecun_uniform': return uniform(math.sqrt(3.0 / fan_in))
    if name == 'glorot_uniform': return uniform(math.sqrt(6.0 / (fan_in + fan_out)))
    if name == 'he_uniform': return uniform(math.sqrt(6.0 / fan_in))
    if name == 'random_normal': return out.normal_(0.0, 0.05, generator=generator)
    if name == 'random_uniform': return uniform(0.05)
    raise ValueError(f'unsupported initializer {name!r}')


#######################################################################################################################
## LAYERS #############################################################################################################
#######################################################################################################################
class Layer:
    trainable_variables: list
    non_trainable_variables: list

    def to(self, device):
        for t in self.trainable_variables + self.non_trainable_variables:
            t.data = t.data.to(device)
        return self


class Dense(Layer):
    """ ``act(x @ kernel[in, out] + bias[out])`` """

    def __init__(self, units: int, activation=None, kernel_initializer='glorot_uniform', bias_initializer='zeros',
                 kernel_regularizer=None, bias_regularizer=None, input_shape=None, **_):
        self.units = int(units)
        self.activation = _act_name(activation)
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.input_dim = None if input_shape is None else int(input_shape[0])
        self.kernel: Optional[torch.Tensor] = None
        self.bias: Optional[torch.Tensor] = None

    def build(self, input_dim: int, device, generator=None):
        self.input_dim = int(input_dim)
        self.kernel = initialize(self.kernel_initializer, (self.input_dim, self.units), generator).to(device).requires_grad_()
        self.bias = initialize(self.bias_initializer, (self.units,), generator).to(device).requires_grad_()
        return self.units

    @property
    def trainable_variables(self): return [self.kernel, self.bias]

    @property
    def non_trainable_variables(self): return []

    def config(self):
        return dict(units=self.units, activation=self.activation, kernel_initializer=self.kernel_initializer,
                    bias_initializer=self.bias_initializer, kernel_regularizer=self.kernel_regularizer,
                    bias_regularizer=self.bias_regularizer)


class Dropout(Layer):
    """ training: ``x * keep / (1 - rate)``; inference: identity """
    alpha = False

    def __init__(self, rate: float, **_):
        self.rate = float(rate)

    def build(self, input_dim, device, generator=None): return input_dim

    trainable_variables = property(lambda self: [])
    non_trainable_variables = property(lambda self: [])

    def config(self): return dict(rate=self.rate)


class AlphaDropout(Dropout):
    """ kept for API compatibility (MLP.py:61); the CUDA path rejects it explicitly """
    alpha = True


class BatchNormalization(Layer):
    """ Keras defaults: axis=-1, momentum=0.99, epsilon=1e-3, gamma=1, beta=0; biased batch variance in training,
    moving <- moving * momentum + batch * (1 - momentum) at every call with training=True """

    def __init__(self, momentum: float = 0.99, epsilon: float = 1e-3, **_):
        self.momentum, self.epsilon = float(momentum), float(epsilon)
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    def build(self, input_dim, device, generator=None):
        self.gamma = torch.ones(input_dim, dtype=torch.float32, device=device).requires_grad_()
        self.beta = torch.zeros(input_dim, dtype=torch.float32, device=device).requires_grad_()
        self.moving_mean = torch.zeros(input_dim, dtype=torch.float32, device=device)
        self.moving_variance = torch.ones(input_dim, dtype=torch.float32, device=device)
        return input_dim

    @property
    def trainable_variables(self): return [self.gamma, self.beta]

    @property
    def non_trainable_variables(self): return [self.moving_mean, self.moving_variance]

    def config(self): return dict(momentum=self.momentum, epsilon=self.epsilon)


class _TallDense(torch.autograd.Function):
    """ ``x @ kernel + bias`` for MANY rows and few units (the output net of a million-node graph).  Forward as usual; in the
    backward the weight gradient ``x^T g`` is a reduction over the rows, for which the library GEMM picks a split-K kernel
    that runs at a few percent of the memory bandwidth (``sgemm_largek``: 1.4 ms for 1M x 35 against 2 output units, 3 % of a C4
    training step and a third of a C5 one).  Here: partial products over blocks of 512 rows as ONE batched GEMM, then a sum over
    the blocks -- a fixed order, so the result is deterministic, and more accurate than one long fp32 accumulation. """
    BLOCK = 512

    @staticmethod
    def forward(ctx, x, kernel, bias):
        ctx.save_for_backward(x, kernel)
        return x @ kernel + bias

    @staticmethod
    def backward(ctx, g):
        x, kernel = ctx.saved_tensors
        g = g.contiguous()
        gx = g @ kernel.t() if ctx.needs_input_grad[0] else None
        gk = gb = None
        if ctx.needs_input_grad[1]:
            xc, S = x.contiguous(), _TallDense.BLOCK
            n0 = (xc.shape[0] // S) * S
            gk = torch.bmm(xc[:n0].view(-1, S, xc.shape[1]).transpose(1, 2), g[:n0].view(-1, S, g.shape[1])).sum(0)
            if n0 < xc.shape[0]: gk = gk + xc[n0:].t() @ g[n0:]
        if ctx.needs_input_grad[2]: gb = g.sum(0)
        return gx, gk, gb


def dense_affine(x: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """ ``x @ kernel + bias``; tall inputs (>= 16 384 rows) take the block-wise weight gradient of _TallDense """
    if x.dim() == 2 and x.shape[0] >= 16384 and torch.is_grad_enabled() and (kernel.requires_grad or bias.requires_grad or x.requires_grad):
        return _TallDense.apply(x, kernel, bias)
    return x @ kernel + bias


#######################################################################################################################
## SEQUENTIAL #########################################################################################################
#######################################################################################################################
class MLPSpec:
    """ what the kernels need to know about a Sequential: Dense chain, dropout in front of each Dense (index l) or
    behind the last one (index L), optional trailing BatchNormalization """

    def __init__(self, dims, activations, dense_layers, dropout_rates, alpha_dropout, batchnorm):
        self.dims: list[int] = dims
        self.activations: list[str] = activations
        self.dense_layers: list[Dense] = dense_layers
        self.dropout_rates: list[float] = dropout_rates
        self.alpha_dropout: bool = alpha_dropout
        self.batchnorm: Optional[BatchNormalization] = batchnorm

    @property
    def n_layers(self): return len(self.dense_layers)

    @property
    def has_dropout(self): return any(r > 0 for r in self.dropout_rates)


class Sequential:
    """ ordered list of Dense / Dropout / BatchNormalization layers with the Keras weight ordering """

    def __init__(self, layers: list[Layer], input_dim: Optional[int] = None, device=None, seed: Optional[int] = None):
        self.layers = list(layers)
        if input_dim is None:
            first_dense = next(l for l in self.layers if isinstance(l, Dense))
            input_dim = first_dense.input_dim
        if input_dim is None: raise ValueError('Sequential needs the input dimension (input_shape on the first Dense)')
        self.input_dim = int(input_dim)
        self.device = torch.device('cpu') if device is None else torch.device(device)
        generator = None
        if seed is not None:
            generator = torch.Generator().manual_seed(int(seed))
        dim = self.input_dim
        for layer in self.layers: dim = layer.build(dim, self.device, generator)
        self.output_dim = dim
        self._check_structure()

    # structure -------------------------------------------------------------------------------------------------------
    def _check_structure(self):
        kinds = [type(l) for l in self.layers]
        if BatchNormalization in kinds[:-1]:
            raise ValueError('BatchNormalization is supported only as the last layer (what MLP() builds, MLP.py:63)')
        if not any(k is Dense for k in kinds): raise ValueError('Sequential needs at least one Dense layer')

    def spec(self) -> MLPSpec:
        dense = [l for l in self.layers if isinstance(l, Dense)]
        rates = [0.0] * (len(dense) + 1)
        alpha, seen = False, 0
        for layer in self.layers:
            if isinstance(layer, Dense): seen += 1
            elif isinstance(layer, Dropout):
                # two dropouts at the same position compose: keep-prob multiply is NOT what Keras does (two masks);
                # the reference builder never produces it, so reject
                if rates[seen] > 0: raise ValueError('two Dropout layers at the same position are not supported')
                rates[seen] = layer.rate
                alpha = alpha or layer.alpha
        bn = self.layers[-1] if isinstance(self.layers[-1], BatchNormalization) else None
        return MLPSpec([self.input_dim] + [d.units for d in dense], [d.activation for d in dense], dense, rates, alpha, bn)

    def to(self, device):
        self.device = torch.device(device)
        for layer in self.layers: layer.to(self.device)
        return self

    # weights ---------------------------------------------------------------------------------------------------------
    @property
    def trainable_variables(self) -> list[torch.Tensor]:
        return [v for l in self.layers for v in l.trainable_variables]

    @property
    def variables(self) -> list[torch.Tensor]:
        return [v for l in self.layers for v in l.trainable_variables + l.non_trainable_variables]

    def get_weights(self) -> list[np.ndarray]:
        return [v.detach().cpu().numpy().copy() for v in self.variables]

    def set_weights(self, weights) -> None:
        variables = self.variables
        if len(weights) != len(variables):
            raise ValueError(f'expected {len(variables)} weight arrays, got {len(weights)}')
        with torch.no_grad():
            for v, w in zip(variables, weights):
                w = torch.as_tensor(np.asarray(w), dtype=torch.float32)
                if tuple(w.shape) != tuple(v.shape): raise ValueError(f'weight shape {tuple(w.shape)} != {tuple(v.shape)}')
                v.copy_(w)

    def clone(self, seed: Optional[int] = None) -> 'Sequential':
        """ ``tf.keras.models.clone_model``: same architecture, freshly initialised weights """
        new_layers = [type(l)(**l.config()) for l in self.layers]
        return Sequential(new_layers, input_dim=self.input_dim, device=self.device, seed=seed)

    # plain torch evaluation (used for small heads and by tests; the state loop runs in CUDA) --------------------------
    def __call__(self, x: torch.Tensor, training: bool = False, *, dropout_seed: int = 0, stream_base: int = 0,
                 step: int = 0, update_moving: bool = True) -> torch.Tensor:
        seen = 0
        for layer in self.layers:
            if isinstance(layer, Dense):
                x = apply_activation(layer.activation, dense_affine(x, layer.kernel, layer.bias))
                seen += 1
            elif isinstance(layer, Dropout):
                if layer.alpha: raise NotImplementedError('AlphaDropout is not implemented on this path')
                if training and layer.rate > 0:
                    keep = dropout_keep_mask(dropout_seed, stream_base + seen, step, x.shape[0], x.shape[1], layer.rate, x.device)
                    x = torch.where(keep, x * (1.0 / (1.0 - layer.rate)), torch.zeros_like(x))
            elif isinstance(layer, BatchNormalization):
                if training:
                    mean = x.mean(dim=0)
                    var = x.var(dim=0, unbiased=False)
                    if update_moving:
                        with torch.no_grad():
                            layer.moving_mean.mul_(layer.momentum).add_(mean.detach() * (1 - layer.momentum))
                            layer.moving_variance.mul_(layer.momentum).add_(var.detach() * (1 - layer.momentum))
                else:
                    mean, var = layer.moving_mean, layer.moving_variance
                x = (x - mean) * torch.rsqrt(var + layer.epsilon) * layer.gamma + layer.beta
        return x


#######################################################################################################################
## LOSSES #############################################################################################################
#######################################################################################################################
_EPS = 1e-7


def categorical_crossentropy(y_true, y_pred, from_logits: bool = False, **_):
    """ Keras semantics: probabilities are re-normalised, clipped to [1e-7, 1-1e-7]; one value per row """
    if from_logits:
        return -(y_true * torch.log_softmax(y_pred, dim=-1)).sum(dim=-1)
    y_pred = y_pred / y_pred.sum(dim=-1, keepdim=True)
    y_pred = torch.clamp(y_pred, _EPS, 1.0 - _EPS)
    return -(y_true * torch.log(y_pred)).sum(dim=-1)


def binary_crossentropy(y_true, y_pred, from_logits: bool = False, **_):
    if from_logits:
        return torch.nn.functional.binary_cross_entropy_with_logits(y_pred, y_true, reduction='none').mean(dim=-1)
    y_pred = torch.clamp(y_pred, _EPS, 1.0 - _EPS)
    return -(y_true * torch.log(y_pred) + (1 - y_true) * torch.log(1 - y_pred)).mean(dim=-1)


def mean_squared_error(y_true, y_pred, **_): return ((y_pred - y_true) ** 2).mean(dim=-1)


def mean_absolute_error(y_true, y_pred, **_): return (y_pred - y_true).abs().mean(dim=-1)


class losses:
    """ namespace mirroring ``tf.keras.losses`` """
    categorical_crossentropy = staticmethod(categorical_crossentropy)
    binary_crossentropy = staticmethod(binary_crossentropy)
    mean_squared_error = staticmethod(mean_squared_error)
    mean_absolute_error = staticmethod(mean_absolute_error)
    mse = staticmethod(mean_squared_error)
    mae = staticmethod(mean_absolute_error)

    @staticmethod
    def serialize(fn) -> str: return fn.__name__

    @staticmethod
    def deserialize(name: str): return getattr(losses, name)


#######################################################################################################################
## OPTIMIZERS #########################################################################################################
#######################################################################################################################
class Optimizer:
    def get_config(self) -> dict: raise NotImplementedError

    def apply_gradients(self, grads_and_vars) -> None: raise NotImplementedError

    @classmethod
    def from_config(cls, config): return cls(**config)


class Adam(Optimizer):
    """ Keras Adam: ``lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)``; ``theta -= lr_t * m / (sqrt(v) + eps)``, eps = 1e-7 """

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7,
                 **_):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots: dict[int, tuple[torch.Tensor, torch.Tensor]] = dict()
        self.capturable = False            # True: nothing of a step depends on host state (CUDA-graph capture of training_step)
        self._t_dev: Optional[torch.Tensor] = None

    def get_config(self):
        return dict(learning_rate=self.learning_rate, beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon)

    @torch.no_grad()
    def apply_gradients(self, grads_and_vars) -> None:
        self.iterations += 1
        grads, params, ms, vs = [], [], [], []
        for g, p in grads_and_vars:
            if g is None: continue
            if id(p) not in self._slots: self._slots[id(p)] = (torch.zeros_like(p), torch.zeros_like(p))
            m, v = self._slots[id(p)]
            grads.append(g.to(p.dtype)), params.append(p), ms.append(m), vs.append(v)
        if not params: return
        torch._foreach_mul_(ms, self.beta_1)
        torch._foreach_add_(ms, grads, alpha=1.0 - self.beta_1)
        torch._foreach_mul_(vs, self.beta_2)
        torch._foreach_addcmul_(vs, grads, grads, value=1.0 - self.beta_2)
        denom = torch._foreach_sqrt(vs)
        torch._foreach_add_(denom, self.epsilon)
        if self.capturable and params[0].is_cuda:
            # step counter and bias correction on the device: the same captured CUDA graph serves every replay
            # (BaseClass.training_step(graph=True)); float64 so that lr_t equals the host formula to the last float32 bit
            if self._t_dev is None or self._t_dev.device != params[0].device:
                self._t_dev = torch.full((), float(self.iterations - 1), dtype=torch.float64, device=params[0].device)
            self._t_dev += 1.0
            lr_t = self.learning_rate * torch.sqrt(1.0 - self.beta_2 ** self._t_dev) / (1.0 - self.beta_1 ** self._t_dev)
            step = torch._foreach_div(ms, denom)
            torch._foreach_mul_(step, lr_t.to(torch.float32))
            torch._foreach_sub_(params, step)
        else:
            t = self.iterations
            lr_t = self.learning_rate * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
            torch._foreach_addcdiv_(params, ms, denom, value=-lr_t)
            if self._t_dev is not None: self._t_dev.fill_(float(t))


class SGD(Optimizer):
    def __init__(self, learning_rate: float = 0.01, momentum: float = 0.0, **_):
        self.learning_rate, self.momentum = learning_rate, momentum
        self._slots: dict[int, torch.Tensor] = dict()

    def get_config(self): return dict(learning_rate=self.learning_rate, momentum=self.momentum)

    @torch.no_grad()
    def apply_gradients(self, grads_and_vars) -> None:
        for g, p in grads_and_vars:
            if g is None: continue
            if self.momentum:
                buf = self._slots.setdefault(id(p), torch.zeros_like(p))
                buf.mul_(self.momentum).add_(g, alpha=-self.learning_rate)
                p.add_(buf)
            else:
                p.add_(g, alpha=-self.learning_rate)


class optimizers:
    """ namespace mirroring ``tf.keras.optimizers`` """
    Adam = Adam
    SGD = SGD

    @staticmethod
    def serialize(opt) -> dict: return {'class_name': type(opt).__name__, 'config': opt.get_config()}

    @staticmethod
    def deserialize(config: dict): return getattr(optimizers, config['class_name'])(**config['config'])
