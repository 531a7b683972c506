# coding=utf-8
"""MLP factory with the reference's signature (``GNN/MLP.py:11-13,68-70``).

``MLP(...)`` returns a :class:`keras_compat.Sequential` (Dense / Dropout chain with an optional trailing
BatchNormalization) whose description the CUDA kernels evaluate; ``get_inout_dims`` restates the dimension rules of
``GNN/MLP.py:68-122``.
"""
from __future__ import annotations

from typing import Optional, Union

from .keras_compat import Dense, Dropout, AlphaDropout, BatchNormalization, Sequential


def _per_layer(value, n: int) -> list:
    return list(value) if type(value) == list else [value] * n


# ---------------------------------------------------------------------------------------------------------------------
def MLP(input_dim: int, layers: list[int], activations, kernel_initializer, bias_initializer,
        kernel_regularizer=None, bias_regularizer=None, dropout_rate: Union[list[float], float, None] = None,
        dropout_pos: Optional[Union[list[int], int]] = None, alphadropout: bool = False, batch_normalization: bool = True,
        *, device=None, seed: Optional[int] = None) -> Sequential:
    """ Quick building function for MLP model (MLP.py:11-64).

    :param input_dim: (int) input dimension of the model
    :param layers: (list of int) units of every Dense layer
    :param activations: activation or list of activations (Keras strings)
    :param kernel_initializer / bias_initializer: initializer or list of initializers (Keras strings)
    :param kernel_regularizer / bias_regularizer: callables or None (or lists), applied in training as extra loss
    :param dropout_rate: float or list of floats in [0, 1]
    :param dropout_pos: int or list of int: position of every dropout layer in the Dense list (0 = before first Dense)
    :param alphadropout: (bool) AlphaDropout instead of Dropout
    :param batch_normalization: (bool) append a BatchNormalization layer after the last Dense. Default True, as reference
    :return: Sequential model
    """
    layers = [layers] if type(layers) == int else list(layers)
    n = len(layers)
    if dropout_rate is None or dropout_pos is None: dropout_rate, dropout_pos = [], []
    if type(dropout_pos) == int: dropout_pos = [dropout_pos]
    if type(dropout_rate) == float: dropout_rate = [dropout_rate] * len(dropout_pos)

    per_dense = [_per_layer(v, n) for v in (activations, kernel_initializer, bias_initializer, kernel_regularizer, bias_regularizer)]
    if any(len(v) != n for v in per_dense):
        raise ValueError('Dense parameters must have the same length to be correctly processed')
    if len(dropout_rate) != len(dropout_pos):
        raise ValueError('Dropout parameters must have the same length to be correctly processed')

    # Dense chain, then dropout layers inserted in front of Dense number <pos> (pos == n: behind the last Dense)
    chain = [Dense(units=u, activation=a, kernel_initializer=ki, bias_initializer=bi, kernel_regularizer=kr, bias_regularizer=br)
             for u, a, ki, bi, kr, br in zip(layers, *per_dense)]
    drop_cls = AlphaDropout if alphadropout else Dropout
    for shift, (pos, rate) in enumerate(zip(dropout_pos, dropout_rate)):
        chain.insert(int(pos) + shift, drop_cls(rate=rate))
    if batch_normalization: chain.append(BatchNormalization())
    return Sequential(chain, input_dim=input_dim, device=device, seed=seed)


# ---------------------------------------------------------------------------------------------------------------------
def get_inout_dims(net_name: str, dim_node_label: int, dim_arc_label: int, dim_target: int, problem_based: str, dim_state: int,
                   hidden_units: Union[None, int, list[int]],
                   *, layer: int = 0, get_state: bool = False, get_output: bool = False) -> tuple[int, list[int]]:
    """ Input dimension and layer list for the state / output MLP (MLP.py:68-122).

    state net : input = AL + 2 * (NL + DS), output = DS if DS else NL
    output net: input = NL + DS (+ NL + AL + DS for arc-based problems), output = T
    For LGNN layers > 0 the label dims grow by the state and/or output of the previous layer (MLP.py:93-100).
    """
    assert layer >= 0
    assert problem_based in ['a', 'n', 'g']
    assert dim_state >= 0
    NL, AL, T, DS = dim_node_label, dim_arc_label, dim_target, dim_state
    arc_based = problem_based == 'a'

    if layer > 0:
        out_on_nodes = T * (not arc_based) * get_output
        if DS != 0:
            NL = NL + DS * get_state + out_on_nodes
        else:
            NL = NL + layer * NL * get_state + ((layer - 1) * get_state + 1) * out_on_nodes
        AL = AL + T * arc_based * get_output

    if net_name == 'state':
        input_shape, output_shape = AL + 2 * (NL + DS), (DS if DS else NL)
    elif net_name == 'output':
        input_shape, output_shape = arc_based * (NL + AL + DS) + NL + DS, T
    else:
        raise ValueError(':param net_name: not in [\'state\', \'output\']')

    if hidden_units is None or (type(hidden_units) == int and hidden_units <= 0): hidden_units = []
    hidden = hidden_units if type(hidden_units) == list else [hidden_units]
    return input_shape, hidden + [output_shape]
