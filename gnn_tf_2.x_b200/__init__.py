"""B200-native Scarselli-GNN state-convergence engine behind the API of sailab-code/GNN_tf_2.x.

Import as ``gnn_b200`` (see ``gnn_b200.py`` at the repository root: the directory name contains a dot).
Module names follow the reference package ``GNN`` so that ``from GNN.X import Y`` becomes ``from gnn_b200.X import Y``.
The hot path (state loop forward/backward, CSR build, prologue) runs in ``csrc/`` behind the C ABI declared in
``include/gnn_b200.h``; there is no CPU fallback.
"""
from .graph_class import GraphObject, GraphTensor, SparseCSR  # noqa: F401
from .MLP import MLP, get_inout_dims  # noqa: F401

__all__ = ['GraphObject', 'GraphTensor', 'SparseCSR', 'MLP', 'get_inout_dims']
