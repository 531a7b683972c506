"""Import alias for the package directory ``gnn_tf_2.x_b200/``.

The directory name required by the repository layout contains a dot, which the
``import`` statement cannot spell.  ``import gnn_b200`` executes this file once;
it loads ``gnn_tf_2.x_b200/__init__.py`` as a regular package registered under the
name ``gnn_b200`` (so ``from gnn_b200.graph_class import GraphObject`` and relative
imports inside the package both work) and replaces itself in ``sys.modules``.
"""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gnn_tf_2.x_b200")
_spec = importlib.util.spec_from_file_location(
    "gnn_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gnn_b200"] = _mod
_spec.loader.exec_module(_mod)
