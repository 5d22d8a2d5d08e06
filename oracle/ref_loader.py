"""Load the reference's own hot-path sources, unmodified, from /root/reference.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only in the build
container (``/root/reference`` does not exist on the GPU box); its single job is
to pin ``oracle/intree.py`` / ``oracle/graph.py`` and to generate the committed
golden vectors (``tests/golden/make_golden.py``).

``import analysisgnn`` is never executed: the package ``__init__`` drags in
partitura / torch_geometric / graphmuse, none of which exist here.  Instead

* ``analysisgnn/models/core/gnn.py`` and ``hgnn.py`` are loaded by file path
  under stub parent packages, with ``oracle.scatter_shim`` registered as
  ``torch_scatter``;
* the body of ``hetero_graph_from_note_array`` (``analysisgnn/utils/hgraph.py:214``)
  and of ``HeteroScoreGraph.add_beat_nodes / add_measure_nodes`` (``:41-73``) are
  cut out of the file with ``ast`` and exec'd with numpy only;
* ``onsetwise_logit_aggregation`` (``analysisgnn/models/analysis.py:44-101``) is cut
  out the same way and exec'd with torch + the ``torch_scatter`` shim.
"""
import ast
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AGNN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "analysisgnn/models/core/gnn.py"))


_cache = {}


def load_core():
    """Returns (gnn_module, hgnn_module) = the reference's in-tree layer files."""
    if "core" in _cache:
        return _cache["core"]
    if not available():
        raise FileNotFoundError(f"reference sources not found under {REFERENCE_ROOT}")
    from . import scatter_shim

    saved = {k: sys.modules.get(k) for k in ("torch_scatter",)}
    sys.modules["torch_scatter"] = scatter_shim
    try:
        for name in ("analysisgnn", "analysisgnn.models", "analysisgnn.models.core"):
            if name not in sys.modules:
                pkg = types.ModuleType(name)
                pkg.__path__ = []  # mark as package, never searched
                sys.modules[name] = pkg
        mods = []
        for stem in ("gnn", "hgnn"):
            full = f"analysisgnn.models.core.{stem}"
            path = os.path.join(REFERENCE_ROOT, "analysisgnn/models/core", stem + ".py")
            spec = importlib.util.spec_from_file_location(full, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            mods.append(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["core"] = tuple(mods)
    return _cache["core"]


def _hgraph_tree():
    if "hgraph_tree" not in _cache:
        path = os.path.join(REFERENCE_ROOT, "analysisgnn/utils/hgraph.py")
        with open(path) as fh:
            _cache["hgraph_tree"] = ast.parse(fh.read(), filename=path)
    return _cache["hgraph_tree"]


def _exec_function(node, extra=None):
    import numpy as np
    import numpy.lib.recfunctions as rfn
    import warnings

    node.decorator_list = []
    module = ast.Module(body=[node], type_ignores=[])
    ast.fix_missing_locations(module)
    scope = {"np": np, "rfn": rfn, "warnings": warnings}
    scope.update(extra or {})
    exec(compile(module, "<reference hgraph.py>", "exec"), scope)
    return scope[node.name]


def load_edge_builder():
    """The reference's ``hetero_graph_from_note_array`` (utils/hgraph.py:214-300)."""
    if "edges" not in _cache:
        for node in _hgraph_tree().body:
            if isinstance(node, ast.FunctionDef) and node.name == "hetero_graph_from_note_array":
                _cache["edges"] = _exec_function(node)
                break
        else:
            raise LookupError("hetero_graph_from_note_array not found in the reference")
    return _cache["edges"]


def load_metrical_edge_builders():
    """(add_beat_nodes, add_measure_nodes) of the reference's HeteroScoreGraph
    (utils/hgraph.py:61-73, :41-59) as free functions taking a namespace with
    ``note_array`` (and ``name``) and setting ``*_nodes`` / ``*_edges`` on it."""
    if "metrical" not in _cache:
        found = {}
        for node in _hgraph_tree().body:
            if isinstance(node, ast.ClassDef) and node.name == "HeteroScoreGraph":
                for item in node.body:
                    if isinstance(item, ast.FunctionDef) and item.name in ("add_beat_nodes", "add_measure_nodes"):
                        found[item.name] = _exec_function(item)
        _cache["metrical"] = (found["add_beat_nodes"], found["add_measure_nodes"])
    return _cache["metrical"]


def load_onsetwise_decode():
    """The reference's ``onsetwise_logit_aggregation`` (models/analysis.py:44-101), unmodified."""
    if "decode" not in _cache:
        import torch
        from . import scatter_shim
        path = os.path.join(REFERENCE_ROOT, "analysisgnn/models/analysis.py")
        with open(path) as fh:
            tree = ast.parse(fh.read(), filename=path)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name == "onsetwise_logit_aggregation":
                node.decorator_list = []
                module = ast.Module(body=[node], type_ignores=[])
                ast.fix_missing_locations(module)
                scope = {"torch": torch, "torch_scatter": scatter_shim}
                exec(compile(module, "<reference models/analysis.py>", "exec"), scope)
                _cache["decode"] = scope[node.name]
                break
        else:
            raise LookupError("onsetwise_logit_aggregation not found in the reference")
    return _cache["decode"]


def load_multitask_loss():
    """The reference's ``MultiTaskLoss`` class (models/chord.py:16-49), unmodified (the module imports graphmuse /
    partitura at the top, so only the class definition is taken)."""
    if "mtl" not in _cache:
        import torch
        import torch.nn as nn
        path = os.path.join(REFERENCE_ROOT, "analysisgnn/models/chord.py")
        with open(path) as fh:
            tree = ast.parse(fh.read(), filename=path)
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name == "MultiTaskLoss":
                module = ast.Module(body=[node], type_ignores=[])
                ast.fix_missing_locations(module)
                scope = {"torch": torch, "nn": nn}
                exec(compile(module, "<reference models/chord.py>", "exec"), scope)
                _cache["mtl"] = scope[node.name]
                break
        else:
            raise LookupError("MultiTaskLoss not found in the reference")
    return _cache["mtl"]

