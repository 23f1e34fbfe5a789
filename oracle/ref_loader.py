"""TEST INFRASTRUCTURE ONLY -- loader for the *real* PPNet reference (read-only tree).

Only `tests/golden/make_golden.py` and `oracle/validate_against_ref.py` use this, and only in
the build container where the reference tree exists.  Nothing under `ppnet_b200/`, `bench.py`
or the `-m gpu` tests may import it (the GPU box has no reference tree).

The reference (EDaGe-PP/*.py, experiments/MPNet/neuralplanner.py) imports `matplotlib` and
`imgviz` at module top level; neither is installed here.  We inject inert stub modules so that
the numerical code runs unmodified.  `plot_obstacles` (matplotlib -> jpg -> PIL dither) cannot
run, so it is monkey-patched to a white tensor by `load_edage()` (raster parity is unpinned,
see DESIGN.md).

`neuralplanner.py` loads model weights from hard-coded paths at import time
(experiments/MPNet/neuralplanner.py:21-32), so its checker functions are lifted out with `ast`
and executed in a namespace that provides the globals they read (`clearance`, `obc`).
"""
import ast
import importlib
import os
import sys
import types


def find_reference():
    for cand in (os.environ.get("PPNET_REF"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "EDaGe-PP")):
            return cand
    return None


class _Stub(types.ModuleType):
    """Module whose every (non-dunder) attribute is a callable no-op object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()


class _Inert:
    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __iter__(self):
        return iter(())


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm",
                 "matplotlib.ticker", "imgviz"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "patches", "cm", "ticker"):
        setattr(mpl, sub, sys.modules["matplotlib." + sub])


_EDAGE = {}


def load_edage():
    """Import the reference's EDaGe-PP modules unmodified.  Returns a dict of modules."""
    if _EDAGE:
        return _EDAGE
    ref = find_reference()
    if ref is None:
        raise RuntimeError("PPNet reference tree not found (set $PPNET_REF)")
    _install_stubs()
    sys.path.insert(0, os.path.join(ref, "EDaGe-PP"))
    try:
        for name in ("PathSeg", "Path", "GMM", "process_map", "PathGenerate", "MapGenerate"):
            _EDAGE[name] = importlib.import_module(name)
    finally:
        sys.path.pop(0)

    import torch

    def _white(size, obstacles, resolution=(224, 224)):
        return torch.ones([3, resolution[0], resolution[1]])

    _EDAGE["Path"].plot_obstacles = _white
    _EDAGE["MapGenerate"].plot_obstacles = _white
    return _EDAGE


def load_mpnet_checker(obc, clearance=1 / 50 * 224):
    """AST-lift collision_check_circle_edge / steerTo / feasibility_check / lvc from
    experiments/MPNet/neuralplanner.py:43-138 and bind them to the given `obc` global."""
    ref = find_reference()
    if ref is None:
        raise RuntimeError("PPNet reference tree not found (set $PPNET_REF)")
    import numpy as np
    import torch
    from scipy.spatial import distance

    path = os.path.join(ref, "experiments", "MPNet", "neuralplanner.py")
    with open(path, "r", encoding="utf-8") as f:
        tree = ast.parse(f.read())
    wanted = {"collision_check_circle_edge", "steerTo", "feasibility_check", "lvc",
              "IsInCollision_circle"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    mod = ast.Module(body=body, type_ignores=[])
    ns = {"torch": torch, "np": np, "distance": distance, "clearance": clearance, "obc": obc}
    exec(compile(mod, path, "exec"), ns)
    return ns
