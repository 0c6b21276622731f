"""TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED DeepRank2 model files so their own arithmetic can be executed to
(a) validate ``oracle/restate.py``, (b) generate the golden vectors under ``tests/golden/``
(``oracle/make_golden.py``) and (c) serve as ``bench.py``'s reference arm / ``cpu_baseline``
(``kind: "reference"``).  The files come from ``/root/reference`` in the build container, or
from ``oracle/_ref/`` -- the reference package installed there by ``oracle/install_reference.py``
(``pip install --no-deps --target oracle/_ref``; git-ignored, travels to the GPU box with the
snapshot).

``import deeprank2`` itself cannot work here (torch_geometric, torch_scatter,
h5py, markov_clustering, community, matplotlib are not installed and there is no
network), so the nine third-party symbols the model files import are provided
by ``oracle/thirdparty.py`` through ``sys.modules`` and the five model files are
then loaded by path.  ``/root/reference`` does not exist on the GPU box: there only the
``oracle/_ref`` install is available (``bench.py``'s reference legs); the ``-m gpu`` tests and
``smoke()`` never call this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """First existing of: $DRK_REFERENCE_ROOT, /root/reference (build container), oracle/_ref (installed copy)."""
    for cand in (os.environ.get("DRK_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "deeprank2", "neuralnets", "gnn")):
            return cand
    return os.environ.get("DRK_REFERENCE_ROOT") or "/root/reference"


REFERENCE_ROOT = _find_root()

_STUBBED = False


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deeprank2", "neuralnets", "gnn"))


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install_thirdparty_stubs() -> None:
    """Register pure-torch stand-ins for the absent wheels (idempotent)."""
    global _STUBBED
    if _STUBBED:
        return
    from oracle import thirdparty as tp

    if "torch_scatter" not in sys.modules:
        _module(
            "torch_scatter",
            scatter_sum=tp.scatter_sum,
            scatter_add=tp.scatter_add,
            scatter_mean=tp.scatter_mean,
            scatter_max=tp.scatter_max,
        )
    if "torch_geometric" not in sys.modules:
        tg = _module("torch_geometric")
        tg.nn = _module("torch_geometric.nn", max_pool_x=tp.max_pool_x)
        tg.nn.inits = _module("torch_geometric.nn.inits", uniform=tp.uniform)
        tg.nn.pool = _module("torch_geometric.nn.pool")
        tg.nn.pool.consecutive = _module("torch_geometric.nn.pool.consecutive", consecutive_cluster=tp.consecutive_cluster)
        tg.nn.pool.pool = _module("torch_geometric.nn.pool.pool", pool_batch=tp.pool_batch, pool_edge=tp.pool_edge)
        tg.data = _module("torch_geometric.data", Batch=tp.Batch, Data=tp.Data)
    # imported at module scope by utils/community_pooling.py:3-5 but only used by the
    # (out-of-scope) community *detection* functions
    for name in ("community", "markov_clustering"):
        if name not in sys.modules:
            _module(name)
    if "matplotlib" not in sys.modules:
        mpl = _module("matplotlib")
        mpl.pyplot = _module("matplotlib.pyplot")
    _STUBBED = True


def _load_by_path(qualname: str, relpath: str) -> types.ModuleType:
    if qualname in sys.modules:
        return sys.modules[qualname]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(qualname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[qualname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Return a namespace with the reference's own classes, executed from /root/reference.

    Package ``__init__`` files are bypassed (they pull h5py etc.); bare namespace
    packages named ``deeprank2...`` are registered so that the intra-package import at
    ``ginet.py:8`` (``from deeprank2.utils.community_pooling import ...``) resolves.
    """
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (it only exists in the build container)")
    install_thirdparty_stubs()
    for pkg in ("deeprank2", "deeprank2.utils", "deeprank2.neuralnets", "deeprank2.neuralnets.gnn"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []  # namespace-like; children are loaded explicitly below
            sys.modules[pkg] = m
    cp = _load_by_path("deeprank2.utils.community_pooling", "deeprank2/utils/community_pooling.py")
    sys.modules["deeprank2.utils"].community_pooling = cp
    ns = types.SimpleNamespace(community_pooling=cp)
    for name in ("ginet", "ginet_nocluster", "foutnet", "sgat", "vanilla_gnn"):
        mod = _load_by_path(f"deeprank2.neuralnets.gnn.{name}", f"deeprank2/neuralnets/gnn/{name}.py")
        setattr(ns, name, mod)
    return ns
