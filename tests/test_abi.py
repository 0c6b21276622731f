"""CPU: the C-ABI shared library loads and exports every symbol include/drk_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
from __future__ import annotations

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "drk_b200.h")).read()
    return sorted(set(re.findall(r"DRK_API\s+[\w\s\*]+?\b(drk_\w+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared_symbols()
    for must in ("drk_graph_index_build", "drk_batch_offsets", "drk_node_linear", "drk_weight_grad", "drk_spmm", "drk_segment_mean", "drk_segment_mean_bwd"):
        assert must in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    from deeprank2_b200 import _lib

    _lib.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/drk_b200.h but not exported"
    loaded = _lib.load()
    assert loaded.drk_abi_version() == _lib.ABI_VERSION
    assert sorted(_lib.SIGNATURES) == _declared_symbols(), "ctypes signature table and header disagree"


def test_pure_host_entry_points():
    from deeprank2_b200 import _lib

    lib = _lib.load()
    assert lib.drk_graph_index_workspace_bytes(1000, 100) > 0
    assert lib.drk_weight_grad_workspace_bytes(50, 32) > 0
    assert lib.drk_launch_count() >= 0
    # argument validation happens before any CUDA call
    rc = lib.drk_spmm(None, None, None, None, 0, None, 0, None, 0, None, 0, 5, 4, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.drk_last_error()
    rc = lib.drk_node_linear(None, 0, None, 0, 1, None, None, 0, None, 0, -1, 1, 1, 0, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    import torch

    from deeprank2_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.node_linear(torch.zeros(2, 2), torch.zeros(2, 2))
