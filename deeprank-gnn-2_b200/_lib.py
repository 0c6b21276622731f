"""ctypes binding of ``libdrk_b200.so`` (the C ABI declared in ``include/drk_b200.h``).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).  There is
no fallback: if it is missing, ``load()`` raises, and every op in this package calls
``load()``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRK_B200_LIB") or os.path.join(PKG_DIR, "libdrk_b200.so")  # override: A/B builds of the kernels
CSRC_DIR = os.path.join(PKG_DIR, "csrc")

# constants mirrored from include/drk_b200.h
ABI_VERSION = 1
OK = 0
STATUS_INDEX_RANGE = 1
STATUS_CROSS_GRAPH = 2
STATUS_UNSORTED = 4
ACT_NONE, ACT_RELU = 0, 1
REDUCE_SUM, REDUCE_MEAN_CLAMP, REDUCE_MEAN_NAN = 0, 1, 2
LOSS_MSE, LOSS_CROSS_ENTROPY = 0, 1
EDGES_DIRECTED, EDGES_UNDIRECTED_PAIRS, EDGES_LOCAL_PAIRS16 = 0, 1, 2
POOL_JUNK_SEGMENTS = 4096

_P = c_void_p
_I32 = c_int32
_I64 = c_int64

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
SIGNATURES = {
    "drk_abi_version": (c_int32, []),
    "drk_last_error": (c_char_p, []),
    "drk_launch_count": (c_int64, []),
    "drk_runtime_init": (c_int32, []),
    "drk_graph_index_workspace_bytes": (c_size_t, [_I64, _I32]),
    "drk_graph_index_build": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "drk_segment_index_workspace_bytes": (c_size_t, [_I64, _I32]),
    "drk_segment_index_build": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P, c_size_t, _P]),
    "drk_batch_offsets": (c_int32, [_P, _I32, _I32, _P, _P, _P, _P]),
    "drk_gather_rows": (c_int32, [_P, _I64, _P, _I64, _I32, _P, _I64, _P]),
    "drk_node_linear": (c_int32, [_P, _I64, _P, _I64, _I32, _P, _P, _I64, _P, _I64, _I64, _I32, _I32, _I32, _P]),
    "drk_node_linear2": (c_int32, [_P, _I64, _P, _I64, _I32, _P, _I64, _P, _I64, _I32, _I32, _P, _P, _I64, _P, _I64, _I64, _I32, _I32, _P]),
    "drk_weight_grad_workspace_bytes": (c_size_t, [_I32, _I32]),
    "drk_weight_grad": (c_int32, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P, _I64, _P, _I32, _P, c_size_t, _P]),
    "drk_spmm": (c_int32, [_P, _P, _P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _P]),
    "drk_spmm_tiled_supported": (c_int32, [_I32, _I32]),
    "drk_spmm_tiled": (c_int32, [_P, _P, _P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "drk_segment_mean": (c_int32, [_P, _I64, _P, _I32, _I32, _P, _I64, _P]),
    "drk_segment_mean_rows": (c_int32, [_P, _I64, _P, _I32, _I64, _I32, _P, _I64, _P]),
    "drk_segment_mean_bwd": (c_int32, [_P, _I64, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P]),
    "drk_ginet_fused_max_nodes": (c_int32, [_I32]),
    "drk_ginet_fused_fwd": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "drk_ginet_fused_bwd_workspace_bytes": (c_size_t, []),
    "drk_ginet_fused_bwd": (c_int32, [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, c_size_t, _P]),
    "drk_edge_ptr": (c_int32, [_P, _I64, _P, _I32, _P, _P]),
    "drk_graph_index_blocked_supported": (c_int32, [_I32, _I32]),
    "drk_graph_index_build_blocked": (c_int32, [_P, _I64, _I32, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "drk_ginet_step_ctas": (c_int32, [_I32]),
    "drk_ginet_step_set_phase_clocks": (c_int32, [_P, _I32]),
    "drk_ginet_step_exchange_floats": (c_int32, [_I32, _I32]),
    "drk_ginet_step_supported": (c_int32, [_I32, _I32, _I32, _I32]),
    "drk_ginet_step_workspace_bytes": (c_size_t, [_I32, _I32, _I32, _I32, _I32]),
    "drk_ginet_step": (c_int32, [_P, _I64, _I32, _P, _I64, _I32, _P, _P, _P, _I32, _I32, _I32, _I32,   # x .. edge_layout .. order, outputs_by_slot .. max_graph_edges
                                 _P, _P, _P, _P, _P, _P, _P, _P, _I32,                       # weights, out_dim
                                 _I32, _P, c_float, c_float, c_uint64, _P, _I32,             # loss, dropout, train
                                 _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,                     # pred, loss, 8 gradients
                                 _P, _P, _P, _P, c_size_t, _P]),                             # adam, peers, status, workspace, stream
    "drk_segment_max": (c_int32, [_P, _P, _P, _I64, _I32, _I32, _I32, _P, _I64, _P, _P]),
    "drk_segment_max_bwd": (c_int32, [_P, _I64, _P, _I32, _I32, _I32, _P, _I64, _P]),
    "drk_cluster_offsets_workspace_bytes": (c_size_t, [_I32]),
    "drk_cluster_offsets": (c_int32, [_P, _P, _P, _I32, _I32, _P, _P, c_size_t, _P]),
    "drk_compact_segments_workspace_bytes": (c_size_t, [_I32]),
    "drk_compact_segments": (c_int32, [_P, _I32, _P, _P, _P, _P, _P, _I32, _P, _P, _P, c_size_t, _P]),
    "drk_pool_edge_keys": (c_int32, [_P, _I64, _P, _I32, _P, _P, _P, _I32, _I64, _P, _P, _P]),
    "drk_pool_edge_decode": (c_int32, [_P, _I32, _P, _P, _P, _I32, _P, _P]),
    "drk_consecutive_blocked_supported": (c_int32, [_I32, _I32]),
    "drk_consecutive_blocked": (c_int32, [_P, _I32, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "drk_pool_edge_blocked_supported": (c_int32, [_I32, _I32]),
    "drk_pool_edge_blocked": (c_int32, [_P, _I64, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _P, _I64, _I32, _P, _I64, _P, _P, _P]),
    "drk_edge_msg_fwd": (c_int32, [_P, _P, _P, _P, _I64, _P, _I64, _I32, _P, _I64, _P, _I64, _P, _P, _I32, _P]),
    "drk_edge_msg_bwd_src": (c_int32, [_P, _P, _P, _P, _I64, _P, _P, _I64, _I32, _P]),
    "drk_edge_msg_bwd_c_workspace_bytes": (c_size_t, []),
    "drk_edge_msg_bwd_c": (c_int32, [_P, _P, _P, _I64, _P, _P, _I64, _I32, _P, _I64, _I32, _P, c_size_t, _P]),
    "drk_vanilla_layer_supported": (c_int32, [_I32, _I32, _I32]),
    "drk_vanilla_layer_fwd": (c_int32, [_P, _I32, _P, _P, _P, _I32, _P, _P, _I32, _I32, _P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "drk_vanilla_layer_bwd_workspace_bytes": (c_size_t, [_I32, _I32]),
    "drk_vanilla_layer_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _I32, _P, _I64, _P, _I64, _P, _P, _I64, _P,
                                        _P, _I64, _P, _P, _P, c_size_t, _P]),
    "drk_attn_supported": (c_int32, [_I32, _I32]),
    "drk_attn_fwd": (c_int32, [_P, _P, _P, _I64, _P, _P, _I64, _I32, _P, ctypes.c_float, _P, _I64, _P, _P, _I32, _I32, _I32, _P]),
    "drk_attn_bwd_dst": (c_int32, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _P, ctypes.c_float, _P, _P, _I64, _I32, _I32, _I32, _P]),
    "drk_attn_slot_map": (c_int32, [_P, _P, _I64, _P, _P, _P]),
    "drk_attn_bwd_src": (c_int32, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _P]),
    "drk_attn_edge_grad_workspace_bytes": (c_size_t, [_I32]),
    "drk_attn_edge_grad": (c_int32, [_P, _P, _I64, _I64, _I32, _P, _P, c_size_t, _P]),
}



class Peers(ctypes.Structure):
    """``DrkPeers`` of include/drk_b200.h"""

    _fields_ = [("world", c_int32), ("rank", c_int32), ("capacity", c_int64), ("flag_capacity", c_int64), ("grad_buf", c_void_p * 8), ("flags", c_void_p * 8)]


class AdamTensor(ctypes.Structure):
    """``DrkAdamTensor`` of include/drk_b200.h"""

    _fields_ = [("param", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("step", c_void_p), ("numel", c_int64)]


class Adam(ctypes.Structure):
    """``DrkAdam`` of include/drk_b200.h"""

    _fields_ = [("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("weight_decay", c_float), ("num_dead", c_int32),
                ("live", AdamTensor * 8), ("dead", AdamTensor * 8)]


_lib = None


class DrkError(RuntimeError):
    """A C-ABI call returned a negative DRK_E* code."""


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libdrk_b200.so failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C deeprank-gnn-2_b200/csrc`). "
            "deeprank2_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.drk_abi_version() != ABI_VERSION:
        raise ImportError(f"libdrk_b200.so has ABI {lib.drk_abi_version()}, this package expects {ABI_VERSION}: rebuild it")
    try:
        import torch

        if torch.cuda.is_available():
            torch.cuda.init()
            lib.drk_runtime_init()
    except ImportError:  # a host without torch binds the C ABI directly
        pass
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().drk_last_error()
        raise DrkError(f"{what or 'drk call'} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().drk_launch_count())
