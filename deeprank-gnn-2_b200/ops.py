"""Tensor-level wrappers over the C ABI and the autograd Functions built from them.

Everything here runs on the current CUDA stream of the tensors' device through
``libdrk_b200.so``; nothing falls back to torch ops on a missing extension or on CPU
tensors (``_f32_cuda`` raises).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, REDUCE_MEAN_CLAMP, REDUCE_MEAN_NAN, REDUCE_SUM  # noqa: F401
from .graph import GraphIndex, stream_ptr, workspace


def _f32_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor, got {t.device} (deeprank2_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def _p(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------ plain ops
def node_linear(a, b, trans_b=True, bias=None, mask=None, act=ACT_NONE, out=None):
    """``act(a @ b.T + bias)`` (``trans_b``) or ``act(a @ b + bias)``; optional ReLU-backward mask."""
    lib = _lib.load()
    a = _f32_cuda(a, "a")
    b = _f32_cuda(b, "b")
    n, k = a.shape
    m = b.shape[0] if trans_b else b.shape[1]
    if (b.shape[1] if trans_b else b.shape[0]) != k:
        raise ValueError(f"node_linear: inner sizes differ: a {tuple(a.shape)}, b {tuple(b.shape)}, trans_b={trans_b}")
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=a.device)
    if mask is not None:
        mask = _f32_cuda(mask, "mask")
    if bias is not None:
        bias = bias.contiguous()
    with torch.cuda.device(a.device):
        rc = lib.drk_node_linear(_p(a), _ld(a), _p(b), _ld(b), 1 if trans_b else 0, _p(bias), _p(mask), _ld(mask) if mask is not None else 0,
                                 _p(out), _ld(out), n, k, m, act, stream_ptr())
    _lib.check(rc, "drk_node_linear")
    return out


def weight_grad(dy, x, want_bias=False, dw=None, dbias=None, accumulate=False):
    """``dW = dy.T @ x`` [M,K] (and ``dy.sum(0)``), two-stage fixed-order reduction."""
    lib = _lib.load()
    dy = _f32_cuda(dy, "dy")
    x = _f32_cuda(x, "x")
    n, m = dy.shape
    k = x.shape[1]
    if x.shape[0] != n:
        raise ValueError("weight_grad: row counts differ")
    if dw is None:
        dw = torch.empty((m, k), dtype=torch.float32, device=dy.device)
    if want_bias and dbias is None:
        dbias = torch.empty((m,), dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        ws_bytes = lib.drk_weight_grad_workspace_bytes(k, m)
        ws = workspace(ws_bytes, dy.device)
        rc = lib.drk_weight_grad(_p(dy), _ld(dy), _p(x), _ld(x), n, k, m, _p(dw), _ld(dw), _p(dbias) if want_bias else None,
                                 1 if accumulate else 0, _p(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_weight_grad")
    return (dw, dbias) if want_bias else dw


def spmm(ptr, idx, src, n_out, w=None, addend=None, mask=None, reduce=REDUCE_SUM, act=ACT_NONE, out=None):
    """``out[i] = epi(reduce_{s in [ptr[i], ptr[i+1])} w[s] * src[idx[s]])`` -- gather + segmented reduce."""
    lib = _lib.load()
    src = _f32_cuda(src, "src")
    width = src.shape[1]
    if out is None:
        out = torch.empty((n_out, width), dtype=torch.float32, device=src.device)
    if addend is not None:
        addend = _f32_cuda(addend, "addend")
    if mask is not None:
        mask = _f32_cuda(mask, "mask")
    with torch.cuda.device(src.device):
        rc = lib.drk_spmm(_p(ptr), _p(idx), _p(w), _p(src), _ld(src), _p(addend), _ld(addend) if addend is not None else 0,
                          _p(mask), _ld(mask) if mask is not None else 0, _p(out), _ld(out), n_out, width, reduce, act, stream_ptr())
    _lib.check(rc, "drk_spmm")
    return out


def segment_mean(x, graph_ptr, num_graphs):
    lib = _lib.load()
    x = _f32_cuda(x, "x")
    out = torch.empty((num_graphs, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.drk_segment_mean(_p(x), _ld(x), _p(graph_ptr), num_graphs, x.shape[1], _p(out), _ld(out), stream_ptr())
    _lib.check(rc, "drk_segment_mean")
    return out


def segment_mean_bwd(dg, graph_ptr, batch32, num_nodes, mask=None):
    lib = _lib.load()
    dg = _f32_cuda(dg, "dg")
    width = dg.shape[1]
    dx = torch.empty((num_nodes, width), dtype=torch.float32, device=dg.device)
    if mask is not None:
        mask = _f32_cuda(mask, "mask")
    with torch.cuda.device(dg.device):
        rc = lib.drk_segment_mean_bwd(_p(dg), _ld(dg), _p(graph_ptr), _p(batch32), _p(mask), _ld(mask) if mask is not None else 0,
                                      num_nodes, width, _p(dx), _ld(dx), stream_ptr())
    _lib.check(rc, "drk_segment_mean_bwd")
    return dx


def gather_rows(src, perm):
    lib = _lib.load()
    src = _f32_cuda(src, "src")
    n = int(perm.numel())
    out = torch.empty((n, src.shape[1]), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        rc = lib.drk_gather_rows(_p(src), _ld(src), _p(perm), n, src.shape[1], _p(out), _ld(out), stream_ptr())
    _lib.check(rc, "drk_gather_rows")
    return out


def segment_index(index: torch.Tensor, num_segments: int):
    """Stable counting sort of an int64 key vector -> (ptr int32 [S+1], perm int32 [n], status int32 [1])."""
    lib = _lib.load()
    if not index.is_cuda or index.dtype != torch.int64:
        raise TypeError("segment_index: expected an int64 CUDA tensor")
    index = index.contiguous()
    n = int(index.numel())
    dev = index.device
    ptr = torch.empty(num_segments + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(lib.drk_segment_index_workspace_bytes(n, num_segments), dev)
        rc = lib.drk_segment_index_build(_p(index), n, num_segments, _p(ptr), _p(perm), _p(status), _p(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_segment_index_build")
    return ptr, perm, status


# ------------------------------------------------------------------------------ autograd
class GINetConvFunction(torch.autograd.Function):
    """``z = [relu]( A (x W^T + b) )`` with ``A[i,j]`` = number of edges (row=i, col=j).

    This is what ``GINetConvLayer.forward`` computes (reference ``ginet.py:40-60``): its attention
    coefficient is a softmax over a singleton axis, i.e. 1 for every edge.  The two attention
    weights are inputs of this Function only so that they receive the same exact-zero gradient
    tensors the reference's autograd gives them.

    The contraction order is chosen by width: project first when ``Fo <= Fi`` (or a bias is
    present), aggregate first otherwise, so the gather always runs on the narrower tensor.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, dead_a, dead_b, graph: GraphIndex, relu: bool):
        fo, fi = weight.shape
        project_first = fo <= fi or bias is not None
        act = ACT_RELU if relu else ACT_NONE
        n = x.shape[0]
        if project_first:
            p = node_linear(x, weight, True, bias)
            z = spmm(graph.rowptr, graph.colidx, p, n, act=act)
            saved = x
        else:
            a = spmm(graph.rowptr, graph.colidx, x, n)
            z = node_linear(a, weight, True, None, act=act)
            saved = a
        ctx.graph = graph
        ctx.relu = relu
        ctx.project_first = project_first
        ctx.has_bias = bias is not None
        ctx.dead = (dead_a, dead_b)
        ctx.save_for_backward(saved, weight, z if relu else None)
        return z

    @staticmethod
    def backward(ctx, dz):
        saved, weight, z = ctx.saved_tensors
        g = ctx.graph
        if g.colptr is None:
            raise RuntimeError("backward needs the CSC half of the graph index (build it with with_csc=True)")
        dz = dz.contiguous()
        if ctx.relu:
            dz = torch.ops.aten.threshold_backward(dz, z, 0.0)
        n = saved.shape[0]
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = dw = db = None
        if ctx.project_first:
            dp = spmm(g.colptr, g.rowidx, dz, n)  # A^T dz
            if need_w or ctx.has_bias:
                if ctx.has_bias:
                    dw, db = weight_grad(dp, saved, want_bias=True)
                else:
                    dw = weight_grad(dp, saved)
            if need_x:
                dx = node_linear(dp, weight, False)
        else:
            if need_w:
                dw = weight_grad(dz, saved)
            if need_x:
                da = node_linear(dz, weight, False)
                dx = spmm(g.colptr, g.rowidx, da, n)
        dead = [torch.zeros_like(t) if (t is not None and ctx.needs_input_grad[3 + i]) else None for i, t in enumerate(ctx.dead)]
        return dx, dw, db, dead[0], dead[1], None, None


def ginet_conv(x, weight, graph, bias=None, relu=False, dead_params=(None, None)):
    return GINetConvFunction.apply(x, weight, bias, dead_params[0], dead_params[1], graph, relu)


class MeanReadoutFunction(torch.autograd.Function):
    """``scatter_mean(x, batch, dim=0)`` over the sorted batch vector (``ginet_nocluster.py:103``)."""

    @staticmethod
    def forward(ctx, x, graph: GraphIndex):
        ctx.graph = graph
        ctx.n = x.shape[0]
        return segment_mean(x, graph.graph_ptr, graph.num_graphs)

    @staticmethod
    def backward(ctx, dg):
        g = ctx.graph
        return segment_mean_bwd(dg.contiguous(), g.graph_ptr, g.batch32, ctx.n), None


def mean_readout(x, graph):
    if graph.graph_ptr is None:
        raise RuntimeError("graph index was built without a batch vector")
    return MeanReadoutFunction.apply(x, graph)
