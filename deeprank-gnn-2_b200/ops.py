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


class GINetStackFunction(torch.autograd.Function):
    """Both branches of the no-cluster ``GINet`` up to and including the readout, as ONE autograd node.

    ``ginet_nocluster.py:88-106``: conv1/conv1_ext (F->16) and conv2/conv2_ext (16->32) act on the same
    graph, so the two branches are stacked along the feature axis: one 32-wide projection of ``x``, two
    32-wide aggregations, two 16->32 projections writing the halves of one [N,64] tensor, one readout.
    Halves the number of launches and doubles the bytes each aggregation moves per index read.

      fwd:  P = x [W1;W1e]^T -> H1 = relu(A P) -> A2 = A H1 -> H2 = relu([A2a W2^T | A2b W2e^T]) -> G = mean_g(H2)
      bwd:  dZ2 = dG[batch]/n_g * (H2>0);  dW2 = dZ2a^T A2a, dW2e = dZ2b^T A2b;  dA2 = [dZ2a W2 | dZ2b W2e]
            dZ1 = (A^T dA2) * (H1>0);  Q = A^T dZ1;  d[W1;W1e] = Q^T x
    The eight attention parameters only receive exact-zero gradients (alpha == 1).
    """

    @staticmethod
    def forward(ctx, x, w1, w1e, w2, w2e, graph: GraphIndex, *dead):
        n = x.shape[0]
        f1 = w1.shape[0]          # 16
        f2 = w2.shape[0]          # 32
        w1s = torch.cat([w1, w1e], dim=0)
        p = node_linear(x, w1s, True)
        h1 = spmm(graph.rowptr, graph.colidx, p, n, act=ACT_RELU)
        a2 = spmm(graph.rowptr, graph.colidx, h1, n)
        h2 = torch.empty((n, 2 * f2), dtype=torch.float32, device=x.device)
        node_linear(a2[:, :f1], w2, True, act=ACT_RELU, out=h2[:, :f2])
        node_linear(a2[:, f1:], w2e, True, act=ACT_RELU, out=h2[:, f2:])
        g = segment_mean(h2, graph.graph_ptr, graph.num_graphs)
        ctx.graph = graph
        ctx.dead = dead
        ctx.f1, ctx.f2 = f1, f2
        ctx.save_for_backward(x, w1s, w2, w2e, h1, a2, h2)
        return g

    @staticmethod
    def backward(ctx, dg):
        x, w1s, w2, w2e, h1, a2, h2 = ctx.saved_tensors
        g = ctx.graph
        f1, f2 = ctx.f1, ctx.f2
        n = x.shape[0]
        dz2 = segment_mean_bwd(dg.contiguous(), g.graph_ptr, g.batch32, n, mask=h2)
        dw2 = weight_grad(dz2[:, :f2], a2[:, :f1])
        dw2e = weight_grad(dz2[:, f2:], a2[:, f1:])
        da2 = torch.empty_like(a2)
        node_linear(dz2[:, :f2], w2, False, out=da2[:, :f1])
        node_linear(dz2[:, f2:], w2e, False, out=da2[:, f1:])
        dz1 = spmm(g.colptr, g.rowidx, da2, n, mask=h1)
        q = spmm(g.colptr, g.rowidx, dz1, n)
        dw1s = weight_grad(q, x)
        dx = node_linear(q, w1s, False) if ctx.needs_input_grad[0] else None
        dead = tuple(torch.zeros_like(t) if ctx.needs_input_grad[6 + i] else None for i, t in enumerate(ctx.dead))
        return (dx, dw1s[:f1], dw1s[f1:], dw2, dw2e, None) + dead


def ginet_stack(x, conv1, conv1_ext, conv2, conv2_ext, graph):
    dead = []
    for layer in (conv1, conv2, conv1_ext, conv2_ext):
        dead += [layer.fc_edge_attr.weight, layer.fc_attention.weight]
    return GINetStackFunction.apply(x, conv1.fc.weight, conv1_ext.fc.weight, conv2.fc.weight, conv2_ext.fc.weight, graph, *dead)
