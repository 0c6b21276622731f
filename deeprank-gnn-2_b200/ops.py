"""Tensor-level wrappers over the C ABI and the autograd Functions built from them.

Everything here runs on the current CUDA stream of the tensors' device through
``libdrk_b200.so``; nothing falls back to torch ops on a missing extension or on CPU
tensors (``_f32_cuda`` raises).
"""
from __future__ import annotations

import os

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


# drk_spmm_tiled is OPT-IN (DRK_SPMM_TILED=1): on B200 it measured slower than the L2-gather kernel on the C3 adjacency (59 vs 43 us at
# width 16, 107 vs 62 us at width 32, profiles/spmm_probe.py) -- one 1024-thread CTA per SM cannot hide the index stream's latency the way
# 32-64 resident warps of the generic kernel do, and L2 already delivers 5.7 TB/s of gathered rows.  Kept because it is bit-identical
# and the right shape once the index stream is staged too.
SPMM_TILED = bool(__import__("os").environ.get("DRK_SPMM_TILED"))
SPMM_TILE_MIN_NODES = int(__import__("os").environ.get("DRK_SPMM_TILE_MIN_NODES", "1024"))


def _tileable(t) -> bool:
    return t is None or (t.data_ptr() % 16 == 0 and _ld(t) % 4 == 0)


def spmm(ptr, idx, src, n_out, w=None, addend=None, mask=None, reduce=REDUCE_SUM, act=ACT_NONE, out=None, graph=None):
    """``out[i] = epi(reduce_{s in [ptr[i], ptr[i+1])} w[s] * src[idx[s]])`` -- gather + segmented reduce.

    ``graph`` (the :class:`GraphIndex` ``ptr`` / ``idx`` belong to) + ``DRK_SPMM_TILED=1``: for collated batches of LARGE graphs (atom
    level, >= 1024 nodes, <= 3584) the block-diagonal kernel ``drk_spmm_tiled`` stages every graph's source rows in shared memory
    instead of gathering through L2; same result bit for bit (opt-in: not faster yet, see above)."""
    lib = _lib.load()
    src = _f32_cuda(src, "src")
    width = src.shape[1]
    if out is None:
        out = torch.empty((n_out, width), dtype=torch.float32, device=src.device)
    if addend is not None:
        addend = _f32_cuda(addend, "addend")
    if mask is not None:
        mask = _f32_cuda(mask, "mask")
    big = getattr(graph, "max_graph_nodes", None) if graph is not None else None
    if (SPMM_TILED and big is not None and big >= SPMM_TILE_MIN_NODES and idx is not None and graph.graph_ptr is not None and n_out == graph.num_nodes == src.shape[0]
            and lib.drk_spmm_tiled_supported(int(big), width) and all(_tileable(t) for t in (src, out, addend, mask))):
        with torch.cuda.device(src.device):
            rc = lib.drk_spmm_tiled(_p(ptr), _p(idx), _p(w), _p(src), _ld(src), _p(addend), _ld(addend) if addend is not None else 0,
                                    _p(mask), _ld(mask) if mask is not None else 0, _p(out), _ld(out), _p(graph.graph_ptr), graph.num_graphs, int(big),
                                    width, reduce, act, stream_ptr())
        _lib.check(rc, "drk_spmm_tiled")
        return out
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
        # the row count (known on the host) lets the library give big graphs a cluster of CTAs each
        rc = lib.drk_segment_mean_rows(_p(x), _ld(x), _p(graph_ptr), num_graphs, int(x.shape[0]), x.shape[1], _p(out), _ld(out), stream_ptr())
    _lib.check(rc, "drk_segment_mean_rows")
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


def segment_index(index: torch.Tensor, num_segments: int, status: torch.Tensor | None = None):
    """Stable counting sort of an int64 key vector -> (ptr int32 [S+1], perm int32 [n], status int32 [1]).  ``status``: an existing
    status word to OR the faults into (saves the fill launch of a fresh one)."""
    lib = _lib.load()
    if not index.is_cuda or index.dtype != torch.int64:
        raise TypeError("segment_index: expected an int64 CUDA tensor")
    index = index.contiguous()
    n = int(index.numel())
    dev = index.device
    ptr = torch.empty(num_segments + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    if status is None:
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
            z = spmm(graph.rowptr, graph.colidx, p, n, act=act, graph=graph)
            saved = x
        else:
            a = spmm(graph.rowptr, graph.colidx, x, n, graph=graph)
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
            dp = spmm(g.colptr, g.rowidx, dz, n, graph=g)  # A^T dz
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
                dx = spmm(g.colptr, g.rowidx, da, n, graph=g)
        dead = [torch.zeros_like(t) if (t is not None and ctx.needs_input_grad[3 + i]) else None for i, t in enumerate(ctx.dead)]
        return dx, dw, db, dead[0], dead[1], None, None


def ginet_conv(x, weight, graph, bias=None, relu=False, dead_params=(None, None)):
    return GINetConvFunction.apply(x, weight, bias, dead_params[0], dead_params[1], graph, relu)


LEAKY_SLOPE = 0.01  # torch.nn.functional.leaky_relu default, ginet.py:52


class GINetAttentionConvFunction(torch.autograd.Function):
    """``GINetConvLayer`` with the attention normalised per destination node (``attention="segment_softmax"``).

    The logit is the reference's (``ginet.py:45-52``): ``q_e = fc_attention([fc(x)[row_e], fc(x)[col_e], fc_edge_attr(edge_attr_e)])``,
    ``leaky_relu``; the softmax then runs over the edges that share ``row_e`` instead of the reference's singleton axis
    (``ginet.py:54``), and ``z[i] = sum_e alpha_e fc(x)[col_e]``.  With ``fc_attention.weight = [a_r | a_c | a_e]`` the logit
    splits into two per-node scalars ``s = P [a_r a_c]^T`` and ``u . edge_attr_e`` with ``u = We^T a_e``, so nothing of size
    ``[E, F]`` is formed.  Oracle: ``oracle/restate.py:ginet_conv_segment_softmax`` (+ torch autograd).
    """

    @staticmethod
    def forward(ctx, x, edge_attr, weight, w_edge, w_att, graph: GraphIndex, relu: bool):
        lib = _lib.load()
        x = _f32_cuda(x, "x")
        fo = weight.shape[0]
        if edge_attr.dim() == 1:
            edge_attr = edge_attr.unsqueeze(-1)  # ginet.py:43
        edge_attr = _f32_cuda(edge_attr, "edge_attr")
        fe = edge_attr.shape[1]
        n, e = x.shape[0], graph.num_edges
        if edge_attr.shape[0] != e:
            raise ValueError(f"edge_attr has {edge_attr.shape[0]} rows for {e} edges")
        if tuple(w_att.shape) != (1, 2 * fo + fe) or tuple(w_edge.shape) != (fe, fe):
            raise ValueError(f"attention weights {tuple(w_att.shape)} / {tuple(w_edge.shape)} do not match {fo} channels and {fe} edge features")
        if not lib.drk_attn_supported(fo, fe):
            raise NotImplementedError(f"segment-softmax attention needs out_channels % 4 == 0, <= 128 and <= 32 edge features (got {fo}, {fe})")
        ctx.graph, ctx.relu, ctx.no_edges = graph, relu, e == 0
        if e == 0:  # scatter into zeros with nothing to scatter (ginet.py:57-58)
            ctx.save_for_backward(x, edge_attr, weight, w_edge, w_att)
            return torch.zeros((n, fo), dtype=torch.float32, device=x.device)
        p = node_linear(x, weight, True)
        att = w_att[0, : 2 * fo].reshape(2, fo).contiguous()
        s = node_linear(p, att, True)  # [n, 2]: a_r.P[i], a_c.P[i]
        a_e = w_att[0, 2 * fo :]
        u = (a_e @ w_edge).contiguous()  # u[l] = sum_k a_e[k] We[k,l]
        attr = graph.attr_in_slot_order(edge_attr)
        z = torch.empty((n, fo), dtype=torch.float32, device=x.device)
        adq = torch.empty((e, 2), dtype=torch.float32, device=x.device)  # per CSR slot: (signed alpha, dq)
        with torch.cuda.device(x.device):
            scratch = workspace(4 * e, x.device)
            rc = lib.drk_attn_fwd(_p(graph.rowptr), _p(graph.colidx), _p(p), _ld(p), _p(s), _p(attr), _ld(attr), fe, _p(u), LEAKY_SLOPE,
                                  _p(z), _ld(z), _p(adq), _p(scratch), n, fo, ACT_RELU if relu else ACT_NONE, stream_ptr())
        _lib.check(rc, "drk_attn_fwd")
        ctx.save_for_backward(x, attr, weight, w_edge, w_att, p, z, adq)
        return z

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        if ctx.no_edges:
            x, edge_attr, weight, w_edge, w_att = ctx.saved_tensors
            grads = [torch.zeros_like(t) if need else None for t, need in zip((x, edge_attr, weight, w_edge, w_att), ctx.needs_input_grad[:5])]
            grads[1] = None
            return (*grads, None, None)
        x, attr, weight, w_edge, w_att, p, z, adq = ctx.saved_tensors
        g = ctx.graph
        if g.colptr is None:
            raise RuntimeError("backward needs the CSC half of the graph index (build it with with_csc=True)")
        dy = _f32_cuda(dy, "grad_output")
        n, fo = z.shape
        fe = attr.shape[1]
        e = g.num_edges
        dev = x.device
        ds = torch.empty((n, 2), dtype=torch.float32, device=dev)
        dz = torch.empty_like(z) if ctx.relu else None
        dp = torch.empty_like(p)
        att = w_att[0, : 2 * fo].contiguous()
        slot_map = g.slot_map()
        with torch.cuda.device(dev):
            rc = lib.drk_attn_bwd_dst(_p(g.rowptr), _p(g.colidx), _p(p), _ld(p), _p(dy), _ld(dy), _p(z), _ld(z), _p(adq), LEAKY_SLOPE, _p(ds),
                                      _p(dz), _ld(dz) if dz is not None else 0, n, fo, ACT_RELU if ctx.relu else ACT_NONE, stream_ptr())
            _lib.check(rc, "drk_attn_bwd_dst")
            grad_rows = dz if ctx.relu else dy
            rc = lib.drk_attn_bwd_src(_p(g.colptr), _p(g.rowidx), _p(slot_map), _p(grad_rows), _ld(grad_rows), _p(adq), _p(ds), _p(att),
                                      _p(dp), _ld(dp), n, fo, stream_ptr())
            _lib.check(rc, "drk_attn_bwd_src")
        dx = dw = dwe = dwa = None
        if ctx.needs_input_grad[2]:
            dw = weight_grad(dp, x)
        if ctx.needs_input_grad[0]:
            dx = node_linear(dp, weight, False)
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            a_e = w_att[0, 2 * fo :]
            gsum = torch.empty(fe, dtype=torch.float32, device=dev)  # sum_e dq_e attr_e, fixed order
            with torch.cuda.device(dev):
                ws_bytes = lib.drk_attn_edge_grad_workspace_bytes(fe)
                ws = workspace(ws_bytes, dev)
                rc = lib.drk_attn_edge_grad(_p(adq), _p(attr), _ld(attr), e, fe, _p(gsum), _p(ws), ws.numel(), stream_ptr())
            _lib.check(rc, "drk_attn_edge_grad")
            if ctx.needs_input_grad[3]:
                dwe = torch.outer(a_e, gsum)  # q_e = a_e^T We attr_e
            if ctx.needs_input_grad[4]:
                d_rc = weight_grad(ds, p).reshape(1, 2 * fo)  # d[a_r | a_c] = ds^T P
                dwa = torch.cat([d_rc, (w_edge @ gsum).reshape(1, fe)], dim=1)
        return dx, None, dw, dwe, dwa, None, None


def ginet_attention_conv(x, edge_attr, weight, w_edge, w_att, graph, relu=False):
    return GINetAttentionConvFunction.apply(x, edge_attr, weight, w_edge, w_att, graph, relu)


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
        h1 = spmm(graph.rowptr, graph.colidx, p, n, act=ACT_RELU, graph=graph)
        a2 = spmm(graph.rowptr, graph.colidx, h1, n, graph=graph)
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
        dz1 = spmm(g.colptr, g.rowidx, da2, n, mask=h1, graph=g)
        q = spmm(g.colptr, g.rowidx, dz1, n, graph=g)
        dw1s = weight_grad(q, x)
        dx = node_linear(q, w1s, False) if ctx.needs_input_grad[0] else None
        dead = tuple(torch.zeros_like(t) if ctx.needs_input_grad[6 + i] else None for i, t in enumerate(ctx.dead))
        return (dx, dw1s[:f1], dw1s[f1:], dw2, dw2e, None) + dead


def ginet_stack(x, conv1, conv1_ext, conv2, conv2_ext, graph):
    dead = []
    for layer in (conv1, conv2, conv1_ext, conv2_ext):
        dead += [layer.fc_edge_attr.weight, layer.fc_attention.weight]
    return GINetStackFunction.apply(x, conv1.fc.weight, conv1_ext.fc.weight, conv2.fc.weight, conv2_ext.fc.weight, graph, *dead)


# ------------------------------------------------------------------------------ Vanilla ("Naive") convolution
def node_linear2(a, b, a2, b2, trans_b=True, bias=None, mask=None, act=ACT_NONE, out=None):
    """``act(a @ op(b) + a2 @ op(b2) + bias)``: a Linear over the concatenation [a | a2] without building it."""
    lib = _lib.load()
    a, b, a2, b2 = _f32_cuda(a, "a"), _f32_cuda(b, "b"), _f32_cuda(a2, "a2"), _f32_cuda(b2, "b2")
    n, k = a.shape
    k2 = a2.shape[1]
    m = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=a.device)
    if mask is not None:
        mask = _f32_cuda(mask, "mask")
    if bias is not None:
        bias = bias.contiguous()
    with torch.cuda.device(a.device):
        rc = lib.drk_node_linear2(_p(a), _ld(a), _p(b), _ld(b), k, _p(a2), _ld(a2), _p(b2), _ld(b2), k2, 1 if trans_b else 0, _p(bias),
                                  _p(mask), _ld(mask) if mask is not None else 0, _p(out), _ld(out), n, m, act, stream_ptr())
    _lib.check(rc, "drk_node_linear2")
    return out


MESSAGE_SIZE = 32  # vanilla_gnn.py:20


def edge_msg_fwd(graph: GraphIndex, uv, edge_attr, cmat):
    """S[i] = sum_e relu(U[i] + V[col_e] + C attr_e); returns (S [N,32], cnt [N,32], mask uint32 [E] in CSR-SLOT order).
    The edge attributes are read in slot order too (``graph.attr_in_slot_order``: one gather per batch, shared by both layers and
    their backward passes), so the kernel streams them instead of chasing ``perm -> attr`` per edge."""
    lib = _lib.load()
    uv = _f32_cuda(uv, "uv")
    n = uv.shape[0]
    fe = 0 if edge_attr is None else edge_attr.shape[1]
    if fe:
        edge_attr = graph.attr_in_slot_order(_f32_cuda(edge_attr, "edge_attr"))
    s = torch.empty((n, MESSAGE_SIZE), dtype=torch.float32, device=uv.device)
    cnt = torch.empty_like(s)
    mask = torch.empty(max(graph.num_edges, 1), dtype=torch.int32, device=uv.device)
    with torch.cuda.device(uv.device):
        rc = lib.drk_edge_msg_fwd(_p(graph.rowptr), _p(graph.colidx), None, _p(uv), _ld(uv), _p(edge_attr) if fe else None,
                                  _ld(edge_attr) if fe else 0, fe, _p(cmat) if fe else None, int(cmat.stride(0)) if fe else 0, _p(s), _ld(s),
                                  _p(cnt), _p(mask), n, stream_ptr())
    _lib.check(rc, "drk_edge_msg_fwd")
    return s, cnt, mask


def edge_msg_bwd_src(graph: GraphIndex, ds, mask, out):
    lib = _lib.load()
    ds = _f32_cuda(ds, "ds")
    with torch.cuda.device(ds.device):
        # masks are kept per CSR slot: the source-side walk reaches them through the CSC-slot -> CSR-slot map
        rc = lib.drk_edge_msg_bwd_src(_p(graph.colptr), _p(graph.rowidx), _p(graph.slot_map()), _p(ds), _ld(ds), _p(mask), _p(out), _ld(out),
                                      ds.shape[0], stream_ptr())
    _lib.check(rc, "drk_edge_msg_bwd_src")
    return out


def edge_msg_bwd_c(graph: GraphIndex, ds, mask, edge_attr, out):
    lib = _lib.load()
    ds = _f32_cuda(ds, "ds")
    edge_attr = graph.attr_in_slot_order(_f32_cuda(edge_attr, "edge_attr"))
    with torch.cuda.device(ds.device):
        ws = workspace(lib.drk_edge_msg_bwd_c_workspace_bytes(), ds.device)
        rc = lib.drk_edge_msg_bwd_c(_p(graph.rowptr), None, _p(ds), _ld(ds), _p(mask), _p(edge_attr), _ld(edge_attr),
                                    edge_attr.shape[1], _p(out), int(out.stride(0)), ds.shape[0], _p(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_edge_msg_bwd_c")
    return out


# ------------------------------------------------------------------------------ per-graph fused Vanilla layer (drk_vanilla.cu)
VANILLA_FUSED = os.environ.get("DRK_VANILLA_FUSED", "1") != "0"
# One CTA per graph: a batch of few graphs leaves most of the 148 SMs idle and the batch-level kernels win.  Measured on the C2 graphs
# (profiles/vanilla_crossover_probe.py, train step in ms, fused / batch-level): 16 graphs 0.42 / 0.30, 64: 0.45 / 0.42, 96: 0.46 / 0.51,
# 128: 0.48 / 0.59, 256: 0.78 / 0.89.
VANILLA_FUSED_MIN_GRAPHS = int(os.environ.get("DRK_VANILLA_FUSED_MIN_GRAPHS", "80"))


def _vanilla_fused_ok(x, we, wn, graph: GraphIndex, f: int, fe: int) -> bool:
    """One CTA per graph needs the graphs' node ranges (collated batches) and every graph to fit one SM's shared memory."""
    if not VANILLA_FUSED or graph.graph_ptr is None or not graph.max_graph_nodes or graph.colptr is None or graph.num_graphs < max(1, VANILLA_FUSED_MIN_GRAPHS):
        return False
    if not (x.is_cuda and x.dtype == torch.float32 and we.dtype == torch.float32 and wn.dtype == torch.float32):
        return False
    lib = _lib.load()
    if lib.drk_vanilla_layer_bwd_workspace_bytes(f, graph.num_graphs) > (1 << 28):  # one gradient partial per graph: batches of many thousand tiny graphs
        return False
    return bool(lib.drk_vanilla_layer_supported(f, fe, int(graph.max_graph_nodes)))


def _rows16(t: torch.Tensor) -> torch.Tensor:
    """Row-contiguous and 16-byte aligned (what the bulk copies of the per-graph kernels need)."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def vanilla_layer_fwd(x, edge_attr, we, be, wn, bn, graph: GraphIndex, for_backward: bool = True):
    """``VanillaConvolutionalLayer.forward`` in one launch; returns (out, S, cnt, tf, mask) -- see include/drk_b200.h.
    ``for_backward=False`` (inference): the per-node counts of active edges / attribute sums are not produced."""
    lib = _lib.load()
    x = _rows16(x)
    n, f = x.shape
    fe = 0 if edge_attr is None else edge_attr.shape[1]
    attr = graph.attr_in_slot_order(_f32_cuda(edge_attr, "edge_attr")).contiguous() if fe else None
    we = we if we.stride(1) == 1 else we.contiguous()
    wn = wn if wn.stride(1) == 1 else wn.contiguous()
    dev = x.device
    out = torch.empty((n, f), dtype=torch.float32, device=dev)
    s = torch.empty((n, MESSAGE_SIZE), dtype=torch.float32, device=dev)
    cnt = torch.empty_like(s) if for_backward else None
    tf = torch.empty((n, fe, MESSAGE_SIZE), dtype=torch.float32, device=dev) if fe and for_backward else None
    mask = torch.empty(max(graph.num_edges, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.drk_vanilla_layer_fwd(_p(x), f, _p(graph.rowptr), _p(graph.colidx), _p(attr), fe, _p(graph.graph_ptr), _p(graph.order), graph.num_graphs,
                                       int(graph.max_graph_nodes), _p(we), int(we.stride(0)), _p(be), _p(wn), int(wn.stride(0)), _p(bn), _p(out), _p(s),
                                       _p(cnt), _p(tf), _p(mask), _p(graph.status), stream_ptr())
    _lib.check(rc, "drk_vanilla_layer_fwd")
    return out, s, cnt, tf, mask


def vanilla_layer_bwd(x, s, out, dout, cnt, tf, mask, we, wn, graph: GraphIndex, has_be: bool, has_bn: bool, need_dx: bool):
    lib = _lib.load()
    x, dout = _rows16(x), _rows16(dout)
    n, f = x.shape
    fe = we.shape[1] - 2 * f
    dev = x.device
    we = we if we.stride(1) == 1 else we.contiguous()
    wn = wn if wn.stride(1) == 1 else wn.contiguous()
    dx = torch.empty_like(x) if need_dx else None
    dwe, dwn = torch.empty(we.shape, dtype=torch.float32, device=dev), torch.empty(wn.shape, dtype=torch.float32, device=dev)
    dbe = torch.empty(MESSAGE_SIZE, dtype=torch.float32, device=dev) if has_be else None
    dbn = torch.empty(f, dtype=torch.float32, device=dev) if has_bn else None
    with torch.cuda.device(dev):
        ws = workspace(lib.drk_vanilla_layer_bwd_workspace_bytes(f, graph.num_graphs), dev)
        rc = lib.drk_vanilla_layer_bwd(_p(x), _p(s), _p(out), _p(dout), _p(cnt), _p(tf), f, fe, _p(mask), _p(graph.colptr), _p(graph.rowidx),
                                       _p(graph.slot_map()), _p(graph.graph_ptr), _p(graph.order), graph.num_graphs, int(graph.max_graph_nodes),
                                       _p(we), int(we.stride(0)), _p(wn), int(wn.stride(0)), _p(dx), _p(dwe), int(dwe.stride(0)), _p(dbe),
                                       _p(dwn), int(dwn.stride(0)), _p(dbn), _p(graph.status), _p(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_vanilla_layer_bwd")
    return dx, dwe, dbe, dwn, dbn


class VanillaConvFunction(torch.autograd.Function):
    """``VanillaConvolutionalLayer.forward`` (reference ``vanilla_gnn.py:26-38``) as one autograd node.

    The edge MLP ``Linear(2F+Fe, 32)`` on ``cat[x_i, x_j, e]`` is split by columns into a destination part
    Wa, a source part Wb and an edge-feature part C, so that
        message_e = relu( (x Wa^T + b)[i] + (x Wb^T)[j] + C e ),
    i.e. two NODE-level projections (one 64-wide GEMM over N rows instead of a 32-wide one over E rows of
    width 2F+Fe) and a 32-wide gather per edge.  The node MLP on ``cat[x, sums]`` is a two-source Linear.
    """

    @staticmethod
    def forward(ctx, x, edge_attr, we, be, wn, bn, graph: GraphIndex):
        n, f = x.shape
        fe = we.shape[1] - 2 * f
        if we.shape[0] != MESSAGE_SIZE or fe < 0:
            raise ValueError(f"edge MLP weight has shape {tuple(we.shape)}, expected [32, 2*{f}+Fe]")
        if fe > 0 and (edge_attr is None or edge_attr.dim() != 2 or edge_attr.shape[1] != fe):
            raise ValueError(f"edge_attr must be [E, {fe}] (2-D, like the reference requires), got {None if edge_attr is None else tuple(edge_attr.shape)}")
        ctx.graph = graph
        ctx.f, ctx.fe = f, fe
        ctx.has_be, ctx.has_bn = be is not None, bn is not None
        ctx.fused = _vanilla_fused_ok(x, we, wn, graph, f, fe)
        if ctx.fused:
            out, s, cnt, tf, mask = vanilla_layer_fwd(x, edge_attr if fe else None, we, be, wn, bn, graph, for_backward=any(ctx.needs_input_grad))
            ctx.save_for_backward(x, None, we, wn, tf, s, cnt, mask, out)
            return out
        wab = torch.cat([we[:, :f], we[:, f : 2 * f]], dim=0)                       # [64, F]
        bias64 = torch.cat([be, torch.zeros_like(be)]) if be is not None else None  # b belongs to U only
        uv = node_linear(x, wab, True, bias64)                                        # [N, 64] = U | V
        cmat = we[:, 2 * f :]                                                         # [32, Fe] view, row stride 2F+Fe
        s, cnt, mask = edge_msg_fwd(graph, uv, edge_attr if fe else None, cmat)
        out = node_linear2(x, wn[:, :f], s, wn[:, f:], True, bn, act=ACT_RELU)        # relu(cat[x, s] Wn^T + bn)
        ctx.save_for_backward(x, edge_attr if fe else None, we, wn, wab, s, cnt, mask, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, edge_attr, we, wn, wab, s, cnt, mask, out = ctx.saved_tensors
        g = ctx.graph
        f, fe = ctx.f, ctx.fe
        n = x.shape[0]
        if ctx.fused:  # the fifth saved tensor is tf (sum of active edge attributes), not the stacked weight
            dx, dwe, dbe, dwn, dbn = vanilla_layer_bwd(x, s, out, dout, cnt, wab, mask, we, wn, g, ctx.has_be, ctx.has_bn, ctx.needs_input_grad[0])
            return dx, None, dwe, dbe, dwn, dbn, None
        dz = torch.ops.aten.threshold_backward(dout.contiguous(), out, 0.0)          # ReLU of the node MLP
        # node MLP: dWn = dZ^T [x | s], dbn = sum dZ, dS = dZ Wn[:, F:]
        dwn = torch.empty_like(wn)
        dbn = None
        if ctx.has_bn:
            _, dbn = weight_grad(dz, x, want_bias=True, dw=dwn[:, :f])
        else:
            weight_grad(dz, x, dw=dwn[:, :f])
        weight_grad(dz, s, dw=dwn[:, f:])
        ds = node_linear(dz, wn[:, f:], False)                                        # [N, 32]
        # edge messages: dU = dS * cnt (every active edge of row i adds dS[i]), dV over the CSC half
        duv = torch.empty((n, 2 * MESSAGE_SIZE), dtype=torch.float32, device=x.device)
        torch.mul(ds, cnt, out=duv[:, :MESSAGE_SIZE])
        edge_msg_bwd_src(g, ds, mask, duv[:, MESSAGE_SIZE:])
        dwe = torch.empty_like(we)
        dwab, dbe64 = weight_grad(duv, x, want_bias=True)                             # [64, F], [64]
        dwe[:, :f].copy_(dwab[:MESSAGE_SIZE])
        dwe[:, f : 2 * f].copy_(dwab[MESSAGE_SIZE:])
        if fe:
            edge_msg_bwd_c(g, ds, mask, edge_attr, dwe[:, 2 * f :])
        dbe = dbe64[:MESSAGE_SIZE] if ctx.has_be else None
        dx = node_linear2(dz, wn[:, :f], duv, wab, False) if ctx.needs_input_grad[0] else None
        return dx, None, dwe, dbe, dwn, dbn, None


def vanilla_conv(x, edge_attr, edge_mlp: torch.nn.Linear, node_mlp: torch.nn.Linear, graph):
    return VanillaConvFunction.apply(x, edge_attr, edge_mlp.weight, edge_mlp.bias, node_mlp.weight, node_mlp.bias, graph)


# ------------------------------------------------------------------------------ Fout convolution
class FoutConvFunction(torch.autograd.Function):
    """``FoutLayer.forward`` (reference ``foutnet.py:48-66``): ``x Wc + mean_{j in N(i)} (x Wn)[j] + b``.

    The reference loops over nodes in Python (O(N*E)); here one [N, 2Fo] projection ``x [Wc | Wn]`` and one
    segmented mean over the CSR.  A node without neighbours gets a NaN row exactly like ``torch.mean`` of
    the empty slice ``beta[index]`` does (``foutnet.py:58``)."""

    @staticmethod
    def forward(ctx, x, wc, wn, bias, graph: GraphIndex, relu: bool):
        n = x.shape[0]
        fo = wc.shape[1]
        wcat = torch.cat([wc, wn], dim=1)                                    # [Fi, 2Fo]
        bias2 = torch.cat([bias, torch.zeros_like(bias)]) if bias is not None else None
        ab = node_linear(x, wcat, False, bias2)                              # alpha + b | beta
        out = spmm(graph.rowptr, graph.colidx, ab[:, fo:], n, addend=ab[:, :fo], reduce=REDUCE_MEAN_NAN, act=ACT_RELU if relu else ACT_NONE, graph=graph)
        ctx.graph, ctx.relu, ctx.fo, ctx.has_bias = graph, relu, fo, bias is not None
        ctx.save_for_backward(x, wcat, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wcat, out = ctx.saved_tensors
        g, fo = ctx.graph, ctx.fo
        n = x.shape[0]
        dout = dout.contiguous()
        if ctx.relu:
            dout = torch.ops.aten.threshold_backward(dout, out, 0.0)
        dab = torch.empty((n, 2 * fo), dtype=torch.float32, device=x.device)
        dab[:, :fo].copy_(dout)
        # d beta[j] = sum_{e: col_e = j} dout[row_e] / deg(row_e): scale rows once, then A^T through the CSC half
        scaled = dout / g.degree().unsqueeze(1)
        spmm(g.colptr, g.rowidx, scaled, n, out=dab[:, fo:], graph=g)
        dwcat = weight_grad(x, dab)                                          # x^T dab = [Fi, 2Fo]
        db2 = dout.sum(0) if ctx.has_bias else None
        dx = node_linear(dab, wcat, True) if ctx.needs_input_grad[0] else None
        return dx, dwcat[:, :fo], dwcat[:, fo:], db2, None, None


def fout_conv(x, wc, wn, bias, graph, relu=False):
    return FoutConvFunction.apply(x, wc, wn, bias, graph, relu)


# ------------------------------------------------------------------------------ torch_scatter-compatible functions
# Function-level seam of the reference: scatter_sum / scatter_mean / scatter_max with an arbitrary
# int64 index along dim 0.  The index is counting-sorted on the device (drk_segment_index_build) and the reduction
# walks each segment in ascending element id -- the visiting order of the reference's CPU scatter_add_ -- without
# atomics.  As in torch_scatter, the output size is `dim_size` or index.max()+1 (the latter costs a host sync).
class SegmentPlan:
    """ptr/perm of one index vector; re-usable across scatter calls that share the index."""

    def __init__(self, index: torch.Tensor, dim_size: int | None = None):
        if index.dim() != 1:
            raise ValueError("only 1-D indices along dim 0 are supported (the reference uses nothing else)")
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
        self.index = index
        self.n_seg = int(dim_size)
        self.n_src = int(index.numel())
        self.ptr, self.perm, self.status = segment_index(index, self.n_seg)

    @classmethod
    def from_parts(cls, index: torch.Tensor, ptr: torch.Tensor, perm: torch.Tensor, n_seg: int, status: torch.Tensor | None = None) -> "SegmentPlan":
        """A plan whose grouping already exists (e.g. the compacted cluster index of ``utils.community_pooling``)."""
        plan = cls.__new__(cls)
        plan.index, plan.n_seg, plan.n_src = index, int(n_seg), int(index.numel())
        plan.ptr, plan.perm, plan.status = ptr, perm, status
        return plan

    def count(self) -> torch.Tensor:
        return (self.ptr[1:] - self.ptr[:-1]).to(torch.float32)


def _as_2d(src: torch.Tensor):
    if src.dim() == 1:
        return src.unsqueeze(1), True
    if src.dim() != 2:
        raise ValueError("scatter_*: src must be 1-D or 2-D")
    return src, False


class _ScatterReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, plan: SegmentPlan, reduce: int, init):
        out = spmm(plan.ptr, plan.perm, src, plan.n_seg, addend=init, reduce=reduce)
        ctx.plan, ctx.reduce = plan, reduce
        return out

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        g = dout
        if ctx.reduce == REDUCE_MEAN_CLAMP:
            g = dout / plan.count().clamp(min=1).unsqueeze(1)
        dsrc = g.index_select(0, plan.index) if ctx.needs_input_grad[0] else None
        dinit = dout if ctx.needs_input_grad[3] else None
        return dsrc, None, None, dinit


def scatter_sum(src, index, dim=0, out=None, dim_size=None, plan: SegmentPlan | None = None):
    """``torch_scatter.scatter_sum`` along dim 0 (``ginet.py:58``, ``vanilla_gnn.py:35``).  ``out`` is summed into."""
    if dim not in (0, -src.dim()):
        raise NotImplementedError("scatter_sum: only dim=0 is used on the DeepRank2 path")
    src2, squeeze = _as_2d(src)
    if plan is None:
        plan = SegmentPlan(index, dim_size if dim_size is not None else (out.shape[0] if out is not None else None))
    init = None if out is None else _as_2d(out)[0]
    res = _ScatterReduce.apply(src2, plan, REDUCE_SUM, init)
    res = res.squeeze(1) if squeeze else res
    if out is not None:
        out.data.copy_(res.detach())
    return res


def scatter_mean(src, index, dim=0, out=None, dim_size=None, plan: SegmentPlan | None = None):
    """``torch_scatter.scatter_mean``: sum / max(count, 1); a passed ``out`` joins the sum before the divide
    (``sgat.py:72``; readout ``ginet.py:117-118``; position pooling ``community_pooling.py:216``)."""
    if dim not in (0, -src.dim()):
        raise NotImplementedError("scatter_mean: only dim=0 is used on the DeepRank2 path")
    src2, squeeze = _as_2d(src)
    if plan is None:
        plan = SegmentPlan(index, dim_size if dim_size is not None else (out.shape[0] if out is not None else None))
    if out is None:
        res = _ScatterReduce.apply(src2, plan, REDUCE_MEAN_CLAMP, None)
    else:
        total = _ScatterReduce.apply(src2, plan, REDUCE_SUM, _as_2d(out)[0])
        res = total / plan.count().clamp(min=1).unsqueeze(1)
    res = res.squeeze(1) if squeeze else res
    if out is not None:
        out.data.copy_(res.detach())
    return res


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, plan: SegmentPlan):
        lib = _lib.load()
        src = _f32_cuda(src, "src")
        width = src.shape[1]
        out = torch.empty((plan.n_seg, width), dtype=torch.float32, device=src.device)
        arg = torch.empty((plan.n_seg, width), dtype=torch.int32, device=src.device)
        with torch.cuda.device(src.device):
            rc = lib.drk_segment_max(_p(plan.ptr), _p(plan.perm), _p(src), _ld(src), src.shape[0], plan.n_seg, width, _p(out), _ld(out), _p(arg), stream_ptr())
        _lib.check(rc, "drk_segment_max")
        ctx.plan, ctx.n_src = plan, src.shape[0]
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, dout, _darg):
        lib = _lib.load()
        (arg,) = ctx.saved_tensors
        dout = _f32_cuda(dout, "dout")
        width = dout.shape[1]
        dsrc = torch.zeros((ctx.n_src, width), dtype=torch.float32, device=dout.device)
        with torch.cuda.device(dout.device):
            rc = lib.drk_segment_max_bwd(_p(dout), _ld(dout), _p(arg), ctx.n_src, ctx.plan.n_seg, width, _p(dsrc), _ld(dsrc), stream_ptr())
        _lib.check(rc, "drk_segment_max_bwd")
        return dsrc, None


def scatter_max(src, index, dim=0, out=None, dim_size=None, plan: SegmentPlan | None = None):
    """``torch_scatter.scatter_max`` -> ``(out, argmax)``: first maximum wins, empty segment -> (0, len(src)),
    gradient flows to the argmax element only (``community_pooling.py:209``)."""
    if dim not in (0, -src.dim()) or out is not None:
        raise NotImplementedError("scatter_max: only dim=0 without out= is used on the DeepRank2 path")
    src2, squeeze = _as_2d(src)
    if plan is None:
        plan = SegmentPlan(index, dim_size)
    res, arg = _ScatterMax.apply(src2, plan)
    arg = arg.to(torch.int64)
    return (res.squeeze(1), arg.squeeze(1)) if squeeze else (res, arg)


def cluster_offsets(cluster: torch.Tensor, graph: GraphIndex) -> torch.Tensor:
    """In-place ``get_preloaded_cluster`` on the device; returns the device scalar holding the id count."""
    lib = _lib.load()
    if not cluster.is_cuda or cluster.dtype != torch.int64 or not cluster.is_contiguous():
        raise TypeError("cluster must be a contiguous int64 CUDA tensor")
    total = torch.empty(1, dtype=torch.int64, device=cluster.device)
    with torch.cuda.device(cluster.device):
        ws = workspace(lib.drk_cluster_offsets_workspace_bytes(graph.num_graphs), cluster.device)
        rc = lib.drk_cluster_offsets(_p(cluster), _p(graph.graph_ptr), _p(graph.batch32), cluster.numel(), graph.num_graphs, _p(total), _p(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_cluster_offsets")
    return total


# ------------------------------------------------------------------------------ SGAT convolution
class SGATConvFunction(torch.autograd.Function):
    """``SGraphAttentionLayer.forward`` (``sgat.py:56-84``, undirected): see the module docstring of
    ``neuralnets/gnn/sgat.py`` for the algebra.  ``a`` is the edge attribute in ORIGINAL edge order."""

    @staticmethod
    def forward(ctx, x, a, weight, bias, graph: GraphIndex):
        n, fi = x.shape
        fo = weight.shape[1]
        wcat = torch.cat([weight[:fi], weight[fi:]], dim=1)          # [Fi, 2Fo] = Wt | Wb
        pq = node_linear(x, wcat, False)                             # P | Q
        a_csr = a.index_select(0, graph.perm.long()).contiguous()    # edge attribute in CSR order
        asum = spmm(graph.rowptr, None, a_csr.unsqueeze(1), n).squeeze(1)   # sum_j a_ij per destination
        deg = graph.degree().clamp(min=1)
        wq = spmm(graph.rowptr, graph.colidx, pq[:, fo:], n, w=a_csr)       # sum_j a_ij Q_j
        out = (pq[:, :fo] * asum.unsqueeze(1) + wq) / deg.unsqueeze(1)
        if bias is not None:
            out = out + bias
        ctx.graph, ctx.fo, ctx.has_bias = graph, fo, bias is not None
        ctx.save_for_backward(x, a, wcat, asum, deg)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, a, wcat, asum, deg = ctx.saved_tensors
        g, fo = ctx.graph, ctx.fo
        n = x.shape[0]
        dmean = dout / deg.unsqueeze(1)                               # [N, Fo]
        dpq = torch.empty((n, 2 * fo), dtype=torch.float32, device=x.device)
        torch.mul(dmean, asum.unsqueeze(1), out=dpq[:, :fo])          # dP_i = dmean_i * sum_j a_ij
        a_csc = a.index_select(0, g.permT.long()).contiguous()
        spmm(g.colptr, g.rowidx, dmean, n, w=a_csc, out=dpq[:, fo:])  # dQ_j = sum_{e: col=j} a_e dmean[row_e]
        dwcat = weight_grad(x, dpq)                                   # [Fi, 2Fo]
        dweight = torch.cat([dwcat[:, :fo], dwcat[:, fo:]], dim=0)
        dbias = dout.sum(0) if ctx.has_bias else None
        dx = node_linear(dpq, wcat, True) if ctx.needs_input_grad[0] else None
        return dx, None, dweight, dbias, None


def sgat_conv(x, a, weight, bias, graph):
    return SGATConvFunction.apply(x, a, weight, bias, graph)


# ------------------------------------------------------------------------------ fused per-graph GINet
def ginet_fused_max_nodes(fi: int) -> int:
    return int(_lib.load().drk_ginet_fused_max_nodes(int(fi)))


class GINetFusedFunction(torch.autograd.Function):
    """The same computation as :class:`GINetStackFunction` in TWO kernel launches (+ a 2 KB partial reduction): one CTA
    per graph, all intermediates in shared memory (``csrc/drk_ginet_fused.cu``).  Used when every graph of the batch fits
    the shared-memory budget (``ginet_fused_max_nodes``) and F <= 64."""

    @staticmethod
    def forward(ctx, x, w1, w1e, w2, w2e, graph: GraphIndex, max_nodes: int, max_edges: int, *dead):
        lib = _lib.load()
        x = _f32_cuda(x, "x")
        n, fi = x.shape
        w1s = torch.cat([w1, w1e], dim=0).contiguous()
        w2c, w2ec = w2.contiguous(), w2e.contiguous()
        need_grad = any(ctx.needs_input_grad[1:5])
        h1 = torch.empty((n, 32), dtype=torch.float32, device=x.device) if need_grad else None
        a2 = torch.empty((n, 32), dtype=torch.float32, device=x.device) if need_grad else None
        g = torch.empty((graph.num_graphs, 64), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.drk_ginet_fused_fwd(_p(x), _ld(x), fi, _p(graph.graph_ptr), _p(graph.rowptr), _p(graph.colidx), _p(w1s), _p(w2c), _p(w2ec),
                                         _p(h1), _p(a2), _p(g), graph.num_graphs, max_nodes, max_edges, _p(graph.status), stream_ptr())
        _lib.check(rc, "drk_ginet_fused_fwd")
        ctx.graph, ctx.max_nodes, ctx.max_edges, ctx.dead = graph, max_nodes, max_edges, dead
        ctx.save_for_backward(x, w2c, w2ec, h1, a2)
        return g

    @staticmethod
    def backward(ctx, dg):
        lib = _lib.load()
        x, w2c, w2ec, h1, a2 = ctx.saved_tensors
        graph = ctx.graph
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("the fused GINet path does not produce d/dx (node features are inputs); use the unfused path")
        fi = x.shape[1]
        dg = dg.contiguous()
        dw1s = torch.empty((32, fi), dtype=torch.float32, device=x.device)
        dw2 = torch.empty((32, 16), dtype=torch.float32, device=x.device)
        dw2e = torch.empty((32, 16), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = workspace(lib.drk_ginet_fused_bwd_workspace_bytes(), x.device)
            rc = lib.drk_ginet_fused_bwd(_p(x), _ld(x), fi, _p(graph.graph_ptr), _p(graph.colptr), _p(graph.rowidx), _p(w2c), _p(w2ec), _p(h1), _p(a2),
                                         _p(dg), _p(dw1s), _p(dw2), _p(dw2e), graph.num_graphs, ctx.max_nodes, ctx.max_edges, _p(graph.status), _p(ws), ws.numel(), stream_ptr())
        _lib.check(rc, "drk_ginet_fused_bwd")
        dead = tuple(torch.zeros_like(t) if ctx.needs_input_grad[8 + i] else None for i, t in enumerate(ctx.dead))
        return (None, dw1s[:16], dw1s[16:], dw2, dw2e, None, None, None) + dead


def ginet_fused(x, conv1, conv1_ext, conv2, conv2_ext, graph, max_nodes, max_edges=0):
    dead = []
    for layer in (conv1, conv2, conv1_ext, conv2_ext):
        dead += [layer.fc_edge_attr.weight, layer.fc_attention.weight]
    return GINetFusedFunction.apply(x, conv1.fc.weight, conv1_ext.fc.weight, conv2.fc.weight, conv2_ext.fc.weight, graph, max_nodes, max_edges, *dead)
