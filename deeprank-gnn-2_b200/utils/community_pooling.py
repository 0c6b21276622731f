"""Community pooling on the device (mirror of ``deeprank2/utils/community_pooling.py``).

Only the *pooling* half of the reference file is on the path; community *detection* (networkx + MCL /
Louvain, ``community_pooling.py:30-162``) is CPU preprocessing that runs once in ``Trainer._precluster`` and
is out of scope.

Integer outputs (relabelled clusters, pooled ``edge_index``, pooled ``batch``) are bit-exact against the
reference semantics (PyG 2.4 ``consecutive_cluster`` / ``pool_edge`` / ``pool_batch``); ``x`` is a max (exact),
pooled ``edge_attr`` / ``pos`` are sums in ascending element order (fp32 tolerance).

Everything runs in our own kernels (``csrc/drk_pool.cu`` + the counting sort of ``csrc/drk_index.cu``): no ``torch.unique``, no
sort, and NO host read-back on the step path.  The data-dependent sizes -- distinct clusters, distinct pooled edges -- are
properties of the graphs and their stored clusterings, so the host collate computes them once per graph
(``data.pool_sizes``) and a batch carries their totals (``Batch.meta("pool")``): outputs are allocated from those, the kernels
write the counts they find to device scalars (checked at the end of a pass through the status word), and a train step of
``FoutNet`` / clustered ``GINet`` / ``SGAT`` can be captured into a CUDA graph.  Tensors without that meta (hand-made
batches) take ONE host round trip to size the outputs -- the reference does ``B`` of them in ``get_preloaded_cluster`` alone.
"""
from __future__ import annotations

import os

import torch

from .. import _lib, ops
from ..data import Batch, Data
from ..graph import GraphIndex, stream_ptr, workspace


def _p(t):
    return None if t is None else t.data_ptr()


def _node_graph_index(batch: torch.Tensor, num_graphs: int | None) -> GraphIndex:
    """graph offsets of a (pooled) batch vector, without edges."""
    empty = torch.empty((2, 0), dtype=torch.int64, device=batch.device)
    return GraphIndex.build(empty, batch.numel(), batch=batch, num_graphs=num_graphs, with_csc=False)


def pool_meta(data, level: int):
    """``{"K": id bound, "C": distinct clusters, ...}`` of pooling level 0 (``cluster0`` on the input batch) or 1 (``cluster1`` on the
    pooled batch) if the batch came out of ``Batch.from_data_list`` / ``ResidentGraphSet.collate``; else None."""
    m = data.__dict__.get(getattr(data, "_META_KEY", "_meta"), {}).get("pool") if hasattr(data, "__dict__") else None
    if m is None:
        return None
    node_ptr = getattr(data, "_node_ptr32", None)
    whole = data.__dict__.get(getattr(data, "_META_KEY", "_meta"), {})
    if level == 0:
        cptr, kkptr = getattr(data, "_pool_cptr", None), getattr(data, "_pool_kkptr", None)
        if cptr is None or kkptr is None or not cptr.is_cuda:
            return None
        out = {"K": m["K0"], "C": m["C0"], "E": m["E1"], "KK": m["KK"], "cptr": cptr, "kkptr": kkptr}
        if node_ptr is not None and node_ptr.is_cuda and whole.get("max_graph_nodes") and m.get("max_K0"):
            out["blocks"] = (node_ptr, cptr, int(whole["max_graph_nodes"]), int(m["max_K0"]))  # per-graph consecutive_cluster
        return out
    out = {"K": m["K1"], "C": m["C1"]}
    c1ptr = getattr(data, "_pool_c1ptr", None)
    if node_ptr is not None and node_ptr.is_cuda and c1ptr is not None and c1ptr.is_cuda and whole.get("max_graph_nodes") and m.get("max_K1"):
        out["blocks"] = (node_ptr, c1ptr, int(whole["max_graph_nodes"]), int(m["max_K1"]))
    return out


def get_preloaded_cluster(cluster, batch, num_graphs: int | None = None):
    """``cluster[batch == g] += max(cluster[batch == g-1]) + 1`` for g = 1..B-1, i.e. make the per-graph cluster ids
    of a collated batch globally unique (``community_pooling.py:23-27``).  In place, like the reference; one kernel
    chain instead of B host round trips."""
    if cluster.numel() == 0:
        return cluster
    gi = _node_graph_index(batch, num_graphs)
    ops.cluster_offsets(cluster, gi)
    return cluster


_STATUS: dict = {}
POOL_BLOCKED = os.environ.get("DRK_POOL_BLOCKED", "1") != "0"  # per-graph pool_edge kernel for collated batches


def _pool_blocks(data):
    """(edge_ptr, pooled_edge_ptr, max clusters, max edges per graph) if ``data`` is a collated batch that still has its own edge list."""
    d = data.__dict__
    m = d.get(getattr(data, "_META_KEY", "_meta"), {})
    pool = m.get("pool")
    ei = d.get("edge_index")
    if pool is None or ei is None or m.get("num_edges_total") != int(ei.shape[1]) or m.get("max_graph_edges") is None or not pool.get("max_C0"):
        return None
    edge_ptr, pooled_ptr = getattr(data, "_edge_ptr32", None), getattr(data, "_pool_eptr32", None)
    if edge_ptr is None or pooled_ptr is None or not edge_ptr.is_cuda or not pooled_ptr.is_cuda:
        return None
    return edge_ptr, pooled_ptr, int(pool["max_C0"]), int(m["max_graph_edges"])


def _status_word(dev) -> torch.Tensor:
    """ONE status word per device: every pooling kernel ORs its data-dependent faults into it (no fill launch per call, no host sync);
    ``check_status`` reads and clears it once per pass."""
    dev = torch.device(dev)
    if dev.index is None:
        dev = torch.device(dev.type, torch.cuda.current_device())
    acc = _STATUS.get(dev)
    if acc is None:
        acc = _STATUS[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    return acc


def check_status(device=None) -> None:
    """Host sync: raise if any pooling kernel since the last call found an id outside the sizes it was given (a clustering that does
    not match the batch's collate meta) or an edge joining two graphs.  The Trainer calls it at the end of every pass."""
    for dev, acc in list(_STATUS.items()):
        want = torch.device(device) if device is not None else None
        if want is not None and (want.type != dev.type or (want.index is not None and want.index != dev.index)):
            continue
        flags = int(acc.item())
        acc.zero_()
        if flags & _lib.STATUS_INDEX_RANGE:
            raise IndexError("community pooling: a cluster id / pooled edge count exceeds the sizes recorded for the batch (stale `pool` meta?)")
        if flags & _lib.STATUS_CROSS_GRAPH:
            raise ValueError("community pooling: an edge joins two graphs of the batch, or a cluster spans two graphs")


class _Structure:
    """What ``consecutive_cluster`` yields on the device: ``inv`` int64 [N] (new id of every node), ``last`` int64 [C] (PyG's ``perm``:
    the largest node index of every cluster), a segment plan (nodes grouped by new id, ascending) and the status / count words."""

    __slots__ = ("inv", "last", "plan", "count", "status", "n_clusters")


def _host_sizes(cluster: torch.Tensor) -> tuple[int, int]:
    """(id bound, distinct ids) by ONE host round trip -- only for tensors that did not come with collate meta."""
    if cluster.numel() == 0:
        return 0, 0
    bound = int(cluster.max()) + 1
    present = torch.zeros(bound, dtype=torch.bool, device=cluster.device)
    present[cluster] = True
    return bound, int(present.sum())


def consecutive_cluster(src: torch.Tensor, meta: dict | None = None):
    """PyG ``consecutive_cluster``: ``(inverse, perm, plan)`` with ``perm[c]`` = the LARGEST node index of cluster c (what the
    reference's CPU ``scatter_`` leaves there: last writer wins).  ``meta = {"K": id bound, "C": distinct ids}`` (from the collate)
    avoids the host round trip."""
    st = _consecutive(src, meta)
    return st.inv, st.last, st.plan


def _consecutive(src: torch.Tensor, meta: dict | None) -> _Structure:
    lib = _lib.load()
    if not src.is_cuda or src.dtype != torch.int64 or src.dim() != 1:
        raise TypeError("cluster must be a 1-D int64 CUDA tensor")
    src = src.contiguous()
    bound, n_clusters = (int(meta["K"]), int(meta["C"])) if meta is not None else _host_sizes(src)
    dev, n = src.device, int(src.numel())
    blocks = meta.get("blocks") if meta is not None else None
    if (POOL_BLOCKED and blocks is not None and n > 0 and int(blocks[0].numel()) >= 2 and int(blocks[0].numel()) == int(blocks[1].numel())
            and lib.drk_consecutive_blocked_supported(int(blocks[2]), int(blocks[3]))):
        # collated batch: one CTA per graph relabels the graph's ids in shared memory (one launch instead of eight)
        node_ptr, cptr, max_nodes, max_ids = blocks
        status = _status_word(dev)
        inv = torch.empty(n, dtype=torch.int64, device=dev)
        last = torch.empty(n_clusters, dtype=torch.int64, device=dev)
        ptr_c = torch.empty(n_clusters + 1, dtype=torch.int32, device=dev)
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.drk_consecutive_blocked(_p(src), n, _p(node_ptr), _p(cptr), int(node_ptr.numel()) - 1, max_nodes, max_ids, n_clusters, _p(inv), _p(last),
                                             _p(ptr_c), _p(perm), _p(status), stream_ptr())
        _lib.check(rc, "drk_consecutive_blocked")
        st = _Structure()
        st.inv, st.last, st.count, st.status, st.n_clusters = inv, last, None, status, n_clusters
        st.plan = ops.SegmentPlan.from_parts(inv, ptr_c, perm, n_clusters, status)
        return st
    ptr_k, perm, status = ops.segment_index(src, bound, _status_word(src.device))  # nodes grouped by id, stable: ascending node index inside a cluster
    rank = torch.empty(max(bound, 1), dtype=torch.int64, device=dev)
    ptr_c = torch.empty(n_clusters + 1, dtype=torch.int32, device=dev)
    last = torch.empty(n_clusters, dtype=torch.int64, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(lib.drk_compact_segments_workspace_bytes(bound), dev)
        rc = lib.drk_compact_segments(_p(ptr_k), bound, _p(perm), _p(rank), _p(ptr_c), None, _p(last), n_clusters, _p(count), _p(status), _p(ws), ws.numel(),
                                      stream_ptr())
    _lib.check(rc, "drk_compact_segments")
    st = _Structure()
    st.inv = rank[src] if n else torch.empty(0, dtype=torch.int64, device=dev)
    st.last, st.count, st.status, st.n_clusters = last, count, status, n_clusters
    st.plan = ops.SegmentPlan.from_parts(st.inv, ptr_c, perm, n_clusters, status)
    return st


def pool_batch(perm, batch):
    return batch[perm]


def _host_pool_meta(inv: torch.Tensor, n_clusters: int, edge_index: torch.Tensor, batch: torch.Tensor | None):
    """cptr / kkptr / pooled-edge count for tensors without collate meta: one host round trip (the tensors are read on the CPU)."""
    import numpy as np

    dev = inv.device
    inv_h, ei_h = inv.cpu().numpy(), edge_index.cpu().numpy()
    b_h = batch.cpu().numpy() if batch is not None else np.zeros(inv_h.size, dtype=np.int64)
    n_graphs = int(b_h.max()) + 1 if b_h.size else 1
    graph_of_cluster = np.zeros(n_clusters, dtype=np.int64)
    graph_of_cluster[inv_h] = b_h
    counts = np.bincount(graph_of_cluster, minlength=n_graphs).astype(np.int64)
    cptr = np.zeros(n_graphs + 1, dtype=np.int64)
    np.cumsum(counts, out=cptr[1:])
    kkptr = np.zeros(n_graphs + 1, dtype=np.int64)
    np.cumsum(counts * counts, out=kkptr[1:])
    pr, pc = inv_h[ei_h[0]], inv_h[ei_h[1]]
    keep = pr != pc
    n_pooled = int(np.unique(pr[keep] * max(n_clusters, 1) + pc[keep]).size)
    return {"E": n_pooled, "KK": int(kkptr[-1]), "cptr": torch.from_numpy(cptr).to(dev), "kkptr": torch.from_numpy(kkptr).to(dev)}


def _pool_edge_blocked(cluster, edge_index, edge_attr, meta: dict, blocks: tuple):
    """``pool_edge`` of a collated batch in ONE launch (``drk_pool_edge_blocked``: one CTA per graph, in shared memory)."""
    lib = _lib.load()
    dev = cluster.device
    edge_ptr32, pooled_eptr32, max_clusters, max_edges = blocks
    n, e, n_pooled = int(cluster.numel()), int(edge_index.shape[1]), int(meta["E"])
    cptr = meta["cptr"]
    n_graphs = int(cptr.numel()) - 1
    edge_index = edge_index.contiguous()
    pooled_index = torch.empty((2, n_pooled), dtype=torch.int64, device=dev)
    src = merged = None
    if edge_attr is not None:
        if edge_attr.requires_grad:
            raise NotImplementedError("pool_edge: gradients with respect to edge_attr are not on the DeepRank2 path")
        src = (edge_attr.unsqueeze(1) if edge_attr.dim() == 1 else edge_attr).contiguous().to(torch.float32)
        merged = torch.empty((n_pooled, src.shape[1]), dtype=torch.float32, device=dev)
    status = _status_word(dev)
    with torch.cuda.device(dev):
        rc = lib.drk_pool_edge_blocked(_p(edge_index), e, _p(edge_ptr32), _p(cluster), n, _p(cptr), _p(pooled_eptr32), n_graphs, int(max_clusters), int(max_edges),
                                       _p(src), int(src.shape[1]) if src is not None else 0, int(src.shape[1]) if src is not None else 0, _p(pooled_index),
                                       n_pooled, _p(merged), _p(status), stream_ptr())
    _lib.check(rc, "drk_pool_edge_blocked")
    if merged is not None and edge_attr.dim() == 1:
        merged = merged.squeeze(1)
    return pooled_index, merged


def pool_edge(cluster, edge_index, edge_attr=None, meta: dict | None = None, batch: torch.Tensor | None = None, batch32: torch.Tensor | None = None,
              blocks: tuple | None = None):
    """PyG ``pool_edge(reduce='sum')``: relabel by cluster, drop self loops, sort by (row, col), merge duplicates and
    sum their attributes.  The result is row-major sorted.  ``cluster`` is the CONSECUTIVE relabelling (``inv``); ``meta`` carries the
    collate's sizes (``E`` pooled edges, ``KK`` dense pair ids, ``cptr`` / ``kkptr`` per-graph offsets); ``blocks`` =
    (edge_ptr int32 [G+1], pooled_edge_ptr int32 [G+1], max clusters per graph, max edges per graph) of a collated batch selects the
    per-graph kernel."""
    lib = _lib.load()
    dev = cluster.device
    n, e = int(cluster.numel()), int(edge_index.shape[1])
    if e == 0:
        return edge_index, edge_attr
    if (POOL_BLOCKED and blocks is not None and meta is not None and "cptr" in meta and int(meta["E"]) > 0
            and lib.drk_pool_edge_blocked_supported(int(blocks[2]), int(blocks[3]))):
        return _pool_edge_blocked(cluster, edge_index, edge_attr, meta, blocks)
    if meta is None or "cptr" not in meta:
        n_clusters = int(cluster.max()) + 1 if n else 0
        meta = _host_pool_meta(cluster, n_clusters, edge_index, batch)
    cptr, kkptr = meta["cptr"], meta["kkptr"]
    n_graphs = int(cptr.numel()) - 1
    n_pairs, n_pooled = int(meta["KK"]), int(meta["E"])
    if n_pairs + _lib.POOL_JUNK_SEGMENTS >= 2**31:
        raise NotImplementedError("pool_edge: the batch has more than 2^31 candidate pooled pairs")
    if batch32 is None:
        batch32 = batch.to(torch.int32) if batch is not None else torch.zeros(n, dtype=torch.int32, device=dev)
    edge_index = edge_index.contiguous()
    key = torch.empty(e, dtype=torch.int64, device=dev)
    status = _status_word(dev)
    with torch.cuda.device(dev):
        rc = lib.drk_pool_edge_keys(_p(edge_index), e, _p(cluster), n, _p(batch32), _p(cptr), _p(kkptr), n_graphs, n_pairs, _p(key), _p(status), stream_ptr())
    _lib.check(rc, "drk_pool_edge_keys")
    ptr_k, perm, _ = ops.segment_index(key, n_pairs + _lib.POOL_JUNK_SEGMENTS, status)  # edges grouped by pooled pair (the junk segments of the self loops come last)
    ptr_s = torch.empty(n_pooled + 1, dtype=torch.int32, device=dev)
    ids = torch.empty(max(n_pooled, 1), dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    pooled_index = torch.empty((2, n_pooled), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = workspace(lib.drk_compact_segments_workspace_bytes(n_pairs), dev)
        rc = lib.drk_compact_segments(_p(ptr_k), n_pairs, None, None, _p(ptr_s), _p(ids), None, n_pooled, _p(count), _p(status), _p(ws), ws.numel(), stream_ptr())
        _lib.check(rc, "drk_compact_segments")
        rc = lib.drk_pool_edge_decode(_p(ids), n_pooled, _p(count), _p(cptr), _p(kkptr), n_graphs, _p(pooled_index), stream_ptr())
        _lib.check(rc, "drk_pool_edge_decode")
    if edge_attr is None:
        return pooled_index, None
    if n_pooled == 0:
        return pooled_index, edge_attr[:0]
    if edge_attr.requires_grad:
        raise NotImplementedError("pool_edge: gradients with respect to edge_attr are not on the DeepRank2 path")
    squeeze = edge_attr.dim() == 1
    src = edge_attr.unsqueeze(1) if squeeze else edge_attr
    merged = ops.spmm(ptr_s, perm, src, n_pooled, reduce=ops.REDUCE_SUM)  # sums in ascending edge id, as scatter_add_ over the sorted list does
    return pooled_index, merged.squeeze(1) if squeeze else merged


def max_pool_x(cluster, x, batch, meta: dict | None = None, shared: dict | None = None):
    """PyG ``max_pool_x(cluster, x, batch)`` -> ``(x_pooled, batch_pooled)`` (``ginet.py:103,114``, ``foutnet.py:111``).
    ``shared``: a dict in which the weight-independent part (relabelling, pooled batch vector) is kept for a second call on the same
    clustering -- the two branches of the clustered GINet; ``cluster`` may then be None on the second call."""
    if shared is not None and "level1" in shared:
        st, pooled_batch = shared["level1"]
    else:
        st = _consecutive(cluster, meta)
        pooled_batch = pool_batch(st.last, batch)
        if shared is not None:
            shared["level1"] = (st, pooled_batch)
    pooled, _ = ops.scatter_max(x, st.inv, dim=0, plan=st.plan)
    return pooled, pooled_batch


def community_pooling(cluster, data, meta: dict | None = None, shared: dict | None = None):
    """Pool all members of a cluster into one node (``community_pooling.py:165-242``): feature-wise max of ``x``,
    pooled + coalesced edges with summed attributes, mean position, pooled batch vector; ``cluster0/1`` carried.
    ``shared``: a dict in which everything that does not depend on ``data.x`` (relabelling, pooled edges and attributes, positions,
    batch vector, and later the pooled batch's graph index) is kept for a second call on the same graph and clustering -- the two
    branches of the clustered GINet pool the same batch twice; ``cluster`` may then be None on the second call."""
    if shared is not None and "level0" in shared:
        import copy

        st, template = shared["level0"]
        out = copy.copy(template)
        out.__dict__ = dict(template.__dict__)  # the template's caches (graph index of the pooled batch) are shared, x is this call's
        out.x, _ = ops.scatter_max(data.x, st.inv, dim=0, plan=st.plan)
        return out
    if meta is None:
        meta = pool_meta(data, 0)
    st = _consecutive(cluster, meta)
    inv, last, plan = st.inv, st.last, st.plan
    x, _ = ops.scatter_max(data.x, inv, dim=0, plan=plan)
    batch = getattr(data, "batch", None)
    gi = data.__dict__.get("_graph_index")
    batch32 = gi.batch32 if gi is not None and gi.batch32 is not None else None
    edge_index, edge_attr = pool_edge(inv, data.edge_index, data.edge_attr, meta=meta, batch=batch, batch32=batch32, blocks=_pool_blocks(data))
    pos = ops.scatter_mean(data.pos, inv, dim=0, plan=plan) if getattr(data, "pos", None) is not None else None
    c0, c1 = getattr(data, "cluster0", None), getattr(data, "cluster1", None)
    if batch is not None:
        out = Batch(batch=pool_batch(last, batch), x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
        ng = data.__dict__.get("_num_graphs")
        if ng is None and data.__dict__.get("ptr") is not None:
            ng = int(data.ptr.numel()) - 1
        if ng is not None:
            out.__dict__["_num_graphs"] = ng
        m = data.__dict__.get(getattr(data, "_META_KEY", "_meta"), {}).get("pool")
        eptr = data.__dict__.get("_pool_eptr32")
        if m is not None and meta is not None and "cptr" in meta and eptr is not None and eptr.is_cuda and m["C0"] < 2**31:
            # the pooled batch is a collated batch again: per-graph node / edge slices for the blocked index build, and level-1 sizes
            out.__dict__["_node_ptr32"] = meta["cptr"].to(torch.int32)
            out.__dict__["_edge_ptr32"] = eptr
            if getattr(data, "_pool_c1ptr", None) is not None:
                out.__dict__["_pool_c1ptr"] = data._pool_c1ptr
            out.__dict__[out._META_KEY] = {"num_graphs": ng, "max_graph_nodes": m["max_C0"], "max_graph_edges": m["max_E1"], "num_edges_total": m["E1"],
                                           "pool": m}
    else:
        out = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
    out.cluster0, out.cluster1 = c0, c1
    out.__dict__["_pool_status"] = st.status
    if shared is not None:
        shared["level0"] = (st, out)
    return out
