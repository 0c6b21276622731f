"""Community pooling on the device (mirror of ``deeprank2/utils/community_pooling.py``).

Only the *pooling* half of the reference file is on the path; community *detection* (networkx + MCL /
Louvain, ``community_pooling.py:30-162``) is CPU preprocessing that runs once in ``Trainer._precluster`` and
is out of scope.

Integer outputs (relabelled clusters, pooled ``edge_index``, pooled ``batch``) are bit-exact against the
reference semantics (PyG 2.4 ``consecutive_cluster`` / ``pool_edge`` / ``pool_batch``); ``x`` is a max (exact),
pooled ``edge_attr`` / ``pos`` are sums in ascending element order (fp32 tolerance).

The number of clusters is data dependent and sizes the outputs, so each pooling step reads ONE scalar back
to the host (``torch.unique``); the reference does ``B`` synchronisations in ``get_preloaded_cluster`` alone.
"""
from __future__ import annotations

import torch

from .. import ops
from ..data import Batch, Data
from ..graph import GraphIndex


def _node_graph_index(batch: torch.Tensor, num_graphs: int | None) -> GraphIndex:
    """graph offsets of a (pooled) batch vector, without edges."""
    empty = torch.empty((2, 0), dtype=torch.int64, device=batch.device)
    return GraphIndex.build(empty, batch.numel(), batch=batch, num_graphs=num_graphs, with_csc=False)


def get_preloaded_cluster(cluster, batch, num_graphs: int | None = None):
    """``cluster[batch == g] += max(cluster[batch == g-1]) + 1`` for g = 1..B-1, i.e. make the per-graph cluster ids
    of a collated batch globally unique (``community_pooling.py:23-27``).  In place, like the reference; one kernel
    chain instead of B host round trips."""
    if cluster.numel() == 0:
        return cluster
    gi = _node_graph_index(batch, num_graphs)
    ops.cluster_offsets(cluster, gi)
    return cluster


def consecutive_cluster(src: torch.Tensor):
    """PyG ``consecutive_cluster``: ``(inverse, perm)`` with ``perm[c]`` = the LARGEST node index of cluster c (what
    the reference's CPU ``scatter_`` leaves there: last writer wins)."""
    uniq, inv = torch.unique(src, sorted=True, return_inverse=True)
    n_clusters = int(uniq.numel())
    plan = ops.SegmentPlan(inv, n_clusters)
    last = plan.perm[(plan.ptr[1:] - 1).long()].to(torch.int64)
    return inv, last, plan


def pool_batch(perm, batch):
    return batch[perm]


def pool_edge(cluster, edge_index, edge_attr=None):
    """PyG ``pool_edge(reduce='sum')``: relabel by cluster, drop self loops, sort by (row, col), merge duplicates and
    sum their attributes.  The result is row-major sorted."""
    num_nodes = cluster.size(0)
    ei = cluster[edge_index.view(-1)].view(2, -1)
    keep = ei[0] != ei[1]
    ei = ei[:, keep]
    if edge_attr is not None:
        edge_attr = edge_attr[keep]
    if ei.numel() == 0:
        return ei, edge_attr
    key = ei[0] * num_nodes + ei[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    pooled_index = torch.stack([torch.div(uniq, num_nodes, rounding_mode="floor"), uniq % num_nodes])
    if edge_attr is None:
        return pooled_index, None
    squeeze = edge_attr.dim() == 1
    merged = ops.scatter_sum(edge_attr if not squeeze else edge_attr.unsqueeze(1), inv, dim=0, dim_size=int(uniq.numel()))
    return pooled_index, merged.squeeze(1) if squeeze else merged


def max_pool_x(cluster, x, batch):
    """PyG ``max_pool_x(cluster, x, batch)`` -> ``(x_pooled, batch_pooled)`` (``ginet.py:103,114``, ``foutnet.py:111``)."""
    inv, last, plan = consecutive_cluster(cluster)
    pooled, _ = ops.scatter_max(x, inv, dim=0, plan=plan)
    return pooled, pool_batch(last, batch)


def community_pooling(cluster, data):
    """Pool all members of a cluster into one node (``community_pooling.py:165-242``): feature-wise max of ``x``,
    pooled + coalesced edges with summed attributes, mean position, pooled batch vector; ``cluster0/1`` carried."""
    inv, last, plan = consecutive_cluster(cluster)
    x, _ = ops.scatter_max(data.x, inv, dim=0, plan=plan)
    edge_index, edge_attr = pool_edge(inv, data.edge_index, data.edge_attr)
    pos = ops.scatter_mean(data.pos, inv, dim=0, plan=plan) if getattr(data, "pos", None) is not None else None
    c0, c1 = getattr(data, "cluster0", None), getattr(data, "cluster1", None)
    if getattr(data, "batch", None) is not None:
        out = Batch(batch=pool_batch(last, data.batch), x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
        ng = data.__dict__.get("_num_graphs")
        if ng is None and data.__dict__.get("ptr") is not None:
            ng = int(data.ptr.numel()) - 1
        if ng is not None:
            out.__dict__["_num_graphs"] = ng
    else:
        out = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
    out.cluster0, out.cluster1 = c0, c1
    return out
