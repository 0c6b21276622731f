"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Pure-torch CPU restatement of the third-party graph primitives the DeepRank2
GNN path executes.  The upstream wheels are NOT present
in this image and their sources are NOT under /root/reference, so their
published semantics are restated here from the pinned versions:

* ``torch_scatter`` 2.1.2   (pins: reference ``env/deeprank2.yml:15-23``)
* ``torch_geometric`` 2.4.0

Call sites in the reference that fix the argument conventions:
``deeprank2/neuralnets/gnn/ginet.py:58`` (scatter_sum with ``out=``),
``ginet_nocluster.py:103-104`` (scatter_mean), ``utils/community_pooling.py:206-219``
(consecutive_cluster / scatter_max / pool_edge / pool_batch), ``ginet.py:103``
(max_pool_x), ``ginet.py:34-38`` (inits.uniform), ``trainer.py:541-557`` (DataLoader
-> Batch.from_data_list collate).

Everything here works on ``dim=0`` reductions (the only axis used on the path)
but keeps the general signatures so the reference files import unchanged.
"""
from __future__ import annotations

import copy
import math

import torch


# --------------------------------------------------------------------------- torch_scatter
def _expand_index(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    """torch_scatter.utils.broadcast: make ``index`` the same shape as ``src``."""
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    """sum of ``src`` rows per ``index`` value; ``out`` (if given) is accumulated into in place."""
    index = _expand_index(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter_sum(src, index, dim, out, dim_size)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    """sum / max(count, 1); a passed ``out`` takes part in the sum before the divide."""
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _expand_index(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


class _ScatterMax(torch.autograd.Function):
    """torch_scatter CPU rule: first maximum wins, empty segment -> (0, E); grad goes to arg only."""

    @staticmethod
    def forward(ctx, src, index, dim_size):
        n_src = src.size(0)
        flat = src.reshape(n_src, -1)
        width = flat.size(1)
        best = torch.full((dim_size, width), -math.inf, dtype=src.dtype)
        arg = torch.full((dim_size, width), n_src, dtype=torch.long)
        # vectorised "first max wins": sort edges by (segment, -value, edge id) per column is
        # expensive; instead use amax then pick the smallest edge id attaining it.
        idx2 = index.view(-1, 1).expand(n_src, width)
        if n_src > 0:
            best = best.scatter_reduce(0, idx2, flat, reduce="amax", include_self=True)
            hit = flat == best.gather(0, idx2)
            cand = torch.where(hit, torch.arange(n_src).view(-1, 1).expand(n_src, width), n_src)
            arg = arg.scatter_reduce(0, idx2, cand, reduce="amin", include_self=True)
        empty = arg == n_src
        best = best.masked_fill(empty, 0)
        ctx.save_for_backward(arg)
        ctx.n_src = n_src
        ctx.src_shape = src.shape
        out_shape = (dim_size,) + tuple(src.shape[1:])
        ctx.mark_non_differentiable(arg)
        return best.reshape(out_shape), arg.reshape(out_shape)

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        width = arg.size(1) if arg.dim() > 1 else 1
        g = grad_out.reshape(arg.size(0), -1)
        grad_src = torch.zeros(ctx.n_src + 1, width, dtype=grad_out.dtype)
        grad_src.scatter_(0, arg.reshape(arg.size(0), -1), g)
        return grad_src[: ctx.n_src].reshape(ctx.src_shape), None, None


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    if dim not in (0, -src.dim()):
        raise NotImplementedError("oracle scatter_max restates dim=0 only (the reference's only use)")
    if out is not None:
        raise NotImplementedError("oracle scatter_max: out= is not used by the reference")
    if dim_size is None:
        dim_size = 0 if index.numel() == 0 else int(index.max()) + 1
    return _ScatterMax.apply(src, index, dim_size)


# --------------------------------------------------------------------------- torch_geometric
def uniform(size, value):
    """torch_geometric.nn.inits.uniform: U(-1/sqrt(size), 1/sqrt(size)); None is a no-op."""
    if isinstance(value, torch.Tensor):
        bound = 1.0 / math.sqrt(size)
        value.data.uniform_(-bound, bound)


def consecutive_cluster(src):
    """relabel cluster ids to 0..C-1; ``perm[c]`` = a representative node of cluster c (last writer)."""
    unique, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.arange(inv.size(0), dtype=inv.dtype, device=inv.device)
    perm = inv.new_empty(unique.size(0)).scatter_(0, inv, perm)
    return inv, perm


def pool_batch(perm, batch):
    return batch[perm]


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, keep]
    return edge_index, (None if edge_attr is None else edge_attr[keep])


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="sum"):
    """sort by (row, col), merge duplicates, reduce their attributes (sum on this path)."""
    nnz = edge_index.size(1)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if nnz > 0 else 0
    key = edge_index[0] * num_nodes + edge_index[1]
    key, order = torch.sort(key, stable=True)
    edge_index = edge_index[:, order]
    if edge_attr is not None:
        edge_attr = edge_attr[order]
    first = torch.ones(nnz, dtype=torch.bool)
    first[1:] = key[1:] > key[:-1]
    if bool(first.all()):
        return edge_index, edge_attr
    edge_index = edge_index[:, first]
    if edge_attr is None:
        return edge_index, None
    slot = torch.cumsum(first.to(torch.long), 0) - 1
    if reduce not in ("sum", "add"):
        raise NotImplementedError(reduce)
    merged = scatter_sum(edge_attr, slot, dim=0, dim_size=edge_index.size(1))
    return edge_index, merged


def pool_edge(cluster, edge_index, edge_attr=None, reduce="sum"):
    num_nodes = cluster.size(0)
    edge_index = cluster[edge_index.view(-1)].view(2, -1)
    edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
    if edge_index.numel() > 0:
        edge_index, edge_attr = coalesce(edge_index, edge_attr, num_nodes, reduce=reduce)
    return edge_index, edge_attr


def _segment_amax(cluster, x, size=None):
    """PyG 2.4 ``scatter(..., reduce='max')`` on CPU: zeros.scatter_reduce_('amax', include_self=False)."""
    dim_size = size if size is not None else (int(cluster.max()) + 1 if cluster.numel() > 0 else 0)
    idx = cluster.view(-1, *([1] * (x.dim() - 1))).expand_as(x)
    return x.new_zeros((dim_size,) + tuple(x.shape[1:])).scatter_reduce_(0, idx, x, reduce="amax", include_self=False)


def max_pool_x(cluster, x, batch, batch_size=None, size=None):
    if size is not None:
        raise NotImplementedError("size= is not used on the reference path")
    cluster, perm = consecutive_cluster(cluster)
    x = _segment_amax(cluster, x)
    return x, pool_batch(perm, batch)


class Data:
    """attribute bag standing in for torch_geometric.data.Data (clone/to/num_nodes only)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos).items():
            if v is not None:
                setattr(self, k, v)
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if hasattr(self, "x") and self.x is not None:
            return self.x.size(0)
        if hasattr(self, "pos") and self.pos is not None:
            return self.pos.size(0)
        return int(self.edge_index.max()) + 1

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def clone(self):
        new = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            new.__dict__[k] = v.clone() if isinstance(v, torch.Tensor) else copy.deepcopy(v)
        return new

    def to(self, device, non_blocking=False):
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                self.__dict__[k] = v.to(device, non_blocking=non_blocking)
        return self


class Batch(Data):
    """``Batch.from_data_list`` collate: cat dim 0 except edge_index (dim 1, + node offsets)."""

    def __init__(self, batch=None, **kwargs):
        super().__init__(**kwargs)
        if batch is not None:
            self.batch = batch

    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        keys = []
        for d in data_list:
            for k in d.__dict__:
                if k not in keys:
                    keys.append(k)
        sizes = [d.num_nodes for d in data_list]
        offsets = [0]
        for s in sizes:
            offsets.append(offsets[-1] + s)
        for k in keys:
            vals = [getattr(d, k, None) for d in data_list]
            if all(v is None for v in vals):
                setattr(out, k, None)
            elif isinstance(vals[0], torch.Tensor):
                if "index" in k or k == "face":
                    setattr(out, k, torch.cat([v + o for v, o in zip(vals, offsets)], dim=1))
                else:
                    setattr(out, k, torch.cat(vals, dim=0))
            else:
                setattr(out, k, list(vals))
        out.batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)]) if sizes else torch.empty(0, dtype=torch.long)
        out.ptr = torch.tensor(offsets, dtype=torch.long)
        return out

    @property
    def num_graphs(self):
        return int(self.ptr.numel()) - 1 if hasattr(self, "ptr") else int(self.batch.max()) + 1
