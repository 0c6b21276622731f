"""Device-side graph index of a batch: destination-sorted CSR, source-sorted CSC, graph offsets.

The reference never builds an index: every layer does ``row, col = edge_index`` and lets
``x[col]`` / ``torch_scatter.scatter_sum(h, row)`` walk the raw int64 edge list
(``ginet.py:41-58``, ``vanilla_gnn.py:28-35``, ``foutnet.py:56-58``) and lets ``scatter_mean(x,
batch)`` rediscover the graph boundaries (``ginet_nocluster.py:103``).  Here the Trainer's
collate path builds that structure ONCE per batch on the GPU (``drk_graph_index_build`` +
``drk_batch_offsets``) and all convolutions of the forward and backward pass share it.

Index arrays are int32 and bit-exact against ``torch.sort(stable=True)`` / ``bincount`` /
``cumsum`` (``oracle/restate.py:graph_csr``): inside a destination the edges keep ascending
edge id, which is also the order in which the reference's CPU ``scatter_add_`` adds them.
"""
from __future__ import annotations

import torch

from . import _lib


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: deeprank2_b200 has no CPU path (got {t.device})")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of the current stream of the current device (the raw getter costs ~0.3 us against ~15 us for building a
    ``torch.cuda.Stream`` object; an eager train step of the layer kernels asks ~30 times)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


_workspaces: dict = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Scratch buffer re-used by every call on (device, current stream); stream order makes that safe."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, _raw_stream(index) if _raw_stream is not None else torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


class GraphIndex:
    """CSR (by destination ``edge_index[0]``) + CSC (by source ``edge_index[1]``) + graph offsets."""

    __slots__ = (
        "num_nodes", "num_edges", "num_graphs", "device",
        "rowptr", "colidx", "perm", "colptr", "rowidx", "permT",
        "graph_ptr", "batch32", "status", "_storage", "_key", "_degree", "_slot_map", "_attr_csr", "max_graph_nodes", "order",
    )

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)

    # ------------------------------------------------------------------ construction
    @classmethod
    def build(cls, edge_index: torch.Tensor, num_nodes: int, batch: torch.Tensor | None = None, num_graphs: int | None = None, with_csc: bool = True,
              blocks: tuple | None = None, with_perm: bool = True) -> "GraphIndex":
        """``blocks`` = (node_ptr int32 [B+1], edge_ptr int32 [B+1], max_graph_nodes, max_graph_edges) of a collated batch (the edges
        of a graph are one contiguous slice of ``edge_index``): the index is then built per graph in shared memory
        (``drk_graph_index_build_blocked``, ~10x faster than the global sort, bit-identical result).  ``with_perm=False`` (only
        honoured by the per-graph builder, and only without the CSC half): ``perm`` -- the original edge id of every CSR slot -- is not
        written and ``gi.perm`` is None; for callers that only aggregate node rows (inference without edge attributes)."""
        lib = _lib.load()
        _require_cuda(edge_index, "edge_index")
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise TypeError(f"edge_index must be int64 [2, E], got {edge_index.dtype} {tuple(edge_index.shape)}")
        edge_index = edge_index.contiguous()
        dev = edge_index.device
        n, e = int(num_nodes), int(edge_index.shape[1])
        if batch is not None:
            _require_cuda(batch, "batch")
            if batch.dtype != torch.int64 or batch.numel() != n:
                raise TypeError(f"batch must be int64 [{n}], got {batch.dtype} {tuple(batch.shape)}")
            if num_graphs is None:
                # same as scatter_mean without dim_size (ginet_nocluster.py:103): one host sync.
                # Batches made by Batch.from_data_list carry `ptr`, so the Trainer path never gets here.
                num_graphs = int(batch.max()) + 1 if n > 0 else 0
        b = int(num_graphs) if num_graphs is not None else 0
        gi = cls()
        gi.num_nodes, gi.num_edges, gi.num_graphs, gi.device = n, e, b, dev
        # one allocation for every int32 array (each sub-array 16-byte aligned)
        def pad(v):
            return (v + 3) // 4 * 4
        sizes = [pad(n + 1), pad(e), pad(e)]
        if with_csc:
            sizes += [pad(n + 1), pad(e), pad(e)]
        if batch is not None:
            sizes += [pad(b + 1), pad(n)]
        sizes += [4]
        storage = torch.empty(sum(sizes), dtype=torch.int32, device=dev)
        gi._storage = storage
        views, off = [], 0
        for s in sizes:
            views.append(storage[off : off + s])
            off += s
        gi.rowptr, gi.colidx, gi.perm = views[0][: n + 1], views[1][:e], views[2][:e]
        k = 3
        if with_csc:
            gi.colptr, gi.rowidx, gi.permT = views[3][: n + 1], views[4][:e], views[5][:e]
            k = 6
        if batch is not None:
            gi.graph_ptr, gi.batch32 = views[k][: b + 1], views[k + 1][:n]
        gi.status = views[-1][:1]
        gi.status.zero_()
        use_blocks = (
            blocks is not None and b > 0 and blocks[0].is_cuda and blocks[1].is_cuda and int(blocks[0].numel()) == b + 1
            and bool(lib.drk_graph_index_blocked_supported(int(blocks[2]), int(blocks[3])))
        )
        if use_blocks:
            if not with_perm and not with_csc:
                gi.perm = None
            with torch.cuda.device(dev):
                rc = lib.drk_graph_index_build_blocked(
                    edge_index.data_ptr(), e, n, blocks[0].data_ptr(), blocks[1].data_ptr(), b, int(blocks[2]), int(blocks[3]),
                    gi.rowptr.data_ptr(), gi.colidx.data_ptr(), gi.perm.data_ptr() if gi.perm is not None else None,
                    gi.colptr.data_ptr() if with_csc else None, gi.rowidx.data_ptr() if with_csc else None, gi.permT.data_ptr() if with_csc else None,
                    gi.status.data_ptr(), stream_ptr(),
                )
                _lib.check(rc, "drk_graph_index_build_blocked")
                if batch is not None:
                    rc = lib.drk_batch_offsets(batch.contiguous().data_ptr(), n, b, gi.graph_ptr.data_ptr(), gi.batch32.data_ptr(), gi.status.data_ptr(), stream_ptr())
                    _lib.check(rc, "drk_batch_offsets")
            return gi
        with torch.cuda.device(dev):
            ws_bytes = lib.drk_graph_index_workspace_bytes(e, n)
            ws = workspace(ws_bytes, dev)
            rc = lib.drk_graph_index_build(
                edge_index.data_ptr(), e, n,
                gi.rowptr.data_ptr(), gi.colidx.data_ptr(), gi.perm.data_ptr(),
                gi.colptr.data_ptr() if with_csc else None,
                gi.rowidx.data_ptr() if with_csc else None,
                gi.permT.data_ptr() if with_csc else None,
                gi.status.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(),
            )
            _lib.check(rc, "drk_graph_index_build")
            if batch is not None:
                rc = lib.drk_batch_offsets(batch.contiguous().data_ptr(), n, b, gi.graph_ptr.data_ptr(), gi.batch32.data_ptr(), gi.status.data_ptr(), stream_ptr())
                _lib.check(rc, "drk_batch_offsets")
        return gi

    def degree(self) -> torch.Tensor:
        """float32 [N]: number of edges per destination node (cached)."""
        if self._degree is None:
            self._degree = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.float32)
        return self._degree

    def slot_map(self) -> torch.Tensor:
        """int32 [E]: CSR slot of the edge in every CSC slot (cached; ``drk_attn_slot_map``).  Per-edge state kept in CSR order by
        the destination kernels is reached through it by the source kernels."""
        if self._slot_map is None:
            if self.colptr is None:
                raise RuntimeError("the slot map needs the CSC half of the graph index (build it with with_csc=True)")
            out = torch.empty(2 * max(self.num_edges, 1), dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                rc = _lib.load().drk_attn_slot_map(self.perm.data_ptr(), self.permT.data_ptr(), self.num_edges, out[self.num_edges :].data_ptr(),
                                                   out.data_ptr(), stream_ptr())
            _lib.check(rc, "drk_attn_slot_map")
            self._slot_map = out[: self.num_edges]
        return self._slot_map

    def attr_in_slot_order(self, edge_attr: torch.Tensor) -> torch.Tensor:
        """``edge_attr[perm]`` (float32 [E, Fe], rows in CSR-slot order), cached per edge_attr tensor: the four convolutions of a
        GINet and their backward passes share one gather per batch."""
        key = (edge_attr.data_ptr(), edge_attr._version, tuple(edge_attr.shape))
        if self._attr_csr is None or self._attr_csr[0] != key:
            from . import ops

            self._attr_csr = (key, ops.gather_rows(edge_attr, self.perm))
        return self._attr_csr[1]

    # ------------------------------------------------------------------ validation (host sync: call it off the hot path)
    def check(self) -> None:
        flags = int(self.status.item())
        if flags & _lib.STATUS_INDEX_RANGE:
            raise IndexError("edge_index / batch contains an id outside [0, num_nodes) resp. [0, num_graphs)")
        if flags & _lib.STATUS_UNSORTED:
            raise ValueError("batch vector is not sorted: graphs of a Batch must occupy consecutive node ranges")


def graph_index(data, with_csc: bool = True, with_perm: bool = True) -> GraphIndex:
    """The (cached) :class:`GraphIndex` of a ``Batch``/``Data`` living on the GPU."""
    ei = data.edge_index
    key = (ei.data_ptr(), ei._version, tuple(ei.shape), str(ei.device), bool(with_csc))
    cached = data.__dict__.get("_graph_index")
    if cached is not None and cached._key[:4] == key[:4] and (cached.colptr is not None or not with_csc) and (cached.perm is not None or not with_perm):
        return cached
    batch = getattr(data, "batch", None)
    ptr = data.__dict__.get("ptr")
    num_graphs = int(ptr.numel()) - 1 if ptr is not None else data.__dict__.get("_num_graphs")
    if num_graphs is None and hasattr(data, "meta"):
        num_graphs = data.meta("num_graphs")
    # collated batches know where each graph's nodes and edges start: per-graph index build in shared memory
    blocks = None
    node_ptr, edge_ptr = data.__dict__.get("_node_ptr32"), data.__dict__.get("_edge_ptr32")
    meta = data.__dict__.get(getattr(data, "_META_KEY", "_meta"), {})
    if (node_ptr is not None and edge_ptr is not None and batch is not None and meta.get("num_edges_total") == int(ei.shape[1])
            and meta.get("max_graph_nodes") is not None and meta.get("max_graph_edges") is not None):
        blocks = (node_ptr, edge_ptr, meta["max_graph_nodes"], meta["max_graph_edges"])
    gi = GraphIndex.build(ei, data.num_nodes, batch=batch, num_graphs=num_graphs, with_csc=with_csc, blocks=blocks, with_perm=with_perm or with_csc)
    gi._key = key
    gi.max_graph_nodes = meta.get("max_graph_nodes")  # known on the host for collated batches: lets the aggregation pick the tiled kernel
    # issue order of the per-graph kernels (largest graphs first, data.py:snake_order); only meaningful for the batch it was made for
    order = getattr(data, "_order32", None)
    if order is not None and order.is_cuda and order.device == ei.device and gi.num_graphs and int(order.numel()) == gi.num_graphs:
        gi.order = order
    data.__dict__["_graph_index"] = gi
    return gi


def max_graph_nodes(data, gi: GraphIndex) -> int:
    """Largest graph of the batch (sizes shared memory of the per-graph kernels).  Known on the host for batches made
    by ``Batch.from_data_list``; otherwise read back once (one host sync) and remembered on the batch."""
    known = data.meta("max_graph_nodes") if hasattr(data, "meta") else None
    if known is None:
        known = data.__dict__.get("_max_graph_nodes")
    if known is None:
        if gi.graph_ptr is None or gi.num_graphs == 0:
            known = gi.num_nodes
        else:
            known = int((gi.graph_ptr[1:] - gi.graph_ptr[:-1]).max().item())
        if hasattr(data, "set_meta"):
            data.set_meta("max_graph_nodes", known)
        else:
            data.__dict__["_max_graph_nodes"] = known
    return int(known)
