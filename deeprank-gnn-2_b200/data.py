"""``Data`` / ``Batch`` containers and the collate the Trainer's loader performs.

The reference uses ``torch_geometric.data.{Data,Batch}`` and
``torch_geometric.loader.DataLoader`` (reference ``deeprank2/trainer.py:17,541-557``;
``dataset.py:1044-1052`` builds each ``Data``).  Only the slice of that interface the
GNN path touches is mirrored: attribute access/assignment, ``clone()``, ``to()``,
``num_nodes``, ``Batch.from_data_list`` with PyG's collate rules:

* tensors are concatenated along dim 0, except attributes whose name contains
  ``index`` (``edge_index``), which are concatenated along dim 1 after adding the
  cumulative node offset of their graph;
* ``cluster0`` / ``cluster1`` get NO offset (per-graph local ids; that is why
  ``get_preloaded_cluster`` exists, ``utils/community_pooling.py:23-27``);
* ``batch`` (int64 [N], non-decreasing) and ``ptr`` (int64 [B+1]) are added;
* python attributes (``entry_names``) become lists.

A ``Batch`` also carries the device-side graph index (destination-sorted CSR,
source-sorted CSC, graph offsets) lazily built by ``deeprank2_b200.graph``; it is
cached on the object so the four convolutions of a forward pass and their backward
share one build.
"""
from __future__ import annotations

import copy
from typing import Any, Iterable

import torch

_OFFSET_KEYS = ("index", "face")


def _is_doubled(d, ei) -> bool:
    """Does the graph's edge list have the reference's layout (``dataset.py:944-948``: all (i, j) first, then all (j, i) in the same
    order)?  Remembered on the graph: datasets hand out the same ``Data`` objects epoch after epoch."""
    key = (ei.data_ptr(), ei._version, tuple(ei.shape))
    cached = d.__dict__.get("_doubled")
    if cached is not None and cached[0] == key:
        return cached[1]
    e = int(ei.shape[1]) if ei.dim() == 2 else -1
    ok = ei.dim() == 2 and e % 2 == 0 and bool(torch.equal(ei[:, e // 2 :], ei[:, : e // 2].flip(0)))
    d.__dict__["_doubled"] = (key, ok)
    return ok


def pool_sizes(d) -> tuple | None:
    """Sizes of the two community-pooling levels of ONE graph -- properties of the graph and its stored clustering, not of any
    weight: ``(k0, c0, e1, k1, c1)`` = id bound (max + 1) and number of distinct ids of ``cluster0``, number of distinct pooled
    edges without self loops (PyG ``pool_edge``), id bound and distinct ids of ``cluster1``.  Computed once per graph on the host
    (the dataset hands out the same ``Data`` objects epoch after epoch) so that a collated batch knows the shapes of its pooled
    tensors and the device never has to report them back.  None if the graph carries no (host) clustering."""
    c0, c1, ei = d.__dict__.get("cluster0"), d.__dict__.get("cluster1"), d.__dict__.get("edge_index")
    if not (isinstance(c0, torch.Tensor) and isinstance(c1, torch.Tensor) and isinstance(ei, torch.Tensor)) or c0.is_cuda or c1.is_cuda or ei.is_cuda:
        return None
    key = (c0.data_ptr(), c0._version, c1.data_ptr(), c1._version, ei.data_ptr(), ei._version, tuple(ei.shape))
    cached = d.__dict__.get("_pool_sizes")
    if cached is not None and cached[0] == key:
        return cached[1]
    import numpy as np

    a0, a1, e = c0.numpy().reshape(-1), c1.numpy().reshape(-1), ei.numpy()
    if (a0.size and a0.min() < 0) or (a1.size and a1.min() < 0):
        return None
    uniq0, inv0 = np.unique(a0, return_inverse=True)
    n_c0 = int(uniq0.size)
    e1 = 0
    if e.size and n_c0:
        pr, pc = inv0[e[0]].astype(np.int64), inv0[e[1]].astype(np.int64)
        keep = pr != pc
        e1 = int(np.unique(pr[keep] * n_c0 + pc[keep]).size)
    out = (int(a0.max()) + 1 if a0.size else 0, n_c0, e1, int(a1.max()) + 1 if a1.size else 0, int(np.unique(a1).size))
    d.__dict__["_pool_sizes"] = (key, out)
    return out


def _takes_node_offset(key: str) -> bool:
    return any(tag in key for tag in _OFFSET_KEYS)


class Data:
    """One graph: ``x [n,F]``, ``edge_index [2,E]`` (int64), ``edge_attr [E,Fe]``, ``y``, ``pos`` ..."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **extra: Any):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self.pos = pos
        for key, value in extra.items():
            setattr(self, key, value)

    # -- introspection
    def _fields(self) -> Iterable[str]:
        names = [k for k in self.__dict__ if not k.startswith("_")]
        pending = self.__dict__.get("_pending")
        if pending is not None:
            names += [k for k in pending[1] if not k.startswith("_")]
        return names

    def meta(self, key: str, default=None):
        return self.__dict__.get(self._META_KEY, {}).get(key, default)

    def set_meta(self, key: str, value) -> None:
        self.__dict__.setdefault(self._META_KEY, {})[key] = value

    def _peek(self, key: str):
        """The attribute's value without triggering a deferred host->device copy."""
        if key in self.__dict__:
            return self.__dict__[key]
        pending = self.__dict__.get("_pending")
        return pending[1].get(key) if pending is not None else None

    @property
    def keys(self):
        return [k for k in self._fields() if self._peek(k) is not None]

    @property
    def num_nodes(self) -> int:
        if self.x is not None:
            return int(self.x.shape[0])
        if self.pos is not None:
            return int(self.pos.shape[0])
        if self.edge_index is not None and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return 0

    @property
    def num_edges(self) -> int:
        return 0 if self.edge_index is None else int(self.edge_index.shape[1])

    @property
    def num_node_features(self) -> int:
        return 0 if self.x is None else (1 if self.x.dim() == 1 else int(self.x.shape[1]))

    @property
    def num_edge_features(self) -> int:
        return 0 if self.edge_attr is None else (1 if self.edge_attr.dim() == 1 else int(self.edge_attr.shape[1]))

    def __contains__(self, key: str) -> bool:
        return self._peek(key) is not None

    def __repr__(self) -> str:
        parts = []
        for k in self._fields():
            v = self._peek(k)
            if isinstance(v, torch.Tensor):
                parts.append(f"{k}={list(v.shape)}")
            elif v is not None:
                parts.append(f"{k}={type(v).__name__}")
        return f"{type(self).__name__}({', '.join(parts)})"

    # -- copies / movement (private caches, e.g. the graph index, are dropped on clone and moved on to())
    # host-side facts about the batch that survive clone()/to(): {"num_graphs", "max_graph_nodes"}
    _META_KEY = "_meta"
    _DEVICE_CACHES = ("_graph_index", "_block_info")
    _HOST_CACHES = ("_slab",)  # the pinned slab of pin_memory(): meaningless for copies whose tensors are no longer its views

    def clone(self):
        new = type(self).__new__(type(self))
        for k, v in self.__dict__.items():
            if k in self._DEVICE_CACHES or k in self._HOST_CACHES:
                continue
            if k == "_pending":
                new.__dict__[k] = (v[0], dict(v[1]))
                continue
            new.__dict__[k] = v.clone() if isinstance(v, torch.Tensor) else copy.deepcopy(v)
        return new

    def to(self, device, non_blocking: bool = False, only=None):
        """Move the tensors to ``device``.  With ``only`` (names) and a CUDA target, the other host tensors are NOT copied
        now: they stay on the host and travel on first attribute access (a step that never reads ``edge_attr`` or ``pos``
        never pays their PCIe time)."""
        dev = torch.device(device)
        stale = self.__dict__.pop("_pending", None)
        pending = dict(stale[1]) if stale is not None else {}
        lazy = only is not None and dev.type == "cuda"
        moved = self._slab_to(dev, non_blocking, only)
        for k, v in list(self.__dict__.items()):
            if k in moved or k in self._HOST_CACHES:
                continue
            if isinstance(v, torch.Tensor):
                if lazy and k not in only and not v.is_cuda:
                    pending[k] = self.__dict__.pop(k)
                else:
                    self.__dict__[k] = v.to(dev, non_blocking=non_blocking)
            elif k in self._DEVICE_CACHES:
                self.__dict__.pop(k)
        if lazy:
            if pending:
                self.__dict__["_pending"] = (dev, pending)
        else:
            for k, v in pending.items():
                self.__dict__[k] = v.to(dev, non_blocking=non_blocking)
        return self

    def __getattr__(self, name):
        # only reached when normal lookup fails: a tensor whose host->device copy was deferred by to(..., only=...)
        pending = self.__dict__.get("_pending")
        if pending is not None and name in pending[1]:
            value = pending[1].pop(name).to(pending[0], non_blocking=True)
            self.__dict__[name] = value
            return value
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {name!r}")

    def cuda(self, non_blocking: bool = False):
        return self.to("cuda", non_blocking=non_blocking)

    def cpu(self):
        return self.to("cpu")

    def pin_memory(self, only=None):
        """Page-lock the host tensors (all of them, or the names in ``only``: the ones an input pipeline copies ahead).

        The selected tensors are packed into ONE pinned slab (256-byte aligned slices) and become views of it: ``to(device)`` then moves
        them with a single copy instead of one DMA per tensor -- a C2 step reads seven tensors, five of them ~1 KB, and every extra
        ``cudaMemcpyAsync`` costs microseconds of a ~0.4 ms step that runs at the PCIe rate."""
        picked = [(k, v) for k, v in self.__dict__.items() if isinstance(v, torch.Tensor) and not v.is_cuda and (only is None or k in only)]
        if not picked:
            return self
        layout, total = [], 0
        for k, v in picked:
            nbytes = v.numel() * v.element_size()
            layout.append((k, total, nbytes, v.dtype, tuple(v.shape)))
            total += (nbytes + 255) // 256 * 256
        slab = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        for (k, off, nbytes, dtype, shape), (_, v) in zip(layout, picked):
            view = slab[off : off + nbytes].view(dtype).view(shape)
            view.copy_(v)
            self.__dict__[k] = view
        self.__dict__["_slab"] = (slab, layout)
        return self

    def _slab_to(self, dev, non_blocking: bool, wanted) -> set:
        """Move the pinned slab with one copy and re-create the views on the device; returns the names it covered."""
        packed = self.__dict__.get("_slab")
        if packed is None or dev.type != "cuda":
            return set()
        slab, layout = packed
        names = [k for k, off, nbytes, dtype, shape in layout]
        if any(wanted is not None and k not in wanted for k in names):
            return set()  # the slab holds tensors that are to stay on the host: fall back to per-tensor copies
        for k, off, nbytes, dtype, shape in layout:  # still the views made by pin_memory?
            v = self.__dict__.get(k)
            if not isinstance(v, torch.Tensor) or v.is_cuda or v.data_ptr() != slab.data_ptr() + off or v.dtype != dtype or tuple(v.shape) != shape:
                return set()
        dslab = slab.to(dev, non_blocking=non_blocking)
        for k, off, nbytes, dtype, shape in layout:
            self.__dict__[k] = dslab[off : off + nbytes].view(dtype).view(shape)
        self.__dict__.pop("_slab", None)
        return set(names)


class Batch(Data):
    """A disjoint union of graphs (block-diagonal adjacency)."""

    def __init__(self, batch=None, ptr=None, **kwargs: Any):
        super().__init__(**kwargs)
        self.batch = batch
        if ptr is not None:
            self.ptr = ptr

    @property
    def num_graphs(self) -> int:
        ptr = self.__dict__.get("ptr")
        if ptr is not None:
            return int(ptr.numel()) - 1
        if self.batch is None or self.batch.numel() == 0:
            return 0
        return int(self.batch.max()) + 1  # host sync: only hit for hand-made batches without ptr

    @classmethod
    def from_data_list(cls, data_list: list[Data]) -> "Batch":
        if len(data_list) == 0:
            raise ValueError("cannot collate an empty list of graphs")
        names: list[str] = []
        for d in data_list:
            for k in d._fields():
                if k not in names:
                    names.append(k)
        import numpy as np

        sizes = [d.num_nodes for d in data_list]
        ptr_np = np.zeros(len(sizes) + 1, dtype=np.int64)
        np.cumsum(np.asarray(sizes, dtype=np.int64), out=ptr_np[1:])
        ptr = torch.from_numpy(ptr_np)

        def with_node_offsets(column):
            """cat(dim=1) of per-graph index tensors with each graph's first node id added: one cat + one vectorised add."""
            widths = np.fromiter((int(v.shape[1]) for v in column), dtype=np.int64, count=len(column))
            merged = torch.cat(column, dim=1)
            if merged.dtype == torch.int64 and merged.numel():
                merged += torch.from_numpy(np.repeat(ptr_np[:-1], widths))
                return merged
            return torch.cat([v + int(o) for v, o in zip(column, ptr_np[:-1])], dim=1)

        out = cls()
        for k in names:
            column = [d.__dict__.get(k) for d in data_list]
            present = [v for v in column if v is not None]
            if not present:
                setattr(out, k, None)
            elif isinstance(present[0], torch.Tensor):
                if len(present) != len(column):
                    raise ValueError(f"attribute {k!r} is missing on some graphs of the batch")
                if _takes_node_offset(k):
                    setattr(out, k, with_node_offsets(column))
                else:
                    column = [v.unsqueeze(0) if v.dim() == 0 else v for v in column]
                    setattr(out, k, torch.cat(column, dim=0))
            else:
                setattr(out, k, list(column))
        out.batch = torch.from_numpy(np.repeat(np.arange(len(sizes), dtype=np.int64), np.asarray(sizes, dtype=np.int64)))
        out.ptr = ptr
        # what the per-graph kernels need and PyG's Batch does not keep: int32 node / edge offsets of every graph (the edges of a
        # graph are a contiguous slice of edge_index) and the graphs sorted by size, largest first (issue order of the CTAs).
        # Private attributes: they follow the batch through clone()/to()/pin_memory() but are not part of `keys`.
        e_sizes = [d.num_edges for d in data_list]
        if ptr_np[-1] < 2**31 and sum(e_sizes) < 2**31:
            eptr_np = np.zeros(len(sizes) + 1, dtype=np.int64)
            np.cumsum(np.asarray(e_sizes, dtype=np.int64), out=eptr_np[1:])
            eptr = torch.from_numpy(eptr_np)
            out.__dict__["_node_ptr32"] = ptr.to(torch.int32)
            out.__dict__["_edge_ptr32"] = eptr.to(torch.int32)
            out.__dict__["_order32"] = snake_order([e + 8 * n for e, n in zip(e_sizes, sizes)])
            # The reference stores every contact twice (dataset.py:944-948: all (i, j), then all (j, i) in the same order).  When
            # every graph has that layout the contacts are also kept ONCE (`_pairs`, what the HDF5 files hold): the per-graph
            # kernels rebuild the doubled list on the fly, so half of the edge bytes never cross PCIe.
            halves = []
            for d in data_list:
                ei = d.__dict__.get("edge_index")
                if ei is None or not _is_doubled(d, ei):
                    halves = None
                    break
                halves.append(ei[:, : int(ei.shape[1]) // 2])
            if halves is not None:
                local = torch.cat(halves, dim=1)
                if max(sizes) <= 65536:
                    # ... and packed: one 32-bit word per contact, (i | j << 16) with ids local to the graph -- what the per-graph step
                    # kernel needs and all it reads (DRK_EDGES_LOCAL_PAIRS16): 4 bytes per contact over PCIe instead of 16 (32 doubled)
                    words = (local[0] | (local[1] << 16)).numpy().astype("uint32").view("int32")
                    out.__dict__["_pairs16"] = torch.from_numpy(words.copy())
                if local.dtype == torch.int64:
                    local = local + torch.from_numpy(np.repeat(ptr_np[:-1], np.asarray(e_sizes, dtype=np.int64) // 2))
                else:
                    local = torch.cat([h + int(o) for h, o in zip(halves, ptr_np[:-1])], dim=1)
                out.__dict__["_pairs"] = local.contiguous()
                out.__dict__["_pair_ptr32"] = (eptr // 2).to(torch.int32)
        out.__dict__[cls._META_KEY] = {"num_graphs": len(sizes), "max_graph_nodes": max(sizes), "max_graph_edges": max(e_sizes), "num_edges_total": sum(e_sizes)}
        out._attach_pool_sizes([pool_sizes(d) for d in data_list])
        return out

    def _attach_pool_sizes(self, per_graph) -> None:
        """Batch-level shapes of the community-pooling chain from the per-graph sizes (``pool_sizes``): totals in the meta dict
        (``pool``: K0 C0 E1 KK K1 C1) and, as tensors that travel with the batch, the first pooled node of every graph
        (``_pool_cptr`` int64 [B+1]) and the first dense pooled-pair id of every graph (``_pool_kkptr`` int64 [B+1], blocks of C_g^2)."""
        import numpy as np

        if not per_graph or any(p is None for p in per_graph):
            return
        m = np.asarray(per_graph, dtype=np.int64).reshape(len(per_graph), 5)
        cptr = np.zeros(len(per_graph) + 1, dtype=np.int64)
        np.cumsum(m[:, 1], out=cptr[1:])
        kkptr = np.zeros(len(per_graph) + 1, dtype=np.int64)
        np.cumsum(m[:, 1] * m[:, 1], out=kkptr[1:])
        if kkptr[-1] + 4096 >= 2**31 or m[:, 0].sum() >= 2**31:
            return
        eptr = np.zeros(len(per_graph) + 1, dtype=np.int64)
        np.cumsum(m[:, 2], out=eptr[1:])
        self.__dict__["_pool_cptr"] = torch.from_numpy(cptr)
        self.__dict__["_pool_kkptr"] = torch.from_numpy(kkptr)
        self.__dict__["_pool_eptr32"] = torch.from_numpy(eptr.astype(np.int32))  # edges of the pooled graphs: contiguous slices, like _edge_ptr32
        c1ptr = np.zeros(len(per_graph) + 1, dtype=np.int64)
        np.cumsum(m[:, 4], out=c1ptr[1:])
        self.__dict__["_pool_c1ptr"] = torch.from_numpy(c1ptr)  # first level-1 cluster of every graph
        self.__dict__.setdefault(self._META_KEY, {})["pool"] = {"K0": int(m[:, 0].sum()), "C0": int(m[:, 1].sum()), "E1": int(m[:, 2].sum()), "KK": int(kkptr[-1]),
                                                                 "K1": int(m[:, 3].sum()), "C1": int(m[:, 4].sum()), "max_C0": int(m[:, 1].max()), "max_E1": int(m[:, 2].max()),
                                                                 "max_K0": int(m[:, 0].max()), "max_K1": int(m[:, 3].max())}


STEP_CTAS = 148  # CTAs of the per-graph kernels = SMs of a B200 (drk_ginet_step_ctas)


def snake_order(work: list[int], ctas: int = STEP_CTAS) -> torch.Tensor:
    """Slot -> graph id for the per-graph kernels: CTA b processes slots b, b + ctas, b + 2 ctas, ...

    Longest-processing-time-first: graphs are taken by decreasing work and each goes to the CTA with the least work so far, so
    all CTAs finish at about the same time -- with 256 graphs on 148 CTAs the 40 largest graphs get a CTA of their own and the
    other 216 are paired largest-with-smallest.  CTAs are then numbered by decreasing graph count (a CTA's k-th graph sits in
    slot b + k * ctas, so the CTAs that run an extra round must be the first ones).  Ties keep graph order: deterministic."""
    import heapq

    import numpy as np

    n = len(work)
    g = max(1, min(ctas, n))
    if n <= 2 * g:
        # closed form of the same schedule for at most two rounds (the per-step path of resident graph sets): k = n - g CTAs run two
        # graphs -- the k smallest of the g largest, each paired with one of the k smallest overall (smallest first graph with the
        # largest second) -- and come first; the g - k largest graphs run alone.
        idx = np.argsort(-np.asarray(work, dtype=np.int64), kind="stable")
        k = max(n - g, 0)
        return torch.from_numpy(np.concatenate((idx[g - k : g], idx[: g - k], idx[g:][::-1])).astype(np.int32))
    idx = sorted(range(n), key=lambda i: (-work[i], i))
    heap = [(0, b) for b in range(g)]  # (load, cta)
    mine: list[list[int]] = [[] for _ in range(g)]
    for i in idx:
        load, b = heapq.heappop(heap)
        mine[b].append(i)
        heapq.heappush(heap, (load + work[i], b))
    rank = sorted(range(g), key=lambda b: (-len(mine[b]), b))
    order = [-1] * n
    rounds = max((len(m) for m in mine), default=0)
    slot_of_round = 0
    for k in range(rounds):
        members = [b for b in rank if len(mine[b]) > k]
        for pos, b in enumerate(members):
            order[slot_of_round + pos] = mine[b][k]
        slot_of_round += len(members)
    return torch.tensor(order, dtype=torch.int32)


def collate(data_list: list[Data]) -> Batch:
    """``collate_fn`` for ``torch.utils.data.DataLoader`` (what PyG's ``Collater`` does for ``Data``)."""
    return Batch.from_data_list(data_list)
