"""The whole GINet step as one per-graph kernel (``csrc/drk_ginet_step.cu``).

``Trainer._epoch``'s loop body (reference ``deeprank2/trainer.py:682-694``) is, for the benchmark model
``ginet_nocluster.GINet`` with ``MSELoss`` / ``CrossEntropyLoss``::

    pred = model(batch); loss = lossfunction(pred, y); loss.backward(); optimizer.step()

:class:`GINetFusedStep` performs the first three in TWO kernel launches (one CTA per graph: graph index,
both convolution branches, readout, head, loss term and the full backward pass in shared memory; then a
finalize kernel that sums the per-graph gradient contributions in graph order) and hands the gradients to the
unchanged ``torch.optim`` optimizer.  :func:`ginet_infer` is the forward-only variant used under
``torch.no_grad()``.  Both need a *collated* batch (edges of a graph contiguous in ``edge_index``), which is what
``Batch.from_data_list`` / PyG's collate produce; the kernels verify it edge by edge and raise
``DRK_STATUS_CROSS_GRAPH`` otherwise.

There is no CPU path: everything here raises if the extension is missing or a tensor is not on a CUDA device.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib
from .graph import stream_ptr, workspace


def _p(t):
    return None if t is None else t.data_ptr()


class BlockInfo:
    """Per-graph offsets of a collated batch on the device: ``node_ptr``/``edge_ptr`` int32 [B+1], issue ``order`` int32 [B]
    (or None), the largest graph's node / (directed) edge count, the batch's status word, and the edge tensor the kernels read:
    ``edges`` int64 [2, M] with ``layout`` EDGES_DIRECTED (= ``edge_index``) or EDGES_UNDIRECTED_PAIRS (each contact once), or int32 [M]
    with EDGES_LOCAL_PAIRS16 (each contact once as one packed word of graph-local ids)."""

    __slots__ = ("node_ptr", "edge_ptr", "order", "num_graphs", "max_nodes", "max_edges", "status", "edges", "layout", "by_slot")


def block_info(data) -> BlockInfo:
    """The (cached) :class:`BlockInfo` of a device-resident ``Batch``.

    Batches made by ``Batch.from_data_list`` carry the offsets (computed by the collate on the host, moved with the batch) and,
    when every graph has the reference's doubled edge layout, the contacts once (``_pairs``): the kernels then read those.
    For any other batch the offsets are derived on the device from ``batch`` and ``edge_index`` (``drk_batch_offsets`` +
    ``drk_edge_ptr``) with one host read-back of the largest graph size, remembered on the batch."""
    lib = _lib.load()
    d = data.__dict__
    pairs, pair_ptr = d.get("_pairs16"), d.get("_pair_ptr32")
    packed = pairs is not None and pairs.is_cuda
    if not packed:
        pairs = d.get("_pairs")
    node_ptr, edge_ptr = d.get("_node_ptr32"), d.get("_edge_ptr32")
    meta = d.get(data._META_KEY, {}) if hasattr(data, "_META_KEY") else {}
    use_pairs = (
        pairs is not None and pair_ptr is not None and node_ptr is not None and pairs.is_cuda and pair_ptr.is_cuda and node_ptr.is_cuda
        and meta.get("num_edges_total") == 2 * int(pairs.shape[-1]) and meta.get("max_graph_nodes") is not None
    )
    packed = packed and use_pairs
    ei = pairs if use_pairs else data.edge_index  # (a lazily transferred edge_index is only touched when it is needed)
    if not ei.is_cuda:
        raise RuntimeError(f"the batch must live on a CUDA device: deeprank2_b200 has no CPU path (got {ei.device})")
    key = (ei.data_ptr(), ei._version, tuple(ei.shape), str(ei.device), use_pairs, packed)
    cached = d.get("_block_info")
    if cached is not None and cached[0] == key:
        return cached[1]
    info = BlockInfo()
    info.by_slot = False
    dev = ei.device
    info.edges = ei if ei.is_contiguous() else ei.contiguous()
    info.layout = (_lib.EDGES_LOCAL_PAIRS16 if packed else _lib.EDGES_UNDIRECTED_PAIRS) if use_pairs else _lib.EDGES_DIRECTED
    collated = use_pairs or (
        node_ptr is not None and edge_ptr is not None and node_ptr.is_cuda and edge_ptr.is_cuda
        and meta.get("num_edges_total") == int(ei.shape[-1]) and meta.get("max_graph_nodes") is not None
    )
    if collated:
        info.node_ptr, info.edge_ptr = node_ptr, (pair_ptr if use_pairs else edge_ptr)
        order = d.get("_order32")
        info.order = order if order is not None and order.is_cuda else None
        info.num_graphs = int(node_ptr.numel()) - 1
        info.max_nodes, info.max_edges = int(meta["max_graph_nodes"]), int(meta["max_graph_edges"])
    else:
        from .graph import graph_index

        g = graph_index(data, with_csc=False)
        if g.graph_ptr is None:
            raise ValueError("the batch has no `batch` vector: per-graph kernels need the graph boundaries")
        info.node_ptr = g.graph_ptr
        info.num_graphs = g.num_graphs
        info.edge_ptr = torch.empty(info.num_graphs + 1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.drk_edge_ptr(_p(ei), int(ei.shape[1]), _p(info.node_ptr), info.num_graphs, _p(info.edge_ptr), stream_ptr()), "drk_edge_ptr")
        info.order = None
        sizes = torch.stack([(info.node_ptr[1:] - info.node_ptr[:-1]).max(), (info.edge_ptr[1:] - info.edge_ptr[:-1]).max()]) if info.num_graphs else torch.zeros(2)
        info.max_nodes, info.max_edges = (int(v) for v in sizes.tolist())  # one host sync, remembered on the batch
    info.status = _status_word(dev)
    d["_block_info"] = (key, info)
    return info


_STATUS_WORDS: dict = {}


def _status_word(dev) -> torch.Tensor:
    """ONE status word per device, shared by every batch descriptor (the kernels OR their data-dependent faults into it): a fresh
    batch costs no fill launch; ``check_status`` reads it and clears it."""
    dev = torch.device(dev)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    word = _STATUS_WORDS.get(key)
    if word is None:
        word = _STATUS_WORDS[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return word


def check_status(info: BlockInfo) -> None:
    """Host sync: raise if a per-graph kernel flagged the batch (call off the hot path, e.g. at the end of a pass)."""
    flags = int(info.status.item())
    if flags:
        info.status.zero_()  # the word is shared by all batches of the device: report once
    if flags & _lib.STATUS_CROSS_GRAPH:
        raise ValueError("edge_index is not grouped by graph (an edge joins two graphs of the batch): not a collated batch")
    if flags & _lib.STATUS_INDEX_RANGE:
        raise IndexError("a graph of the batch exceeds the sizes the kernel was planned for, or a class index is out of range")


def _standard_ginet(model) -> bool:
    from .neuralnets.gnn.ginet_nocluster import GINet

    if not isinstance(model, GINet) or not model._stackable():
        return False
    fi = model.conv1.fc.weight.shape[1]
    return (
        tuple(model.conv1.fc.weight.shape) == (16, fi)
        and tuple(model.conv2.fc.weight.shape) == (32, 16)
        and tuple(model.fc1.weight.shape) == (128, 64)
        and model.fc1.bias is not None
        and model.fc2.weight.shape[1] == 128
        and model.fc2.bias is not None
        and 1 <= model.fc2.weight.shape[0] <= 8
        and fi <= 64
        and all(p.dtype == torch.float32 and p.is_cuda and p.is_contiguous() for p in model.parameters())
    )


def _loss_kind(loss_fn):
    if isinstance(loss_fn, nn.MSELoss) and loss_fn.reduction == "mean":
        return _lib.LOSS_MSE
    if isinstance(loss_fn, nn.CrossEntropyLoss) and loss_fn.reduction == "mean" and loss_fn.weight is None and getattr(loss_fn, "label_smoothing", 0.0) == 0.0:
        return _lib.LOSS_CROSS_ENTROPY
    return None


def step_supported(model, data) -> bool:
    """True if ``model`` is the reference GINet architecture on the GPU and every graph of ``data`` fits the kernel's plan."""
    if not _standard_ginet(model) or not data.x.is_cuda or data.x.dtype != torch.float32 or data.x.dim() != 2:
        return False
    if data.x.shape[1] != model.conv1.fc.weight.shape[1] or data.x.requires_grad or data.x.stride(1) != 1:
        return False
    info = block_info(data)
    return bool(_lib.load().drk_ginet_step_supported(int(data.x.shape[1]), int(model.fc2.weight.shape[0]), info.max_nodes, info.max_edges))


def _call_step(model, data, info, *, train, loss_kind, target, inv_loss_count, dropout_p, seed, state, pred, loss, grads, adam=None, peers=None):
    lib = _lib.load()
    x = data.x
    fi = int(x.shape[1])
    out_dim = int(model.fc2.weight.shape[0])
    ei = info.edges
    with torch.cuda.device(x.device):
        ws_bytes = lib.drk_ginet_step_workspace_bytes(fi, out_dim, info.num_graphs, info.max_nodes, info.max_edges) if train else 0
        ws = workspace(ws_bytes, x.device) if train else None
        rc = lib.drk_ginet_step(
            _p(x), x.stride(0), fi, _p(ei), int(ei.shape[-1]), int(info.layout), _p(info.node_ptr), _p(info.edge_ptr), _p(info.order), 1 if getattr(info, "by_slot", False) else 0,
            info.num_graphs, info.max_nodes, info.max_edges,
            _p(model.conv1.fc.weight), _p(model.conv1_ext.fc.weight), _p(model.conv2.fc.weight), _p(model.conv2_ext.fc.weight),
            _p(model.fc1.weight), _p(model.fc1.bias), _p(model.fc2.weight), _p(model.fc2.bias), out_dim,
            int(loss_kind), _p(target), float(inv_loss_count), float(dropout_p), int(seed) & (2**64 - 1), _p(state), 1 if train else 0,
            _p(pred), _p(loss), *[_p(g) for g in grads], ctypes.byref(adam) if adam is not None else None, ctypes.byref(peers) if peers is not None else None, _p(info.status), _p(ws), ws.numel() if ws is not None else 0, stream_ptr(),
        )
    _lib.check(rc, "drk_ginet_step")


def ginet_infer(model, data, selection=None) -> torch.Tensor:
    """``model(data)`` for the reference GINet without autograd: one kernel, [B, out] predictions.  With ``selection`` (from
    ``ResidentGraphSet.select``) ``data`` is the set's packed batch and the predictions come back in slot order."""
    info = selection if selection is not None else block_info(data)
    pred = torch.empty((info.num_graphs, int(model.fc2.weight.shape[0])), dtype=torch.float32, device=data.x.device)
    _call_step(model, data, info, train=False, loss_kind=_lib.LOSS_MSE, target=None, inv_loss_count=0.0, dropout_p=0.0, seed=0, state=None,
               pred=pred, loss=None, grads=[None] * 8)
    return pred


class GINetFusedStep:
    """``loss, pred = step(batch)``: forward + loss + backward of the reference GINet in two launches, then ``optimizer.step()``.

    The gradients land in one flat fp32 buffer whose views are the parameters' ``.grad`` (the dead ``fc_edge_attr`` /
    ``fc_attention`` parameters keep exact-zero gradients, as in the reference).  With ``world_size > 1`` the flat buffer
    is summed over ranks with one NCCL all-reduce; ``global_size`` (graphs in the global mini-batch) scales the loss terms
    so that the sum is the single-process gradient, ragged tails included.
    ``target_fn(batch) -> tensor`` supplies the targets (float [B(,out)] for MSE, int64 class indices for cross entropy).
    """

    #: the batch tensors a step reads (what an input pipeline has to copy ahead; GINet's attention is the identity, so
    #: ``edge_attr`` is never read, and the readout uses the graph offsets instead of ``batch``)
    FIELDS = ("x", "_pairs16", "_pair_ptr32", "y", "_node_ptr32", "_edge_ptr32", "_order32")
    #: the same for batches whose graphs are not in the doubled layout (no ``_pairs``): the full directed edge list
    FIELDS_DIRECTED = ("x", "edge_index", "y", "_node_ptr32", "_edge_ptr32", "_order32")

    def __init__(self, model, optimizer, loss_fn, target_fn=None, world_size: int = 1, group=None, seed: int | None = None):
        if not _standard_ginet(model):
            raise ValueError("GINetFusedStep needs the reference ginet_nocluster.GINet architecture on a CUDA device")
        kind = _loss_kind(loss_fn)
        if kind is None:
            raise ValueError(f"GINetFusedStep supports MSELoss / CrossEntropyLoss (mean reduction, no class weights), got {loss_fn}")
        self.model, self.optimizer, self.loss_fn, self.kind = model, optimizer, loss_fn, kind
        self.target_fn = target_fn or (lambda b: b.y)
        self.world, self.group = int(world_size), group
        dev = model.fc1.weight.device
        self.params = list(model.parameters())
        self.flat_grad = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat_grad[off : off + p.numel()].view_as(p))
            off += p.numel()
        by_id = {id(p): v for p, v in zip(self.params, self.views)}
        m = model
        self.grads = [by_id[id(t)] for t in (m.conv1.fc.weight, m.conv1_ext.fc.weight, m.conv2.fc.weight, m.conv2_ext.fc.weight, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)]
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.int64, device=dev)  # [0] steps done (dropout stream), [1] finalize scratch, [2] exchange epoch
        self._peers = self._peer_exchange() if self.world > 1 else None
        self._adam = self._adam_descriptor() if (self.world == 1 or self._peers is not None) else None
        self.seed = int(torch.initial_seed() if seed is None else seed)
        self._pred = {}

    @staticmethod
    def supports(model, loss_fn, batch=None) -> bool:
        return _standard_ginet(model) and _loss_kind(loss_fn) is not None and (batch is None or step_supported(model, batch))

    def _peer_exchange(self):
        """Symmetric buffers for the in-kernel gradient all-reduce (``DrkPeers``): every rank's slot buffer (one (value, epoch) word
        per gradient element, sending rank and epoch parity; the flag array is reserved and no longer read by the kernel)
        mapped into every process (``torch.distributed._symmetric_memory``: CUDA VMM handles over NVLink).  Returns None
        -- and the step falls back to one NCCL all-reduce -- if the rendezvous is not available (``DRK_NO_PEER_EXCHANGE=1``
        forces that)."""
        import os

        import torch.distributed as dist

        if os.environ.get("DRK_NO_PEER_EXCHANGE") or self.world > 8:
            return None
        try:
            import torch.distributed._symmetric_memory as symm

            lib = _lib.load()
            dev = self.flat_grad.device
            fi, out_dim = int(self.model.conv1.fc.weight.shape[1]), int(self.model.fc2.weight.shape[0])
            total = int(lib.drk_ginet_step_exchange_floats(fi, out_dim))
            n_flags = self.world * ((total + 31) // 32)
            group = self.group if self.group is not None else dist.group.WORLD
            n_words = 4 * self.world * total  # two epochs x one slot array per sending rank x (value, epoch)
            self._sym_grad = symm.empty(n_words, dtype=torch.float32, device=dev)
            self._sym_flags = symm.empty(n_flags, dtype=torch.int32, device=dev)
            self._sym_grad.zero_()
            self._sym_flags.zero_()
            h_grad = symm.rendezvous(self._sym_grad, group)
            h_flags = symm.rendezvous(self._sym_flags, group)
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)  # everybody's flags are zero before anybody's first step
            peers = _lib.Peers()
            peers.world, peers.rank = self.world, int(h_grad.rank)
            peers.capacity, peers.flag_capacity = n_words, n_flags
            for q in range(self.world):
                peers.grad_buf[q] = int(h_grad.buffer_ptrs[q])
                peers.flags[q] = int(h_flags.buffer_ptrs[q])
            self._sym_handles = (h_grad, h_flags)
            return peers
        except Exception as exc:  # noqa: BLE001 - any failure of the optional fast path means "use NCCL"
            import warnings

            warnings.warn(f"peer-memory gradient exchange unavailable ({type(exc).__name__}: {exc}); using one NCCL all-reduce per step", stacklevel=2)
            return None

    def _adam_descriptor(self):
        """``DrkAdam`` over torch.optim.Adam's own state tensors, or None if the optimizer is anything else (then
        ``optimizer.step()`` runs as usual).  Needs ``capturable=True`` or ``fused=True`` (step counters on the device)."""
        opt = self.optimizer
        if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
            return None
        grp = opt.param_groups[0]
        if grp.get("amsgrad") or grp.get("maximize") or grp.get("differentiable") or not (grp.get("capturable") or grp.get("fused")):
            return None
        if not isinstance(grp["lr"], float) or {id(p) for p in grp["params"]} != {id(p) for p in self.params}:
            return None
        for p in self.params:  # lazy state initialisation, as torch.optim.Adam._init_group does on the first step()
            st = opt.state[p]
            if len(st) == 0:
                st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if not (st["step"].is_cuda and st["step"].dtype == torch.float32 and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()):
                return None
        desc = _lib.Adam()
        desc.lr, (desc.beta1, desc.beta2), desc.eps, desc.weight_decay = grp["lr"], grp["betas"], grp["eps"], grp["weight_decay"]
        m = self.model
        live = [m.conv1.fc.weight, m.conv1_ext.fc.weight, m.conv2.fc.weight, m.conv2_ext.fc.weight, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]
        live_ids = {id(p) for p in live}
        dead = [p for p in self.params if id(p) not in live_ids]
        if len(dead) > 8:
            return None

        def fill(slot, p):
            st = opt.state[p]
            slot.param, slot.exp_avg, slot.exp_avg_sq, slot.step, slot.numel = p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["step"].data_ptr(), p.numel()

        for i, p in enumerate(live):
            fill(desc.live[i], p)
        for i, p in enumerate(dead):
            fill(desc.dead[i], p)
        desc.num_dead = len(dead)
        st0 = opt.state[self.params[0]]
        self._adam_key = (st0["step"].data_ptr(), st0["exp_avg"].data_ptr())
        return desc

    def forward_backward(self, batch, global_size: int | None = None, adam=None, selection=None):
        """Everything but the optimizer: returns (loss, pred); ``p.grad`` of every parameter is set.

        ``selection`` (a :class:`BlockInfo` from ``ResidentGraphSet.select``) runs the step on a list of graph ids of a
        device-resident graph set instead of a collated batch: ``batch`` is then the set's packed batch, targets stay indexed
        by graph id and the predictions come back in the selection's slot order."""
        info = selection if selection is not None else block_info(batch)
        out_dim = int(self.model.fc2.weight.shape[0])
        target = self.target_fn(batch)
        n_targets = int(block_info(batch).num_graphs) if selection is not None else info.num_graphs
        if self.kind == _lib.LOSS_MSE:
            target = target.to(torch.float32).contiguous()
            if target.numel() != n_targets * out_dim:
                raise ValueError(f"MSELoss target has {target.numel()} elements, predictions {n_targets * out_dim}")
            count = (global_size or info.num_graphs) * out_dim
        else:
            target = target.to(torch.int64).contiguous()
            if target.numel() != n_targets:
                raise ValueError(f"CrossEntropyLoss target has {target.numel()} elements for {n_targets} graphs")
            count = global_size or info.num_graphs
        key = (info.num_graphs, out_dim)
        pred = self._pred.get(key)
        if pred is None:
            pred = self._pred[key] = torch.empty(key, dtype=torch.float32, device=batch.x.device)
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v
        drop = float(self.model.dropout) if self.model.training else 0.0
        _call_step(self.model, batch, info, train=True, loss_kind=self.kind, target=target, inv_loss_count=1.0 / max(count, 1), dropout_p=drop,
                   seed=self.seed, state=self.state, pred=pred, loss=self.loss, grads=self.grads, adam=adam, peers=self._peers)
        if self.world > 1 and self._peers is None:
            import torch.distributed as dist

            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return self.loss, pred

    def step_selection(self, graph_set, ids=None, global_size: int | None = None, prepared=None):
        """One train step on graphs ``ids`` of a :class:`ResidentGraphSet`: the mini-batch is a list of ids (the only host->device
        traffic of the step), no collate, no copy of the graphs.  Returns (loss, pred, slot_ids): ``pred[s]`` belongs to graph
        ``slot_ids[s]``.  ``prepared = graph_set.select(ids)`` may be computed ahead (e.g. while the previous step runs)."""
        selection, slot_ids = prepared if prepared is not None else graph_set.select(ids)
        self._refresh_adam()
        if self._adam is not None:
            loss, pred = self.forward_backward(graph_set.batch, global_size, adam=self._adam, selection=selection)
        else:
            loss, pred = self.forward_backward(graph_set.batch, global_size, selection=selection)
            self.optimizer.step()
        return loss, pred, slot_ids

    def empty_step(self):
        """A rank without graphs in this (ragged) global mini-batch still joins the gradient exchange and steps the optimizer."""
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v
        if self._peers is not None:
            # zero local contributions through the same finalize kernel: the peers wait for this rank's flags
            m, lib, dev = self.model, _lib.load(), self.flat_grad.device
            out_dim = int(m.fc2.weight.shape[0])
            fi = int(m.conv1.fc.weight.shape[1])
            pred = torch.empty((0, out_dim), dtype=torch.float32, device=dev)
            if self._adam is not None:
                self._adam.lr = self.optimizer.param_groups[0]["lr"]
            with torch.cuda.device(dev):
                ws = workspace(1 << 20, dev)
                rc = lib.drk_ginet_step(
                    None, fi, fi, None, 0, 0, None, None, None, 0, 0, 0, 0,
                    _p(m.conv1.fc.weight), _p(m.conv1_ext.fc.weight), _p(m.conv2.fc.weight), _p(m.conv2_ext.fc.weight),
                    _p(m.fc1.weight), _p(m.fc1.bias), _p(m.fc2.weight), _p(m.fc2.bias), out_dim,
                    int(self.kind), _p(self.loss), 0.0, 0.0, int(self.seed) & (2**64 - 1), _p(self.state), 1,
                    _p(pred) or _p(self.loss), _p(self.loss), *[_p(g) for g in self.grads],
                    ctypes.byref(self._adam) if self._adam is not None else None, ctypes.byref(self._peers), None, _p(ws), ws.numel(), stream_ptr(),
                )
            _lib.check(rc, "drk_ginet_step (empty rank)")
            if self._adam is None:
                self.optimizer.step()
            return
        self.flat_grad.zero_()
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.optimizer.step()

    def _refresh_adam(self):
        """``optimizer.load_state_dict`` replaces the state tensors: re-point the descriptor when that happened."""
        if self._adam is None:
            return
        st = self.optimizer.state.get(self.params[0])
        if not st or (st["step"].data_ptr(), st["exp_avg"].data_ptr()) != self._adam_key:
            self._adam = self._adam_descriptor()
        if self._adam is not None:
            self._adam.lr = self.optimizer.param_groups[0]["lr"]

    def __call__(self, batch, global_size: int | None = None):
        self._refresh_adam()
        if self._adam is not None:  # Adam applied by the finalize kernel on torch's own optimizer state (2 launches per step)
            return self.forward_backward(batch, global_size, adam=self._adam)
        loss, pred = self.forward_backward(batch, global_size)
        self.optimizer.step()
        return loss, pred


class ResidentGraphSet:
    """A whole dataset collated ONCE and kept in HBM (180 GB hold ~1.8 M residue-level graphs): node features, contacts, targets and
    per-graph offsets of all graphs as one packed batch.  A mini-batch is then a list of graph ids -- ``select`` turns it into the
    descriptor the step kernel takes (ids in longest-processing-time-first order, results by slot); nothing else crosses PCIe and
    nothing is gathered or copied on the device: every CTA reads its graph in place.

    This is the device-side replacement of the per-batch collate + transfer of ``Trainer._epoch`` (``trainer.py:682-686``;
    SURVEY 8f rank 2): the reference re-opens the HDF5 file and re-collates every graph in every epoch."""

    def __init__(self, graphs, device):
        from .data import Batch

        host = Batch.from_data_list(list(graphs))
        self.num_graphs = len(host.ptr) - 1
        self.entry_names = list(host.entry_names) if isinstance(getattr(host, "entry_names", None), list) else None
        node_ptr, edge_ptr = host._node_ptr32.tolist(), host._edge_ptr32.tolist()
        self.work = [(edge_ptr[g + 1] - edge_ptr[g]) + 8 * (node_ptr[g + 1] - node_ptr[g]) for g in range(self.num_graphs)]
        import numpy as np

        self.work_np = np.asarray(self.work, dtype=np.int64)
        # per-graph extent of every collated tensor (what `collate` needs to cut a mini-batch out of the packed set)
        from .data import _takes_node_offset

        graphs = list(graphs)
        self._extents = {}
        for k, v in host.__dict__.items():
            if k.startswith("_") or k in ("batch", "ptr") or not isinstance(v, torch.Tensor):
                continue
            dim = 1 if _takes_node_offset(k) else 0
            sizes = np.fromiter(((1 if g.__dict__[k].dim() == 0 else g.__dict__[k].shape[dim]) for g in graphs), dtype=np.int64, count=len(graphs))
            starts = np.zeros(len(graphs) + 1, dtype=np.int64)
            np.cumsum(sizes, out=starts[1:])
            self._extents[k] = (dim, sizes, starts)
        self._lists = {k: v for k, v in host.__dict__.items() if not k.startswith("_") and isinstance(v, list)}
        from .data import pool_sizes

        self._pool_sizes = [pool_sizes(g) for g in graphs]  # per-graph shapes of the community-pooling chain (None: no clustering)
        self.batch = host.to(torch.device(device))
        self._info = None

    @property
    def info(self) -> BlockInfo:
        """Descriptor of the packed set for the per-graph step kernel (built on first use; needs a CUDA device)."""
        if self._info is None:
            self._info = block_info(self.batch)
        return self._info

    def collate(self, ids):
        """The mini-batch of graphs ``ids`` as a device ``Batch`` -- what ``Batch.from_data_list([dataset.get(i) for i in ids]).to(device)``
        returns, bit for bit, but cut out of the packed resident set by a handful of device gathers: no per-graph Python, no
        host tensors, no PCIe traffic beyond one small block of offsets (the reference re-collates on the host every epoch,
        ``trainer.py:541-557``).  Used by the Trainer for every network the per-graph step kernel does not cover."""
        import numpy as np

        from .data import Batch, _takes_node_offset

        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size == 0:
            raise ValueError("empty selection")
        if int(ids.min()) < 0 or int(ids.max()) >= self.num_graphs:
            raise IndexError(f"graph ids must be in [0, {self.num_graphs})")
        dev = self.batch.x.device
        b = int(ids.size)
        # tensors with the same per-graph extents (node-aligned, edge-aligned, one row per graph, ...) share one gather index
        plans, plan_of, seen = [], {}, {}
        for k, (_dim, sizes, starts) in self._extents.items():
            sel, old = sizes[ids], starts[ids]
            key = (sel.tobytes(), old.tobytes())
            if key not in seen:
                new = np.zeros(b + 1, dtype=np.int64)
                np.cumsum(sel, out=new[1:])
                seen[key] = len(plans)
                plans.append((sel, old, new))
            plan_of[k] = seen[key]
        # every small host array in ONE pinned block: ids | per plan: sizes, old starts, new starts
        staging = torch.from_numpy(np.concatenate([ids] + [a for plan in plans for a in plan]))
        block = (staging.pin_memory() if dev.type == "cuda" else staging).to(dev, non_blocking=True)
        arange_b = torch.arange(b, device=dev)
        gathers, off = [], b
        for sel, _old, new in plans:
            sizes_dev, old_dev, new_dev = block[off : off + b], block[off + b : off + 2 * b], block[off + 2 * b : off + 3 * b + 1]
            off += 3 * b + 1
            total = int(new[-1])
            if bool((sel == 1).all()):  # one row per graph: the old starts are the gather index
                gathers.append((old_dev, arange_b, old_dev, new_dev))
                continue
            owner = torch.repeat_interleave(arange_b, sizes_dev, output_size=total)
            src = torch.arange(total, device=dev) + (old_dev - new_dev[:-1])[owner]
            gathers.append((src, owner, old_dev, new_dev))
        node = gathers[plan_of["x"]]
        node_shift = node[3][:-1] - node[2]  # new first node - old first node, per graph
        out = Batch()
        packed = self.batch.__dict__
        for k, pi in plan_of.items():
            src, owner = gathers[pi][0], gathers[pi][1]
            out.__dict__[k] = packed[k].index_select(1, src) + node_shift[owner] if _takes_node_offset(k) else packed[k].index_select(0, src)
        for k, column in self._lists.items():
            out.__dict__[k] = [column[i] for i in ids.tolist()]
        out.batch = node[1]
        out.ptr = node[3]
        if "edge_index" in plan_of:
            n_plan, e_plan = plans[plan_of["x"]], plans[plan_of["edge_index"]]
            if n_plan[2][-1] < 2**31 and e_plan[2][-1] < 2**31:
                out.__dict__["_node_ptr32"] = node[3].to(torch.int32)
                out.__dict__["_edge_ptr32"] = gathers[plan_of["edge_index"]][3].to(torch.int32)
            out.__dict__[Batch._META_KEY] = {"num_graphs": b, "max_graph_nodes": int(n_plan[0].max()), "max_graph_edges": int(e_plan[0].max()),
                                             "num_edges_total": int(e_plan[2][-1])}
        out._attach_pool_sizes([self._pool_sizes[i] for i in ids.tolist()])
        for k in ("_pool_cptr", "_pool_kkptr", "_pool_eptr32", "_pool_c1ptr"):
            if k in out.__dict__:
                out.__dict__[k] = out.__dict__[k].to(dev, non_blocking=True)
        return out

    def select(self, ids):
        """(descriptor, slot_ids) for the graphs ``ids``: ``slot_ids`` is the order in which results come back."""
        from .data import snake_order

        import numpy as np

        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size == 0:
            raise ValueError("empty selection")
        if int(ids.min()) < 0 or int(ids.max()) >= self.num_graphs:
            raise IndexError(f"graph ids must be in [0, {self.num_graphs})")
        order = snake_order(self.work_np[ids]).numpy()
        slot_ids = ids[order].astype(np.int32)
        dev = self.batch.x.device
        # the ids travel through a small ring of pinned staging buffers (no allocation per step); a slot is reused only after
        # the copy that read it has completed (event recorded behind the copy)
        ring = self.__dict__.setdefault("_ring", [])
        if not ring and dev.type == "cuda":
            for _ in range(4):
                ring.append([torch.empty(4096, dtype=torch.int32).pin_memory(), None])
        if dev.type == "cuda" and len(slot_ids) <= 4096:
            slot = ring[self.__dict__.get("_ring_pos", 0) % len(ring)]
            self.__dict__["_ring_pos"] = self.__dict__.get("_ring_pos", 0) + 1
            if slot[1] is not None:
                slot[1].synchronize()
            host_ids = slot[0][: len(slot_ids)]
            host_ids.numpy()[:] = slot_ids
            dev_ids = host_ids.to(dev, non_blocking=True)
            slot[1] = torch.cuda.Event()
            slot[1].record(torch.cuda.current_stream(dev))
        else:
            dev_ids = torch.from_numpy(slot_ids).to(dev)
        sel = BlockInfo()
        base = self.info
        sel.node_ptr, sel.edge_ptr, sel.edges, sel.layout = base.node_ptr, base.edge_ptr, base.edges, base.layout
        sel.max_nodes, sel.max_edges, sel.status = base.max_nodes, base.max_edges, base.status
        sel.order = dev_ids
        sel.num_graphs = len(slot_ids)
        sel.by_slot = True
        return sel, slot_ids.tolist()


    def select_epoch(self, id_lists):
        """``select`` for every mini-batch of an epoch at once: the slot-ordered ids of ALL batches travel in one host->device copy and
        every selection's ``order`` is a view of that one buffer.  Returns ``(all_ids_dev, [(descriptor, slot_ids), ...])`` with
        ``descriptor.order = all_ids_dev[offset : offset + len]``."""
        from .data import snake_order

        import numpy as np

        chunks = []
        for ids in id_lists:
            ids = np.asarray(ids, dtype=np.int64).reshape(-1)
            if ids.size == 0:
                raise ValueError("empty selection")
            if int(ids.min()) < 0 or int(ids.max()) >= self.num_graphs:
                raise IndexError(f"graph ids must be in [0, {self.num_graphs})")
            chunks.append(ids[snake_order(self.work_np[ids]).numpy()].astype(np.int32))
        dev = self.batch.x.device
        flat = torch.from_numpy(np.concatenate(chunks)) if chunks else torch.zeros(0, dtype=torch.int32)
        all_ids = (flat.pin_memory() if dev.type == "cuda" else flat).to(dev, non_blocking=True)
        base = self.info
        out, off = [], 0
        for slot_ids in chunks:
            sel = BlockInfo()
            sel.node_ptr, sel.edge_ptr, sel.edges, sel.layout = base.node_ptr, base.edge_ptr, base.edges, base.layout
            sel.max_nodes, sel.max_edges, sel.status = base.max_nodes, base.max_edges, base.status
            sel.order = all_ids[off : off + len(slot_ids)]
            sel.num_graphs = len(slot_ids)
            sel.by_slot = True
            out.append((sel, slot_ids.tolist()))
            off += len(slot_ids)
        return all_ids, out


class CapturedSelectionStep:
    """The two-launch train step on a :class:`ResidentGraphSet`, captured ONCE into a CUDA graph for a fixed batch size: a step is
    then "write the graph ids into a static device buffer, replay".  The host's share of a step drops to the LPT ordering of the ids
    (numpy) and one small copy; everything else -- index build, forward, loss, backward, gradient exchange, Adam -- replays.

    Capturing does not disturb training: parameters and optimizer state are restored after the warm-up run."""

    def __init__(self, step: "GINetFusedStep", graph_set: "ResidentGraphSet", batch_size: int, global_size: int | None = None):
        import numpy as np

        if batch_size < 1 or batch_size > 4096:
            raise ValueError("batch_size must be in [1, 4096]")
        self.step, self.graph_set, self.batch_size, self.global_size = step, graph_set, int(batch_size), global_size
        dev = graph_set.batch.x.device
        self.ids_dev = torch.zeros(self.batch_size, dtype=torch.int32, device=dev)
        base = graph_set.info
        sel = BlockInfo()
        sel.node_ptr, sel.edge_ptr, sel.edges, sel.layout = base.node_ptr, base.edge_ptr, base.edges, base.layout
        sel.max_nodes, sel.max_edges, sel.status = base.max_nodes, base.max_edges, base.status
        sel.order, sel.num_graphs, sel.by_slot = self.ids_dev, self.batch_size, True
        self.selection = sel
        self._ring = [[torch.empty(self.batch_size, dtype=torch.int32).pin_memory(), None] for _ in range(4)]
        self._pos = 0
        self._np = np
        # warm-up + capture on a side stream, then put parameters and optimizer state back
        params = [p.detach().clone() for p in step.params]
        opt_state = {k: {n: (v.clone() if isinstance(v, torch.Tensor) else v) for n, v in st.items()} for k, st in step.optimizer.state.items()}
        state = step.state.clone()
        self._write_ids(np.arange(self.batch_size) % graph_set.num_graphs)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.pred = self._run()
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for p, saved in zip(step.params, params):
                p.copy_(saved)
            for k, st in step.optimizer.state.items():
                for n, v in st.items():
                    if isinstance(v, torch.Tensor) and k in opt_state and n in opt_state[k]:
                        v.copy_(opt_state[k][n])
            step.state[0:1].copy_(state[0:1])  # the dropout step counter; the exchange epoch ([2]) must keep counting: the peers' flags did

    def _run(self):
        st = self.step
        st._refresh_adam()
        if st._adam is not None:
            return st.forward_backward(self.graph_set.batch, self.global_size, adam=st._adam, selection=self.selection)
        out = st.forward_backward(self.graph_set.batch, self.global_size, selection=self.selection)
        st.optimizer.step()
        return out

    def _write_ids(self, slot_ids):
        slot = self._ring[self._pos % len(self._ring)]
        self._pos += 1
        if slot[1] is not None:
            slot[1].synchronize()
        slot[0].numpy()[:] = slot_ids
        self.ids_dev.copy_(slot[0], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(self.ids_dev.device))

    def __call__(self, ids):
        """One train step on exactly ``batch_size`` graph ids.  Returns (loss, pred, slot_ids) -- views of static buffers, valid until
        the next call."""
        from .data import snake_order

        np = self._np
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size != self.batch_size:
            raise ValueError(f"this step was captured for {self.batch_size} graphs, got {ids.size}")
        slot_ids = ids[snake_order(self.graph_set.work_np[ids]).numpy()]
        self._write_ids(slot_ids)
        self.graph.replay()
        return self.loss, self.pred, slot_ids
