"""CPU: host-side logic added for the per-graph kernels -- collate extras (int32 offsets, snake schedule, contacts kept once for
doubled graphs), the lazy / partial device transfer of a batch, the prefetching loader's bookkeeping."""
from __future__ import annotations

import torch

from deeprank2_b200.data import Batch, Data, snake_order
from deeprank2_b200.pipeline import batch_nbytes, shallow_host_view
from deeprank2_b200.synthetic import make_batch, make_graph


def test_collate_offsets_and_pairs_describe_the_same_edges():
    graphs = [make_graph(g) for g in range(5)]
    batch = Batch.from_data_list(graphs)
    node_ptr, edge_ptr, pair_ptr = batch._node_ptr32, batch._edge_ptr32, batch._pair_ptr32
    assert node_ptr.dtype == edge_ptr.dtype == pair_ptr.dtype == torch.int32
    assert torch.equal(node_ptr.long(), batch.ptr)
    assert int(edge_ptr[-1]) == batch.num_edges and torch.equal(edge_ptr, 2 * pair_ptr)
    for g in range(5):
        e0, e1 = int(edge_ptr[g]), int(edge_ptr[g + 1])
        p0, p1 = int(pair_ptr[g]), int(pair_ptr[g + 1])
        ei = batch.edge_index[:, e0:e1]
        pairs = batch._pairs[:, p0:p1]
        half = (e1 - e0) // 2
        assert torch.equal(ei[:, :half], pairs) and torch.equal(ei[:, half:], pairs.flip(0))  # dataset.py:944-948 layout
        words = batch._pairs16[p0:p1].long() & 0xFFFFFFFF  # the same contacts as one packed word of graph-local ids
        lo_node = int(node_ptr[g])
        assert torch.equal((words & 0xFFFF) + lo_node, pairs[0]) and torch.equal((words >> 16) + lo_node, pairs[1])
        lo, hi = int(node_ptr[g]), int(node_ptr[g + 1])
        assert int(ei.min()) >= lo and int(ei.max()) < hi  # edges of a graph stay inside it and are contiguous
    meta = batch.__dict__[Batch._META_KEY]
    assert meta["num_edges_total"] == batch.num_edges and meta["max_graph_edges"] == int((edge_ptr[1:] - edge_ptr[:-1]).max())


def test_pairs_are_dropped_when_a_graph_is_not_doubled():
    doubled = make_graph(0, 4)
    d = Data(x=torch.zeros(3, 4), edge_index=torch.tensor([[0, 1, 2, 0], [1, 2, 0, 2]]), edge_attr=torch.zeros(4, 1), y=torch.zeros(1),
             pos=torch.zeros(3, 3))
    d.entry_names = "hand-made"
    mixed = Batch.from_data_list([doubled, d])
    assert "_pairs" not in mixed.__dict__ and "_pairs16" not in mixed.__dict__ and "_edge_ptr32" in mixed.__dict__


def test_snake_order_balances_rounds():
    work = [100 - i for i in range(10)]  # already descending
    order = snake_order(work, ctas=4).tolist()
    assert sorted(order) == list(range(10))
    assert order[:4] == [0, 1, 2, 3] and order[4:8] == [7, 6, 5, 4] and order[8:] == [8, 9]
    totals = [sum(work[order[s]] for s in range(b, 10, 4)) for b in range(4)]
    assert max(totals[:2]) - min(totals[:2]) <= 2  # CTAs with the same number of graphs carry (almost) the same work
    # at most two rounds: closed form of the same schedule -- the CTAs that run two graphs come first and pair a mid-sized graph with
    # one of the smallest; deterministic (ties keep graph order)
    two = snake_order([9, 8, 7, 6, 2, 1], ctas=4).tolist()
    assert sorted(two) == list(range(6))
    loads = [sum([9, 8, 7, 6, 2, 1][two[s]] for s in range(b, 6, 4)) for b in range(4)]
    assert sorted(loads) == [8, 8, 8, 9]  # 7+1 and 6+2 paired, 9 and 8 alone
    assert snake_order([5, 5, 5], ctas=2).tolist() == snake_order([5, 5, 5], ctas=2).tolist()
    assert sorted(snake_order([3, 9, 1, 7], ctas=3).tolist()) == [0, 1, 2, 3]
    assert snake_order([4, 2, 6], ctas=148).tolist() == [2, 0, 1]  # fewer graphs than CTAs: largest first


def test_partial_transfer_keeps_deferred_tensors_reachable():
    host = make_batch(3)
    view = shallow_host_view(host)
    only = ("x", "_pairs", "y")
    # a CPU "device" moves everything (the deferral only applies to CUDA targets) ...
    moved = view.to("cpu", only=only)
    assert "_pending" not in moved.__dict__ and moved.edge_attr.shape == host.edge_attr.shape
    # ... so exercise the deferred bookkeeping directly
    lazy = shallow_host_view(host)
    pending = {k: lazy.__dict__.pop(k) for k in ("edge_attr", "pos", "batch")}
    lazy.__dict__["_pending"] = (torch.device("cpu"), pending)
    assert set(lazy.keys) >= {"x", "edge_index", "edge_attr", "pos", "batch", "y"}  # still listed
    assert torch.equal(lazy.edge_attr, host.edge_attr)  # first access materialises it ...
    assert "edge_attr" in lazy.__dict__ and "edge_attr" not in lazy.__dict__["_pending"][1]  # ... once
    clone = lazy.clone()
    assert torch.equal(clone.pos, host.pos) and torch.equal(lazy.pos, host.pos)
    try:
        lazy.no_such_attribute
    except AttributeError:
        pass
    else:
        raise AssertionError("unknown attributes must still raise AttributeError")
    assert host.__dict__.get("_pending") is None  # the cached host batch is never modified


def test_batch_nbytes_counts_what_a_partial_copy_moves():
    host = make_batch(2)
    full = batch_nbytes(host)
    from deeprank2_b200.fused import GINetFusedStep

    step_fields = GINetFusedStep.FIELDS
    assert "_pairs16" in step_fields and "_pairs" not in step_fields and "edge_index" not in step_fields
    part = batch_nbytes(host, step_fields)
    expected = sum(host.__dict__[k].numel() * host.__dict__[k].element_size() for k in step_fields)
    assert part == expected < full
    assert host._pairs.numel() * 2 == host.edge_index.numel() and host._pairs16.numel() * 2 == host._pairs.numel()


def test_resident_set_collate_is_the_host_collate():
    """``ResidentGraphSet.collate`` (gathers out of the packed set; runs on the GPU in production) against ``Batch.from_data_list``
    on the same ids: every tensor, dtype, list attribute and the collate metadata -- here on CPU tensors, which exercises the
    same index arithmetic."""
    from deeprank2_b200.fused import ResidentGraphSet

    graphs = [make_graph(g, 6, 2, with_clusters=True, n=7 + 3 * g) for g in range(9)]
    gset = ResidentGraphSet(graphs, "cpu")
    for ids in ([4, 0, 8, 8, 1], [2], list(range(9)), [8, 7, 6]):
        got = gset.collate(ids)
        ref = Batch.from_data_list([graphs[i] for i in ids])
        for k in ("x", "edge_index", "edge_attr", "y", "pos", "cluster0", "cluster1", "batch", "ptr", "_node_ptr32", "_edge_ptr32"):
            a, b = got.__dict__[k], ref.__dict__[k]
            assert a.dtype == b.dtype and torch.equal(a, b), k
        assert got.entry_names == ref.entry_names
        assert got.__dict__[Batch._META_KEY] == ref.__dict__[Batch._META_KEY]
        assert sorted(got.keys) == sorted(ref.keys)
    try:
        gset.collate([9])
    except IndexError:
        pass
    else:
        raise AssertionError("an id outside the set must raise IndexError")
    try:
        gset.collate([])
    except ValueError:
        pass
    else:
        raise AssertionError("an empty selection must raise ValueError")


def test_resident_batches_shard_like_the_streamed_loader():
    """``ResidentBatches(collate=True)`` hands every rank the same slice of every global mini-batch as ``BatchLoader`` does
    (``parallel.shard_indices``), ragged tail included; the union over ranks is the global batch, in order."""
    from deeprank2_b200.fused import ResidentGraphSet
    from deeprank2_b200.trainer import BatchLoader, ResidentBatches

    class _ListDataset:
        def __init__(self, graphs):
            self.graphs = graphs

        def __len__(self):
            return len(self.graphs)

        def get(self, i):
            return self.graphs[i]

    graphs = [make_graph(g, 5, 1, n=6 + g) for g in range(11)]
    gset = ResidentGraphSet(graphs, "cpu")
    world = 2
    per_rank = []
    for rank in range(world):
        resident = list(ResidentBatches(gset, batch_size=4, shuffle=True, rank=rank, world_size=world, collate=True))
        streamed = list(BatchLoader(_ListDataset(graphs), batch_size=4, shuffle=True, device=None, rank=rank, world_size=world))
        assert len(resident) == len(streamed) == 3
        for (rb, rg), (sb, sg) in zip(resident, streamed):
            assert rg == sg
            assert (rb is None) == (sb is None)
            if rb is not None:
                assert rb.entry_names == sb.entry_names
                assert torch.equal(rb.x, sb.x) and torch.equal(rb.edge_index, sb.edge_index) and torch.equal(rb.batch, sb.batch)
        per_rank.append(resident)
    # union over ranks == the global mini-batches of the shared permutation
    seen = [name for step in range(3) for rank in range(world) if per_rank[rank][step][0] is not None for name in per_rank[rank][step][0].entry_names]
    assert sorted(seen) == sorted(g.entry_names for g in graphs)
    assert [per_rank[0][s][1] for s in range(3)] == [4, 4, 3]


def test_threaded_collate_keeps_order_and_content():
    """The streamed loader's collate threads hand the batches out in submission order: identical to collating in the calling thread."""
    from deeprank2_b200.trainer import BatchLoader

    class _ListDataset:
        def __init__(self, graphs):
            self.graphs = graphs

        def __len__(self):
            return len(self.graphs)

        def get(self, i):
            return self.graphs[i]

    ds = _ListDataset([make_graph(g, 5, 1, n=5 + (g * 7) % 11) for g in range(23)])
    runs = []
    for workers in (0, 3, None):
        loader = BatchLoader(ds, batch_size=4, shuffle=True, device=None, seed=5, num_workers=workers)
        runs.append([(b.entry_names, b.x.clone(), b.edge_index.clone(), gs) for b, gs in loader])
    assert len(runs[0]) == 6
    for other in runs[1:]:
        assert len(other) == len(runs[0])
        for (n0, x0, e0, g0), (n1, x1, e1, g1) in zip(runs[0], other):
            assert n0 == n1 and g0 == g1 and torch.equal(x0, x1) and torch.equal(e0, e1)
    # an abandoned iteration shuts the pool down cleanly
    it = iter(BatchLoader(ds, batch_size=2, shuffle=False, device=None, num_workers=2))
    next(it)
    it.close()
