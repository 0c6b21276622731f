"""GPU: community pooling chain and the torch_scatter-compatible functions -- integer outputs
bit-exact against the golden vectors recorded from the reference, floats at the fp32 bar."""
from __future__ import annotations

import pytest
import torch

from conftest import CLUSTERED_CASES, assert_close, assert_equal_int, load_golden
from oracle import thirdparty as tp

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _batch(g):
    from deeprank2_b200.data import Batch

    b = Batch()
    for k, v in vars(g.inputs()).items():
        setattr(b, k, v.clone())
    return b.to(DEV)


@pytest.mark.parametrize("case", CLUSTERED_CASES)
def test_pooling_chain_vs_reference(case):
    from deeprank2_b200.utils.community_pooling import community_pooling, get_preloaded_cluster, max_pool_x

    g = load_golden(case)
    b = _batch(g)
    ng = int(b.ptr.numel()) - 1
    c0 = get_preloaded_cluster(b.cluster0, b.batch, ng)
    assert c0.data_ptr() == b.cluster0.data_ptr(), "in place, like the reference"
    assert_equal_int(c0, g.t("pooling/out/cluster0_global"), "cluster0 offsets")
    pooled = community_pooling(c0, b)
    assert_equal_int(pooled.edge_index, g.t("pooling/out/pool_edge_index"), "pooled edge_index")
    assert_equal_int(pooled.batch, g.t("pooling/out/pool_batch"), "pooled batch")
    assert torch.equal(pooled.x.cpu(), g.t("pooling/out/pool_x")), "segment max is exact"
    assert_close(pooled.edge_attr, g.t("pooling/out/pool_edge_attr"), "pooled edge_attr")
    assert_close(pooled.pos, g.t("pooling/out/pool_pos"), "pooled pos")
    c1 = get_preloaded_cluster(pooled.cluster1, pooled.batch, ng)
    assert_equal_int(c1, g.t("pooling/out/cluster1_global"), "cluster1 offsets")
    x2, b2 = max_pool_x(c1, pooled.x, pooled.batch)
    assert torch.equal(x2.cpu(), g.t("pooling/out/pool2_x"))
    assert_equal_int(b2, g.t("pooling/out/pool2_batch"), "pool2 batch")


def test_community_pooling_docstring_example():
    """the 2 x 6-node example of the reference docstring (community_pooling.py:181-191) with fixed clusters"""
    from deeprank2_b200.data import Batch, Data
    from deeprank2_b200.utils.community_pooling import community_pooling

    edge_index = torch.tensor([[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]], dtype=torch.long)
    x = torch.tensor([[0.0], [1.0], [2.0], [3.0], [4.0], [5.0]])
    pos = torch.arange(18, dtype=torch.float32).reshape(6, 3)
    d = Data(x=x, edge_index=edge_index, edge_attr=torch.ones(8, 1), pos=pos)
    batch = Batch.from_data_list([d, d]).to(DEV)
    cluster = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3], device=DEV)
    out = community_pooling(cluster, batch)
    assert out.x.flatten().tolist() == [2.0, 5.0, 2.0, 5.0]
    assert out.edge_index.numel() == 0  # the two triangles are disconnected components: only self loops remain
    assert out.batch.tolist() == [0, 0, 1, 1]
    # reference semantics on the same input
    ref = tp.Batch.from_data_list([tp.Data(x=x, edge_index=edge_index, edge_attr=torch.ones(8, 1), pos=pos)] * 2)
    rc, rperm = tp.consecutive_cluster(cluster.cpu())
    assert_close(out.pos, tp.scatter_mean(ref.pos, rc, dim=0), "pos")


@pytest.mark.parametrize("width", [1, 3, 16, 50])
def test_scatter_functions_match_torch_scatter_semantics(width):
    from deeprank2_b200 import ops

    gen = torch.Generator().manual_seed(width)
    n, segs = 777, 41
    index = torch.randint(0, segs, (n,), generator=gen)
    index[index == 7] = 8  # an empty segment
    src = torch.randn(n, width, generator=gen)
    s_dev = src.to(DEV).requires_grad_(True)
    s_cpu = src.clone().requires_grad_(True)
    got = ops.scatter_sum(s_dev, index.to(DEV), dim=0, dim_size=segs)
    ref = tp.scatter_sum(s_cpu, index, dim=0, dim_size=segs)
    assert_close(got, ref, "sum")
    gout = torch.randn(segs, width, generator=gen)
    got.backward(gout.to(DEV))
    ref.backward(gout)
    assert_close(s_dev.grad, s_cpu.grad, "sum grad")

    assert_close(ops.scatter_mean(src.to(DEV), index.to(DEV), dim=0, dim_size=segs), tp.scatter_mean(src, index, dim=0, dim_size=segs), "mean")
    init = torch.randn(segs, width, generator=gen)
    got = ops.scatter_mean(src.to(DEV), index.to(DEV), dim=0, out=init.clone().to(DEV))
    assert_close(got, tp.scatter_mean(src, index, dim=0, out=init.clone()), "mean with out=")

    s_dev = src.to(DEV).requires_grad_(True)
    s_cpu = src.clone().requires_grad_(True)
    gmax, garg = ops.scatter_max(s_dev, index.to(DEV), dim=0, dim_size=segs)
    rmax, rarg = tp.scatter_max(s_cpu, index, dim=0, dim_size=segs)
    assert torch.equal(gmax.detach().cpu(), rmax.detach())
    assert_equal_int(garg, rarg, "argmax (first max wins, empty -> n)")
    gmax.backward(gout.to(DEV))
    rmax.backward(gout)
    assert_close(s_dev.grad, s_cpu.grad, "max grad")
    # without dim_size the output has index.max()+1 rows, like torch_scatter
    assert ops.scatter_sum(src.to(DEV), index.to(DEV), dim=0).shape[0] == int(index.max()) + 1


def test_scatter_max_ties_pick_first():
    from deeprank2_b200 import ops

    src = torch.tensor([[1.0], [3.0], [3.0], [2.0], [3.0]], device=DEV)
    index = torch.tensor([0, 0, 0, 1, 0], device=DEV)
    out, arg = ops.scatter_max(src, index, dim=0, dim_size=3)
    assert out.flatten().tolist() == [3.0, 2.0, 0.0]
    assert arg.flatten().tolist() == [1, 3, 5]


def _clustered_batch(n_graphs=12, first=300):
    from deeprank2_b200.synthetic import make_batch

    return make_batch(n_graphs, first=first, with_clusters=True)


def test_pooling_with_collate_sizes_has_no_readback_and_matches_the_reference_semantics():
    """Batches made by ``Batch.from_data_list`` carry the shapes of their pooled tensors (``meta("pool")``): the chain then runs without a
    single host synchronisation (checked with ``torch.cuda.set_sync_debug_mode("error")``) and gives, bit for bit, what PyG's
    consecutive_cluster / pool_edge / max_pool_x give on the CPU (oracle/thirdparty.py)."""
    from deeprank2_b200.utils import community_pooling as cp

    host = _clustered_batch()
    m = host.meta("pool")
    assert m is not None and m["C0"] > 0 and m["E1"] > 0
    b = host.clone().to(DEV)
    ng = host.num_graphs
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        c0 = cp.get_preloaded_cluster(b.cluster0.clone(), b.batch, ng)
        pooled = cp.community_pooling(c0, b)
        c1 = cp.get_preloaded_cluster(pooled.cluster1.clone(), pooled.batch, ng)
        x2, b2 = cp.max_pool_x(c1, pooled.x, pooled.batch, meta=cp.pool_meta(pooled, 1))
    finally:
        torch.cuda.set_sync_debug_mode("default")
    cp.check_status(DEV)
    # reference semantics on the CPU
    rc0 = host.cluster0.clone()
    for g in range(1, ng):
        rc0[host.batch == g] += rc0[host.batch == g - 1].max() + 1
    inv, perm = tp.consecutive_cluster(rc0)
    r_x, _ = tp.scatter_max(host.x, inv, dim=0)
    r_ei, r_ea = tp.pool_edge(inv, host.edge_index, host.edge_attr)
    assert_equal_int(pooled.edge_index, r_ei, "pooled edge_index")
    assert pooled.edge_index.shape[1] == m["E1"] and pooled.x.shape[0] == m["C0"]
    assert torch.equal(pooled.x.cpu(), r_x)
    assert_close(pooled.edge_attr, r_ea, "pooled edge_attr")
    assert_equal_int(pooled.batch, host.batch[perm], "pooled batch")
    assert_close(pooled.pos, tp.scatter_mean(host.pos, inv, dim=0), "pooled pos")
    rc1 = host.cluster1.clone()
    pb = host.batch[perm]
    for g in range(1, ng):
        rc1[pb == g] += rc1[pb == g - 1].max() + 1
    r_x2, r_b2 = tp.max_pool_x(rc1, r_x, pb)
    assert torch.equal(x2.cpu(), r_x2) and x2.shape[0] == m["C1"]
    assert_equal_int(b2, r_b2, "pool2 batch")
    # the pooled batch is a collated batch again: its graph index comes from the per-graph (blocked) builder
    from deeprank2_b200.graph import graph_index

    gi = graph_index(pooled)
    ref_ptr = torch.zeros(m["C0"] + 1, dtype=torch.int64)
    ref_ptr[1:] = torch.cumsum(torch.bincount(r_ei[0], minlength=m["C0"]), 0)
    assert_equal_int(gi.rowptr, ref_ptr, "rowptr of the pooled graph")
    assert_equal_int(gi.colidx, r_ei[1], "pool_edge's output is already destination sorted")


def test_pooling_flags_sizes_that_do_not_match_the_batch():
    from deeprank2_b200.utils import community_pooling as cp

    host = _clustered_batch(4)
    b = host.clone().to(DEV)
    b.meta("pool")["C0"] -= 3  # stale meta: fewer clusters than the data holds
    c0 = cp.get_preloaded_cluster(b.cluster0.clone(), b.batch, host.num_graphs)
    cp.community_pooling(c0, b)
    with pytest.raises(IndexError):
        cp.check_status(DEV)
    cp.check_status(DEV)  # the accumulator was reset


@pytest.mark.parametrize("net_name", ["foutnet", "ginet", "sgat"])
def test_clustered_train_step_replays_from_a_cuda_graph(net_name):
    """No host read-back anywhere in the step of a clustered network: it can be captured once and replayed; two replays on the same
    weights are bit-identical to two eager steps."""
    import copy
    import importlib

    from deeprank2_b200.step import GraphedTrainStep, TrainStep

    cls = {"foutnet": "FoutNet", "ginet": "GINet", "sgat": "SGAT"}[net_name]
    mod = importlib.import_module(f"deeprank2_b200.neuralnets.gnn.{net_name}")
    batch = _clustered_batch(16).to(DEV)
    loss_fn = torch.nn.MSELoss()

    def build():
        torch.manual_seed(3)
        net = getattr(mod, cls)(50, 1, 1).to(DEV).eval()  # eval: no dropout stream to keep in step
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
        inner = TrainStep(net, opt, loss_fn)

        def step(b):
            view = copy.copy(b)
            view.__dict__ = dict(b.__dict__)
            return inner(view)

        return net, step

    net_e, step_e = build()
    losses_e = [float(step_e(batch)[0]) for _ in range(3)]
    net_g, step_g = build()
    graphed = GraphedTrainStep(step_g, batch, warmup=1)  # the warm-up step is the first of the three
    losses_g = [losses_e[0]]
    for _ in range(2):
        loss, _ = graphed.replay()
        losses_g.append(float(loss))
    assert losses_g[1:] == losses_e[1:], (losses_g, losses_e)
    for (k, p), q in zip(net_e.named_parameters(), net_g.parameters()):
        assert torch.equal(p, q), k


@pytest.mark.parametrize("num_segments,density,slack", [(0, 0.5, 3), (1, 1.0, 0), (4096, 0.3, 0), (4097, 0.9, 5), (50_000, 0.05, 0), (370_001, 0.25, 17), (20_000, 0.6, -40)])
def test_compact_segments_multi_chunk_scan_vs_numpy(num_segments, density, slack):
    """``drk_compact_segments`` (two launches over chunks of 4096 segment sizes) against numpy: rank, compact offsets, kept ids, last member,
    count, trailing capacity filled with the total, and the status flag when the capacity is too small (slack < 0)."""
    import numpy as np

    from deeprank2_b200 import _lib
    from deeprank2_b200.graph import stream_ptr, workspace

    lib = _lib.load()
    rng = np.random.default_rng(num_segments + 1)
    sizes = (rng.random(num_segments) < density) * rng.integers(1, 4, size=num_segments)
    ptr = np.zeros(num_segments + 1, dtype=np.int32)
    np.cumsum(sizes, out=ptr[1:])
    total = int(ptr[-1])
    perm = rng.permutation(max(total, 1)).astype(np.int32)
    kept = np.flatnonzero(sizes > 0)
    cap = max(len(kept) + slack, 0)
    dev = torch.device("cuda")
    t_ptr, t_perm = torch.from_numpy(ptr).to(dev), torch.from_numpy(perm).to(dev)
    rank = torch.full((max(num_segments, 1),), -7, dtype=torch.int64, device=dev)
    ptr_out = torch.full((cap + 1,), -7, dtype=torch.int32, device=dev)
    ids = torch.full((max(cap, 1),), -7, dtype=torch.int32, device=dev)
    last = torch.full((max(cap, 1),), -7, dtype=torch.int64, device=dev)
    count = torch.full((1,), -7, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = workspace(lib.drk_compact_segments_workspace_bytes(num_segments), dev)
    rc = lib.drk_compact_segments(t_ptr.data_ptr(), num_segments, t_perm.data_ptr(), rank.data_ptr(), ptr_out.data_ptr(), ids.data_ptr(), last.data_ptr(), cap,
                                  count.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_compact_segments")
    assert int(count.item()) == len(kept)
    assert bool(int(status.item()) & _lib.STATUS_INDEX_RANGE) == (len(kept) > cap)
    want_rank = np.full(num_segments, -1, dtype=np.int64)
    want_rank[kept] = np.arange(len(kept))
    assert np.array_equal(rank.cpu().numpy()[:num_segments], want_rank)
    m = min(len(kept), cap)
    assert np.array_equal(ptr_out.cpu().numpy()[:m], ptr[kept[:m]])
    assert np.all(ptr_out.cpu().numpy()[m:] == total)
    assert np.array_equal(ids.cpu().numpy()[:m], kept[:m].astype(np.int32))
    assert np.array_equal(last.cpu().numpy()[:m], perm[ptr[kept[:m] + 1] - 1].astype(np.int64))


@pytest.mark.parametrize("n_graphs,fe", [(5, 1), (12, 3)])
def test_per_graph_pool_edge_equals_the_global_route(n_graphs, fe, monkeypatch):
    """``drk_pool_edge_blocked`` (one CTA per graph, shared memory) against the global chain (dense pair ids -> counting sort ->
    compaction -> decode -> segmented sum): same pooled edge_index bit for bit, same merged attributes (both add in ascending edge
    id; the global route's segmented sum splits long lists over lanes, hence a tolerance), and against PyG's pool_edge on the CPU."""
    from deeprank2_b200 import _lib
    from deeprank2_b200.synthetic import make_batch
    from deeprank2_b200.utils import community_pooling as cp

    host = make_batch(n_graphs, first=40, n_edge_features=fe, with_clusters=True)
    outs = []
    for blocked in (True, False):
        monkeypatch.setattr(cp, "POOL_BLOCKED", blocked)
        b = host.clone().to(DEV)
        assert (cp._pool_blocks(b) is not None)
        c0 = cp.get_preloaded_cluster(b.cluster0.clone(), b.batch, host.num_graphs)
        before = _lib.launch_count()
        pooled = cp.community_pooling(c0, b)
        outs.append((pooled.edge_index.clone(), pooled.edge_attr.clone(), pooled.x.clone(), _lib.launch_count() - before))
        cp.check_status(DEV)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][2], outs[1][2])
    assert_close(outs[0][1], outs[1][1], "merged attributes, per-graph kernel vs global route")
    assert outs[0][3] < outs[1][3] - 8, "the per-graph kernel replaces the ten launches of the global chain"
    rc0 = host.cluster0.clone()
    for g in range(1, host.num_graphs):
        rc0[host.batch == g] += rc0[host.batch == g - 1].max() + 1
    inv, _ = tp.consecutive_cluster(rc0)
    r_ei, r_ea = tp.pool_edge(inv, host.edge_index, host.edge_attr)
    assert_equal_int(outs[0][0], r_ei, "pooled edge_index")
    assert_close(outs[0][1], r_ea, "pooled edge_attr")


def test_per_graph_consecutive_cluster_equals_the_global_route(monkeypatch):
    """``drk_consecutive_blocked`` (one CTA per graph) against the global chain (counting sort + compaction + rank gather) on both pooling
    levels of a collated batch: relabelling, PyG's ``perm`` (last member), the segment plan -- bit for bit -- and fewer launches."""
    from deeprank2_b200 import _lib
    from deeprank2_b200.synthetic import make_batch
    from deeprank2_b200.utils import community_pooling as cp

    host = make_batch(9, first=70, with_clusters=True)
    outs = []
    for blocked in (True, False):
        monkeypatch.setattr(cp, "POOL_BLOCKED", blocked)
        b = host.clone().to(DEV)
        m0 = cp.pool_meta(b, 0)
        assert "blocks" in m0
        c0 = cp.get_preloaded_cluster(b.cluster0.clone(), b.batch, host.num_graphs)
        before = _lib.launch_count()
        st0 = cp._consecutive(c0, m0)
        n0 = _lib.launch_count() - before
        pooled = cp.community_pooling(c0, b)
        m1 = cp.pool_meta(pooled, 1)
        assert "blocks" in m1
        c1 = cp.get_preloaded_cluster(pooled.cluster1.clone(), pooled.batch, host.num_graphs)
        st1 = cp._consecutive(c1, m1)
        x2, b2 = cp.max_pool_x(c1, pooled.x, pooled.batch, meta=m1)
        cp.check_status(DEV)
        outs.append((n0, [t.clone() for st in (st0, st1) for t in (st.inv, st.last, st.plan.ptr, st.plan.perm)], x2.clone(), b2.clone()))
    assert outs[0][0] == 1 and outs[1][0] >= 7
    for a, b_ in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b_)
    assert torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][3], outs[1][3])
