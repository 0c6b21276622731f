"""GPU parity tests of the individual C-ABI kernels against the CPU oracle / plain torch fp32.

Run on the B200 box: ``python -m pytest tests -m gpu``.  Integer structures are compared
bit-exactly, floating point at rtol 1e-5 / atol 1e-5*max|ref| (conftest.assert_close).
"""
from __future__ import annotations

import pytest
import torch

from conftest import GOLDEN_CASES, assert_close, assert_equal_int, load_golden
from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from deeprank2_b200 import ops

    return ops


def _graph(edge_index, n, batch=None, num_graphs=None):
    from deeprank2_b200.graph import GraphIndex

    return GraphIndex.build(edge_index.to(DEV), n, batch=None if batch is None else batch.to(DEV), num_graphs=num_graphs)


# ------------------------------------------------------------------ index structures (bit exact)
def _check_index(edge_index, n, batch=None, num_graphs=None):
    g = _graph(edge_index, n, batch, num_graphs)
    g.check()
    rowptr, colidx, perm = R.graph_csr(edge_index, n)
    colptr, rowidx, permT = R.graph_csc(edge_index, n)
    assert_equal_int(g.rowptr, rowptr, "rowptr")
    assert_equal_int(g.perm, perm, "perm")
    assert_equal_int(g.colidx, colidx, "colidx")
    assert_equal_int(g.colptr, colptr, "colptr")
    assert_equal_int(g.permT, permT, "permT")
    assert_equal_int(g.rowidx, rowidx, "rowidx")
    assert g.rowptr.dtype == torch.int32
    if batch is not None:
        assert_equal_int(g.graph_ptr, R.batch_offsets(batch, num_graphs), "graph_ptr")
        assert_equal_int(g.batch32, batch, "batch32")
    return g


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_graph_index_golden(case):
    d = load_golden(case).inputs()
    _check_index(d.edge_index, d.x.shape[0], d.batch, int(d.ptr.numel()) - 1)


def test_graph_index_edge_cases():
    empty = torch.zeros(2, 0, dtype=torch.long)
    _check_index(empty, 5, torch.tensor([0, 0, 1, 1, 1]), 2)
    _check_index(empty, 1)
    _check_index(torch.tensor([[0], [0]]), 1)
    # one hub node with 700 incoming edges (exercises the >32 rank-sort path) and duplicates
    gen = torch.Generator().manual_seed(0)
    src = torch.randint(0, 50, (700,), generator=gen)
    ei = torch.stack([torch.full((700,), 3), src])
    _check_index(torch.cat([ei, ei.flip(0)], dim=1), 50)
    # empty graphs inside the batch vector: graph 1 and 3 have no nodes
    _check_index(torch.tensor([[0, 2], [1, 3]]), 4, torch.tensor([0, 0, 2, 2]), 5)


def test_graph_index_full_size_synthetic():
    from deeprank2_b200.synthetic import make_batch

    b = make_batch(256)
    _check_index(b.edge_index, b.num_nodes, b.batch, 256)


def test_graph_index_random_unstructured():
    gen = torch.Generator().manual_seed(5)
    n, e = 10_000, 300_000
    ei = torch.randint(0, n, (2, e), generator=gen)
    _check_index(ei, n)


def test_graph_index_flags_out_of_range():
    ei = torch.tensor([[0, 1, 7], [1, 0, 2]])
    g = _graph(ei, 3)
    with pytest.raises(IndexError):
        g.check()
    g2 = _graph(torch.tensor([[0, 1], [1, 0]]), 2, torch.tensor([1, 0]), 2)
    with pytest.raises(ValueError):
        g2.check()


def test_segment_index_matches_stable_sort():
    ops = _ops()
    gen = torch.Generator().manual_seed(1)
    idx = torch.randint(0, 97, (5000,), generator=gen)
    ptr, perm, status = ops.segment_index(idx.to(DEV), 97)
    eptr, eperm = R.csr_by_destination(idx, 97)
    assert_equal_int(ptr, eptr, "ptr")
    assert_equal_int(perm, eperm, "perm")
    assert int(status.item()) == 0


# ------------------------------------------------------------------ dense projections
@pytest.mark.parametrize("n,k,m", [(1, 1, 1), (77, 5, 16), (300, 50, 16), (1000, 50, 32), (513, 16, 32), (257, 32, 64), (130, 101, 32), (64, 82, 50), (200, 130, 70), (129, 7, 3)])
@pytest.mark.parametrize("trans_b", [True, False])
def test_node_linear(n, k, m, trans_b):
    ops = _ops()
    gen = torch.Generator().manual_seed(n * 131 + k * 7 + m)
    a = torch.randn(n, k, generator=gen)
    b = torch.randn(m, k, generator=gen) if trans_b else torch.randn(k, m, generator=gen)
    bias = torch.randn(m, generator=gen)
    mask = torch.randn(n, m, generator=gen)
    ref = a.double() @ (b.double().T if trans_b else b.double())
    got = ops.node_linear(a.to(DEV), b.to(DEV), trans_b)
    assert_close(got, ref.float(), "plain")
    got = ops.node_linear(a.to(DEV), b.to(DEV), trans_b, bias=bias.to(DEV), act=ops.ACT_RELU)
    assert_close(got, torch.relu(ref + bias.double()).float(), "bias+relu")
    got = ops.node_linear(a.to(DEV), b.to(DEV), trans_b, mask=mask.to(DEV))
    assert_close(got, torch.where(mask <= 0, torch.zeros_like(ref), ref).float(), "mask")


@pytest.mark.parametrize("n,k,m", [(5000, 50, 64), (4097, 82, 50), (2048, 64, 50), (3000, 13, 40), (2500, 128, 128), (2300, 96, 200), (2049, 1, 33)])
@pytest.mark.parametrize("trans_b", [True, False])
def test_node_linear_wide_outputs_tensor_core_path(n, k, m, trans_b, monkeypatch):
    """With ``DRK_LINEAR_TC=1`` M > 32 on >= 2048 rows runs on the tensor cores (3xTF32, ``k_node_linear_tc``; opt-in until it is
    faster than the SIMT kernel): same contract, same fp32 parity bar -- plain, bias + ReLU, ReLU-mask, ragged last row tile, K not
    a multiple of 8, several 64-column slabs."""
    from deeprank2_b200 import _lib

    monkeypatch.setenv("DRK_LINEAR_TC", "1")
    ops = _ops()
    before = _lib.launch_count()
    gen = torch.Generator().manual_seed(n + 31 * k + 7 * m)
    a = torch.randn(n, k, generator=gen) * 3.0
    b = torch.randn(m, k, generator=gen) if trans_b else torch.randn(k, m, generator=gen)
    bias = torch.randn(m, generator=gen)
    mask = torch.randn(n, m, generator=gen)
    ref = a.double() @ (b.double().T if trans_b else b.double())
    assert_close(ops.node_linear(a.to(DEV), b.to(DEV), trans_b), ref.float(), "plain")
    assert_close(ops.node_linear(a.to(DEV), b.to(DEV), trans_b, bias=bias.to(DEV), act=ops.ACT_RELU), torch.relu(ref + bias.double()).float(), "bias+relu")
    assert_close(ops.node_linear(a.to(DEV), b.to(DEV), trans_b, mask=mask.to(DEV)), torch.where(mask <= 0, torch.zeros_like(ref), ref).float(), "mask")
    r1 = ops.node_linear(a.to(DEV), b.to(DEV), trans_b)
    assert torch.equal(r1, ops.node_linear(a.to(DEV), b.to(DEV), trans_b)), "bit-reproducible"
    assert _lib.launch_count() - before == 5
    # the default (SIMT) kernel agrees with it inside the same bar
    monkeypatch.setenv("DRK_LINEAR_TC", "0")
    assert_close(ops.node_linear(a.to(DEV), b.to(DEV), trans_b), ref.float(), "plain, SIMT")


def test_node_linear2_wide_outputs_tensor_core_path(monkeypatch):
    """The concat-free two-operand Linear (``vanilla_gnn.py:37-38``: [x | message sums] -> node MLP) on the opt-in tensor-core path."""
    monkeypatch.setenv("DRK_LINEAR_TC", "1")
    ops = _ops()
    gen = torch.Generator().manual_seed(9)
    n = 4500
    a, a2 = torch.randn(n, 50, generator=gen), torch.randn(n, 32, generator=gen) * 5.0
    for trans_b in (True, False):
        b = torch.randn(50, 50, generator=gen)
        b2 = torch.randn(50, 32, generator=gen) if trans_b else torch.randn(32, 50, generator=gen)
        bias = torch.randn(50, generator=gen)
        mask = torch.randn(n, 50, generator=gen)
        ref = a.double() @ (b.double().T if trans_b else b.double()) + a2.double() @ (b2.double().T if trans_b else b2.double()) + bias.double()
        ref = torch.relu(ref)
        ref = torch.where(mask <= 0, torch.zeros_like(ref), ref)
        got = ops.node_linear2(a.to(DEV), b.to(DEV), a2.to(DEV), b2.to(DEV), trans_b, bias=bias.to(DEV), mask=mask.to(DEV), act=ops.ACT_RELU)
        assert_close(got, ref.float(), f"two operands, trans_b={trans_b}")


def test_node_linear_strided_views():
    ops = _ops()
    gen = torch.Generator().manual_seed(3)
    big = torch.randn(500, 64, generator=gen).to(DEV)
    w = torch.randn(32, 16, generator=gen).to(DEV)
    out = torch.zeros(500, 64, device=DEV)
    ops.node_linear(big[:, 16:32], w, True, out=out[:, 32:])
    assert_close(out[:, 32:], (big[:, 16:32].double().cpu() @ w.double().cpu().T).float(), "strided")
    assert float(out[:, :32].abs().max()) == 0.0


@pytest.mark.parametrize("n,k,m", [(1, 1, 1), (300, 50, 16), (5000, 50, 32), (777, 16, 32), (1000, 16, 64), (640, 101, 32), (333, 82, 50), (100, 130, 70),
                                   # row-contiguous, n >= 4096, M, K <= 64: the tensor-core kernel (ragged last tile, odd widths, one-column operands)
                                   (4096, 16, 16), (77469, 50, 64), (9999, 50, 50), (5003, 7, 33), (4100, 64, 1), (12345, 1, 64), (8191, 32, 32)])
def test_weight_grad(n, k, m):
    ops = _ops()
    gen = torch.Generator().manual_seed(n + k + m)
    dy = torch.randn(n, m, generator=gen)
    x = torch.randn(n, k, generator=gen)
    dw, db = ops.weight_grad(dy.to(DEV), x.to(DEV), want_bias=True)
    assert_close(dw, (dy.double().T @ x.double()).float(), "dW")
    assert_close(db, dy.double().sum(0).float(), "db")
    dw2 = ops.weight_grad(dy.to(DEV), x.to(DEV), dw=dw.clone(), accumulate=True)
    assert_close(dw2, 2 * (dy.double().T @ x.double()).float(), "accumulate")


# ------------------------------------------------------------------ segmented gather-reduce
def _spmm_ref(edge_index, src, n, w=None):
    vals = src[edge_index[1]].double()
    if w is not None:
        vals = vals * w.double().unsqueeze(1)
    return torch.zeros(n, src.shape[1], dtype=torch.float64).index_add_(0, edge_index[0], vals)


@pytest.mark.parametrize("case", ["toy_edgecases", "synthetic_small", "fixture_1ATN"])
@pytest.mark.parametrize("width", [16, 32, 64, 50, 5, 1, 132])
def test_spmm_sum(case, width):
    ops = _ops()
    d = load_golden(case).inputs()
    n = d.x.shape[0]
    g = _graph(d.edge_index, n)
    gen = torch.Generator().manual_seed(width)
    src = torch.randn(n, width, generator=gen)
    ref = _spmm_ref(d.edge_index, src, n)
    assert_close(ops.spmm(g.rowptr, g.colidx, src.to(DEV), n), ref.float(), "sum")
    assert_close(ops.spmm(g.rowptr, g.colidx, src.to(DEV), n, act=ops.ACT_RELU), torch.relu(ref).float(), "sum+relu")
    # transposed aggregation through the CSC half
    ref_t = torch.zeros(n, width, dtype=torch.float64).index_add_(0, d.edge_index[1], src[d.edge_index[0]].double())
    assert_close(ops.spmm(g.colptr, g.rowidx, src.to(DEV), n), ref_t.float(), "A^T")


@pytest.mark.parametrize("width", [16, 32, 6])
def test_spmm_means_weights_epilogues(width):
    ops = _ops()
    d = load_golden("toy_edgecases").inputs()
    n = d.x.shape[0]
    g = _graph(d.edge_index, n)
    gen = torch.Generator().manual_seed(width)
    src = torch.randn(n, width, generator=gen)
    add = torch.randn(n, width, generator=gen)
    mask = torch.randn(n, width, generator=gen)
    w_edge = torch.rand(d.edge_index.shape[1], generator=gen)
    deg = torch.bincount(d.edge_index[0], minlength=n).double().unsqueeze(1)
    total = _spmm_ref(d.edge_index, src, n)
    assert_close(ops.spmm(g.rowptr, g.colidx, src.to(DEV), n, reduce=ops.REDUCE_MEAN_CLAMP), (total / deg.clamp(min=1)).float(), "mean clamp")
    nan_mean = total / deg  # 0/0 -> NaN rows for isolated nodes, like torch.mean of an empty slice
    got = ops.spmm(g.rowptr, g.colidx, src.to(DEV), n, reduce=ops.REDUCE_MEAN_NAN, addend=add.to(DEV), act=ops.ACT_RELU)
    assert_close(got, torch.relu(nan_mean + add.double()).float(), "mean nan + addend + relu")
    assert bool(torch.isnan(got).any()), "toy case has isolated nodes: NaN rows expected"
    w_csr = w_edge[g.perm.cpu().long()].to(DEV)
    assert_close(ops.spmm(g.rowptr, g.colidx, src.to(DEV), n, w=w_csr), _spmm_ref(d.edge_index, src, n, w_edge).float(), "weighted")
    got = ops.spmm(g.rowptr, g.colidx, src.to(DEV), n, mask=mask.to(DEV))
    assert_close(got, torch.where(mask <= 0, torch.zeros_like(total), total).float(), "relu-mask epilogue")


def test_spmm_full_size_and_deterministic():
    ops = _ops()
    from deeprank2_b200.synthetic import make_batch

    b = make_batch(64)
    n = b.num_nodes
    g = _graph(b.edge_index, n)
    src = torch.randn(n, 32, generator=torch.Generator().manual_seed(0))
    out1 = ops.spmm(g.rowptr, g.colidx, src.to(DEV), n)
    out2 = ops.spmm(g.rowptr, g.colidx, src.to(DEV), n)
    assert torch.equal(out1, out2), "no atomics: two runs must agree bit for bit"
    assert_close(out1, _spmm_ref(b.edge_index, src, n).float(), "C2-size spmm")
    # linearity (size-independent property): A(ax + by) = a Ax + b Ay
    src2 = torch.randn(n, 32, generator=torch.Generator().manual_seed(1))
    lhs = ops.spmm(g.rowptr, g.colidx, (2.0 * src + 0.5 * src2).to(DEV), n)
    rhs = 2.0 * out1 + 0.5 * ops.spmm(g.rowptr, g.colidx, src2.to(DEV), n)
    assert_close(lhs, rhs.cpu(), "linearity", rtol=1e-4, atol_scale=1e-5)


# ------------------------------------------------------------------ readout
@pytest.mark.parametrize("width", [64, 32, 50, 7])
def test_segment_mean_fwd_bwd(width):
    ops = _ops()
    d = load_golden("synthetic_small").inputs()
    n = d.x.shape[0]
    nb = int(d.ptr.numel()) - 1
    g = _graph(d.edge_index, n, d.batch, nb)
    gen = torch.Generator().manual_seed(width)
    x = torch.randn(n, width, generator=gen)
    ref = R.mean_readout(x.double(), d.batch)
    assert_close(ops.segment_mean(x.to(DEV), g.graph_ptr, nb), ref.float(), "mean readout")
    dg = torch.randn(nb, width, generator=gen)
    cnt = torch.bincount(d.batch, minlength=nb).clamp(min=1).double()
    ref_dx = (dg.double() / cnt.unsqueeze(1))[d.batch]
    assert_close(ops.segment_mean_bwd(dg.to(DEV), g.graph_ptr, g.batch32, n), ref_dx.float(), "mean readout bwd")
    mask = torch.randn(n, width, generator=gen)
    got = ops.segment_mean_bwd(dg.to(DEV), g.graph_ptr, g.batch32, n, mask=mask.to(DEV))
    assert_close(got, torch.where(mask <= 0, torch.zeros_like(ref_dx), ref_dx).float(), "masked")


def test_segment_mean_empty_graph_is_zero():
    ops = _ops()
    batch = torch.tensor([0, 0, 2, 2, 2])
    g = _graph(torch.zeros(2, 0, dtype=torch.long), 5, batch, 4)
    x = torch.arange(10, dtype=torch.float32).reshape(5, 2)
    got = ops.segment_mean(x.to(DEV), g.graph_ptr, 4).cpu()
    assert torch.equal(got, torch.tensor([[1.0, 2.0], [0.0, 0.0], [6.0, 7.0], [0.0, 0.0]]))


def test_gather_rows():
    ops = _ops()
    gen = torch.Generator().manual_seed(0)
    src = torch.randn(100, 3, generator=gen)
    perm = torch.randperm(100, generator=gen).to(torch.int32)
    assert torch.equal(ops.gather_rows(src.to(DEV), perm.to(DEV)).cpu(), src[perm.long()])


def test_cpu_tensors_are_rejected():
    ops = _ops()
    with pytest.raises(RuntimeError):
        ops.node_linear(torch.zeros(2, 2), torch.zeros(2, 2))


@pytest.mark.parametrize("width", [16, 32, 64])
def test_spmm_tiled_is_bitwise_the_generic_kernel(width, monkeypatch):
    """drk_spmm_tiled (block-diagonal batches of large graphs: the graph's source rows staged in shared memory) against drk_spmm on an
    atom-level batch: every reduce mode / epilogue / weight combination, forward (CSR) and transposed (CSC) -- bit for bit."""
    from deeprank2_b200 import _lib, ops
    from deeprank2_b200.graph import graph_index
    from deeprank2_b200.synthetic import ATOM, make_batch

    monkeypatch.setattr(ops, "SPMM_TILED", True)  # opt-in kernel (DRK_SPMM_TILED=1)
    host = make_batch(3, first=70, n_node_features=4, n_edge_features=1, level=dict(ATOM, n_lo=1100, n_hi=1500))
    b = host.clone().to(DEV)
    gi = graph_index(b)
    assert gi.max_graph_nodes is not None and gi.max_graph_nodes >= 1024
    assert _lib.load().drk_spmm_tiled_supported(gi.max_graph_nodes, width)
    n, e = b.num_nodes, b.num_edges
    gen = torch.Generator().manual_seed(width)
    src = torch.randn(n, 2 * width, generator=gen).to(DEV)[:, width:]  # a column-sliced view: leading dimension != width
    addend = torch.randn(n, width, generator=gen).to(DEV)
    mask = torch.randn(n, width, generator=gen).to(DEV)
    w = torch.rand(e, generator=gen).to(DEV)
    for ptr, idx in ((gi.rowptr, gi.colidx), (gi.colptr, gi.rowidx)):
        for kw in (dict(), dict(act=ops.ACT_RELU), dict(reduce=ops.REDUCE_MEAN_NAN, addend=addend, act=ops.ACT_RELU), dict(reduce=ops.REDUCE_MEAN_CLAMP, w=w),
                   dict(mask=mask)):
            tiled = ops.spmm(ptr, idx, src, n, graph=gi, **kw)
            plain = ops.spmm(ptr, idx, src, n, **kw)
            assert torch.equal(tiled.isnan(), plain.isnan())
            assert torch.equal(torch.nan_to_num(tiled), torch.nan_to_num(plain)), (width, sorted(kw))
    # and it is the tiled kernel that ran: its launches are counted under their own name
    before = _lib.launch_count()
    ops.spmm(gi.rowptr, gi.colidx, src, n, graph=gi)
    assert _lib.launch_count() == before + 1
