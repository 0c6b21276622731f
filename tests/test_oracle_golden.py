"""CPU: the oracle restatement (oracle/restate.py) against the golden vectors that
oracle/make_golden.py recorded by executing the reference's own model files."""
from __future__ import annotations

import pytest
import torch

from conftest import CLUSTERED_CASES, GOLDEN_CASES, assert_adam_close, assert_close, assert_equal_int, load_golden
from oracle import restate as R
from oracle import thirdparty as tp


def _layer_check(g, tag, fn):
    w = R.as_parameters(g.group(f"{tag}/w"))
    d = g.inputs()
    x = d.x.clone().requires_grad_(True)
    z = fn(x, d, w)
    assert_close(z, g.t(f"{tag}/out/z"), f"{g.case}:{tag}:z")
    z.backward(g.t(f"{tag}/gout/z"))
    assert_close(x.grad, g.t(f"{tag}/grad/x"), f"{g.case}:{tag}:dx")
    for k, p in w.items():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert_close(got, g.t(f"{tag}/grad/{k}"), f"{g.case}:{tag}:d{k}")
    return w


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_ginet_conv_layer(case):
    g = load_golden(case)
    for tag in ("ginet_conv", "ginet_conv_nc"):
        w = _layer_check(g, tag, lambda x, d, w: R.ginet_conv(x, d.edge_index, d.edge_attr, w))
        # the two attention parameters receive exact-zero gradient tensors
        assert torch.count_nonzero(g.t(f"{tag}/grad/fc_attention.weight")) == 0
        assert torch.count_nonzero(g.t(f"{tag}/grad/fc_edge_attr.weight")) == 0
        assert w["fc_attention.weight"].grad is not None and torch.count_nonzero(w["fc_attention.weight"].grad) == 0
        d = g.inputs()
        eff = R.ginet_conv_effective(d.x, d.edge_index, g.t(f"{tag}/w/fc.weight"))
        assert_close(eff, g.t(f"{tag}/out/z"), f"{case}:{tag}:effective")


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_vanilla_conv_layer(case):
    g = load_golden(case)
    _layer_check(g, "vanilla_conv", lambda x, d, w: R.vanilla_conv(x, d.edge_index, d.edge_attr, w))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fout_conv_layer(case):
    g = load_golden(case)
    _layer_check(g, "fout_conv", lambda x, d, w: R.fout_conv(x, d.edge_index, w))
    if g.inputs().x.shape[0] < 400:
        d = g.inputs()
        w = g.group("fout_conv/w")
        assert_close(R.fout_conv_loop(d.x, d.edge_index, w), g.t("fout_conv/out/z"), f"{case}:fout_loop")


@pytest.mark.parametrize("case", [c for c in GOLDEN_CASES if c != "toy_edgecases_fe3" and c != "fixture_variants_fe5"])
def test_sgat_conv_layer(case):
    g = load_golden(case)
    _layer_check(g, "sgat_conv", lambda x, d, w: R.sgat_conv(x, d.edge_index, d.edge_attr, w))


def _net_check(g, tag, forward):
    if not g.has(f"{tag}/out/pred"):
        pytest.skip(f"{tag} not recorded for {g.case}")
    p = R.as_parameters(g.group(f"{tag}/w"))
    d = g.inputs()
    opt = R.make_adam(p)
    pred, loss = R.train_step(forward, p, opt, d)
    assert_close(pred, g.t(f"{tag}/out/pred"), f"{g.case}:{tag}:pred")
    assert_close(torch.tensor(loss), g.t(f"{tag}/out/loss"), f"{g.case}:{tag}:loss")
    for k, v in p.items():
        assert_close(v.grad, g.t(f"{tag}/grad/{k}"), f"{g.case}:{tag}:grad:{k}")
        gk = f"{tag}/grad/{k}"
        assert_adam_close(v, g.t(f"{tag}/adam/{k}"), f"{g.case}:{tag}:adam:{k}", g.t(gk) if g.has(gk) else None, g.t(f"{tag}/w/{k}"))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_ginet_nocluster_train_step(case):
    _net_check(load_golden(case), "ginet_nocluster", R.ginet_nocluster_forward)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_vanilla_train_step(case):
    _net_check(load_golden(case), "vanilla", R.vanilla_forward)


@pytest.mark.parametrize("case", CLUSTERED_CASES)
def test_clustered_ginet_train_step(case):
    _net_check(load_golden(case), "ginet", R.ginet_forward)


@pytest.mark.parametrize("case", CLUSTERED_CASES)
def test_foutnet_train_step(case):
    _net_check(load_golden(case), "foutnet", R.foutnet_forward)


@pytest.mark.parametrize("case", CLUSTERED_CASES)
def test_pooling_chain_integer_exact(case):
    g = load_golden(case)
    d = g.inputs()
    c0 = R.preloaded_cluster(d.cluster0.clone(), d.batch)
    assert_equal_int(c0, g.t("pooling/out/cluster0_global"), "cluster0 offsets")
    assert_equal_int(R.preloaded_cluster_closed_form(d.cluster0.clone(), d.batch), g.t("pooling/out/cluster0_global"), "closed form")
    pooled = R.community_pool(c0, d)
    assert_equal_int(pooled.edge_index, g.t("pooling/out/pool_edge_index"), "pooled edge_index")
    assert_equal_int(pooled.batch, g.t("pooling/out/pool_batch"), "pooled batch")
    assert torch.equal(pooled.x, g.t("pooling/out/pool_x")), "segment max must be exact"
    assert_close(pooled.edge_attr, g.t("pooling/out/pool_edge_attr"), "pooled edge_attr")
    assert_close(pooled.pos, g.t("pooling/out/pool_pos"), "pooled pos")
    c1 = R.preloaded_cluster(pooled.cluster1.clone(), pooled.batch)
    assert_equal_int(c1, g.t("pooling/out/cluster1_global"), "cluster1 offsets")
    x2, b2 = tp.max_pool_x(c1, pooled.x, pooled.batch)
    assert torch.equal(x2, g.t("pooling/out/pool2_x"))
    assert_equal_int(b2, g.t("pooling/out/pool2_batch"), "pool2 batch")


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_csr_matches_stable_sort_and_scatter_order(case):
    g = load_golden(case)
    d = g.inputs()
    n = d.x.shape[0]
    rowptr, colidx, perm = R.graph_csr(d.edge_index, n)
    assert rowptr.dtype == torch.int32 and int(rowptr[-1]) == d.edge_index.shape[1]
    # destination of every CSR slot is non-decreasing and perm is ascending inside a row (stable)
    dst = d.edge_index[0][perm.long()]
    assert bool((dst[1:] >= dst[:-1]).all())
    same = dst[1:] == dst[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())
    assert_equal_int(colidx, d.edge_index[1][perm.long()], "colidx")
    assert_equal_int(R.batch_offsets(d.batch), d.ptr, "ptr")
