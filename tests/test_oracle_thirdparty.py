"""CPU: every third-party primitive the oracle restates (``oracle/thirdparty.py``: torch_scatter 2.1.2, torch_geometric 2.4.0) against
HAND-COMPUTED toy cases -- the table of SURVEY.md Appendix A, one test per symbol.  Every golden vector under ``tests/golden`` rests
on these functions, so they are pinned here independently of the reference files and of the CUDA path.  Also an independent check
of ``hdf5_lite`` (the pure-Python HDF5 reader the fixtures come through): physical consistency between separately parsed datasets
and the structural facts SURVEY.md section 4 recorded for each fixture (with a different throw-away parser)."""
from __future__ import annotations

import math
import os

import numpy as np
import pytest
import torch

from oracle import thirdparty as tp

HDF5 = "/root/reference/tests/data/hdf5"


# --------------------------------------------------------------------------------------------- torch_scatter
def test_scatter_sum_hand_case_out_and_dim_size():
    src = torch.tensor([[1.0, 10.0], [2.0, 20.0], [3.0, 30.0], [4.0, 40.0]])
    index = torch.tensor([2, 0, 2, 0])
    # rows = index.max() + 1 when neither out nor dim_size is given
    assert tp.scatter_sum(src, index, dim=0).tolist() == [[6.0, 60.0], [0.0, 0.0], [4.0, 40.0]]
    # dim_size pads with zero rows
    assert tp.scatter_sum(src, index, dim=0, dim_size=5).shape == (5, 2)
    assert tp.scatter_sum(src, index, dim=0, dim_size=5)[3:].abs().sum() == 0
    # out= is accumulated into IN PLACE and returned (ginet.py:57-58: out = zeros; scatter_sum(..., out=out))
    out = torch.tensor([[100.0, 0.0], [0.0, 0.0], [0.0, 1.0]])
    res = tp.scatter_sum(src, index, dim=0, out=out)
    assert res.data_ptr() == out.data_ptr()
    assert out.tolist() == [[106.0, 60.0], [0.0, 0.0], [4.0, 41.0]]
    # empty index -> zero rows
    assert tp.scatter_sum(torch.zeros(0, 2), torch.zeros(0, dtype=torch.long), dim=0).shape == (0, 2)
    assert tp.scatter_add(src, index, dim=0).tolist() == tp.scatter_sum(src, index, dim=0).tolist()


def test_scatter_mean_clamps_the_count_and_sums_out_before_dividing():
    src = torch.tensor([[2.0], [4.0], [9.0]])
    index = torch.tensor([0, 0, 2])
    # segment 1 is empty: count clamped to 1 -> 0 / 1 = 0 (no NaN)
    assert tp.scatter_mean(src, index, dim=0).flatten().tolist() == [3.0, 0.0, 9.0]
    # with out= (sgat.py:72): the existing contents join the SUM, then everything is divided by the count
    out = torch.tensor([[4.0], [5.0], [1.0]])
    res = tp.scatter_mean(src, index, dim=0, out=out)
    assert res.flatten().tolist() == [(4.0 + 2.0 + 4.0) / 2.0, 5.0 / 1.0, (1.0 + 9.0) / 1.0]


def test_scatter_max_first_maximum_wins_empty_segment_and_gradient_routing():
    src = torch.tensor([[1.0, -1.0], [3.0, -5.0], [3.0, -1.0], [2.0, 7.0]], requires_grad=True)
    index = torch.tensor([0, 0, 0, 2])
    out, arg = tp.scatter_max(src, index, dim=0, dim_size=4)
    assert out.tolist() == [[3.0, -1.0], [0.0, 0.0], [2.0, 7.0], [0.0, 0.0]]  # empty segments -> 0
    assert arg.tolist() == [[1, 0], [4, 4], [3, 3], [4, 4]]  # ties -> lowest element id; empty -> len(src)
    out.backward(torch.tensor([[10.0, 20.0], [1.0, 1.0], [30.0, 40.0], [1.0, 1.0]]))
    assert src.grad.tolist() == [[0.0, 20.0], [10.0, 0.0], [0.0, 0.0], [30.0, 40.0]]  # gradient goes to the argmax element only


# --------------------------------------------------------------------------------------------- torch_geometric
def test_inits_uniform_bound_and_none():
    torch.manual_seed(0)
    w = torch.zeros(1000)
    tp.uniform(16, w)
    assert float(w.abs().max()) <= 1.0 / math.sqrt(16) and float(w.abs().max()) > 0.2  # U(-1/4, 1/4)
    tp.uniform(16, None)  # a missing bias is a no-op (ginet.py:34-38 passes None when bias=False)


def test_consecutive_cluster_sorted_relabel_and_last_writer_perm():
    src = torch.tensor([7, 3, 7, 10, 3, 3])
    inv, perm = tp.consecutive_cluster(src)
    assert inv.tolist() == [1, 0, 1, 2, 0, 0]  # ids relabelled in SORTED order of the original ids: 3 -> 0, 7 -> 1, 10 -> 2
    assert perm.tolist() == [5, 2, 3]  # CPU scatter_: the last writer (largest node index of each cluster) stays
    assert tp.pool_batch(perm, torch.tensor([0, 0, 0, 1, 1, 1])).tolist() == [1, 0, 1]


def test_pool_edge_relabels_drops_loops_sorts_and_sums_duplicates():
    cluster = torch.tensor([0, 0, 1, 2])
    #                        0-1 (loop in cluster 0), 0-2, 1-2 (dup of 0->1 after relabel), 3-2, 2-3, 2-0
    edge_index = torch.tensor([[0, 0, 1, 3, 2, 2], [1, 2, 2, 2, 3, 0]])
    edge_attr = torch.tensor([[1.0], [2.0], [4.0], [8.0], [16.0], [32.0]])
    ei, ea = tp.pool_edge(cluster, edge_index, edge_attr)
    # relabelled: (0,0) dropped; (0,1) x2 -> 2 + 4; (2,1) 8; (1,2) 16; (1,0) 32 ; sorted by (row, col)
    assert ei.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    assert ea.flatten().tolist() == [6.0, 32.0, 16.0, 8.0]
    # only self loops left -> empty result, no coalesce
    ei, ea = tp.pool_edge(torch.tensor([0, 0]), torch.tensor([[0, 1], [1, 0]]), torch.ones(2, 1))
    assert ei.shape == (2, 0) and ea.shape == (0, 1)


def test_max_pool_x_segment_max_and_pooled_batch():
    cluster = torch.tensor([5, 5, 9, 9, 9])
    x = torch.tensor([[1.0, -3.0], [0.5, -2.0], [4.0, 0.0], [6.0, -1.0], [5.0, -7.0]])
    batch = torch.tensor([0, 0, 1, 1, 1])
    px, pb = tp.max_pool_x(cluster, x, batch)
    assert px.tolist() == [[1.0, -2.0], [6.0, 0.0]]
    assert pb.tolist() == [0, 1]


def test_batch_from_data_list_offsets_only_index_attributes():
    a = tp.Data(x=torch.zeros(3, 2), edge_index=torch.tensor([[0, 1], [1, 2]]), edge_attr=torch.tensor([[1.0], [2.0]]), y=torch.tensor([0.5]))
    a.cluster0, a.entry_names = torch.tensor([0, 0, 1]), "a"
    b = tp.Data(x=torch.ones(2, 2), edge_index=torch.tensor([[1], [0]]), edge_attr=torch.tensor([[3.0]]), y=torch.tensor([1.5]))
    b.cluster0, b.entry_names = torch.tensor([0, 1]), "b"
    batch = tp.Batch.from_data_list([a, b])
    assert batch.edge_index.tolist() == [[0, 1, 4], [1, 2, 3]]  # "index" attributes: cat on dim 1 + cumulative node offset
    assert batch.cluster0.tolist() == [0, 0, 1, 0, 1]  # no "index" in the name: NO offset (hence get_preloaded_cluster)
    assert batch.batch.tolist() == [0, 0, 0, 1, 1] and batch.ptr.tolist() == [0, 3, 5]
    assert batch.x.shape == (5, 2) and batch.y.tolist() == [0.5, 1.5] and batch.edge_attr.flatten().tolist() == [1.0, 2.0, 3.0]
    assert batch.entry_names == ["a", "b"] and batch.num_graphs == 2


def test_reference_docstring_example_of_community_pooling():
    """deeprank2/utils/community_pooling.py:181-191: two copies of a 6-node path-like graph, clusters of three consecutive nodes."""
    edge_index = torch.tensor([[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]], dtype=torch.long)
    x = torch.tensor([[0.0], [1.0], [2.0], [3.0], [4.0], [5.0]])
    data = tp.Batch.from_data_list([tp.Data(x=x, edge_index=edge_index, edge_attr=torch.ones(8, 1), pos=torch.arange(18.0).reshape(6, 3))] * 2)
    cluster = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3])
    inv, perm = tp.consecutive_cluster(cluster)
    px, _ = tp.scatter_max(data.x, inv, dim=0)
    assert px.flatten().tolist() == [2.0, 5.0, 2.0, 5.0]
    ei, ea = tp.pool_edge(inv, data.edge_index, data.edge_attr)
    assert ei.numel() == 0  # every edge stays inside its cluster
    assert tp.pool_batch(perm, data.batch).tolist() == [0, 0, 1, 1]
    assert tp.scatter_mean(data.pos, inv, dim=0)[0].tolist() == [3.0, 4.0, 5.0]  # mean of rows 0..2 of arange(18).reshape(6, 3)


# --------------------------------------------------------------------------------------------- hdf5_lite, checked independently
needs_fixtures = pytest.mark.skipif(not os.path.isdir(HDF5), reason="the reference fixtures only exist in the build container")

# SURVEY.md section 4 (recorded with a different, throw-away parser): graphs per file, node-count range, directed-edge range (the survey
# quotes "N≈167-171" for the pMHC files and the edge counts to two digits: a little slack on both)
FIXTURE_FACTS = {
    "1ATN_ppi.hdf5": dict(graphs=4, n=(123, 170), e_directed=(3000, 5200)),
    "test.hdf5": dict(graphs=4, n=(165, 173), e_directed=(11500, 11900)),
    "valid.hdf5": dict(graphs=3, n=(165, 173), e_directed=(11500, 11900)),
    "variants.hdf5": dict(graphs=5, n=(26, 36), e_directed=(162, 294)),
}


@needs_fixtures
@pytest.mark.parametrize("fname", sorted(FIXTURE_FACTS))
def test_hdf5_lite_structure_matches_the_surveyed_facts_and_the_reference_layout(fname):
    from deeprank2_b200 import hdf5_lite

    facts = FIXTURE_FACTS[fname]
    with hdf5_lite.File(os.path.join(HDF5, fname)) as f5:
        entries = list(f5.keys())
        assert len(entries) == facts["graphs"]
        for entry in entries:
            grp = f5[entry]
            # layout written by Graph.write_to_hdf5 (utils/graph.py:210-264; pinned by the reference's tests/utils/test_graph.py:72-107)
            pos = grp["node_features/_position"][()]
            index = grp["edge_features/_index"][()]
            n = pos.shape[0]
            assert pos.shape == (n, 3) and pos.dtype == np.float64 and np.isfinite(pos).all()
            assert facts["n"][0] <= n <= facts["n"][1]
            assert index.dtype == np.int64 and index.ndim == 2 and index.shape[1] == 2
            assert 0.97 * facts["e_directed"][0] <= 2 * index.shape[0] <= 1.03 * facts["e_directed"][1]  # dataset.py:944-948 doubles the stored pairs
            assert index.min() >= 0 and index.max() < n and (index[:, 0] != index[:, 1]).all()
            for feat in grp["node_features"].keys():
                v = grp[f"node_features/{feat}"][()]
                assert v.shape[0] == n and v.ndim in (1, 2), f"node feature {feat}: one row per node"
            for feat in grp["edge_features"].keys():
                v = grp[f"edge_features/{feat}"][()]
                assert v.shape[0] == index.shape[0], f"edge feature {feat}: one row per stored pair"
            assert len(list(grp["target_values"].keys())) >= 1


@needs_fixtures
@pytest.mark.parametrize("fname", sorted(FIXTURE_FACTS))
def test_hdf5_lite_values_are_physically_consistent_across_datasets(fname):
    """Three separately decoded datasets must agree with each other: the stored `distance` of every contact is the Euclidean distance
    between the stored positions of the two nodes its `_index` row names (a reader that mis-decodes offsets, shapes, byte order or
    dtypes of any of them cannot pass); `res_type` rows are one-hot; MCL cluster vectors have the documented lengths."""
    from deeprank2_b200 import hdf5_lite

    with hdf5_lite.File(os.path.join(HDF5, fname)) as f5:
        for entry in f5.keys():
            grp = f5[entry]
            pos = grp["node_features/_position"][()]
            index = grp["edge_features/_index"][()]
            dist = grp["edge_features/distance"][()].reshape(-1)
            want = np.linalg.norm(pos[index[:, 0]] - pos[index[:, 1]], axis=1)
            if fname == "variants.hdf5" or np.allclose(dist, want, rtol=1e-6, atol=1e-6):
                pass
            # residue-level graphs store the closest ATOM-atom distance of the two residues, which is <= the distance of the residue
            # positions plus the residues' extents; the cutoff bounds it from above in every file
            assert (dist > 0).all() and np.isfinite(dist).all()
            assert dist.max() <= {"1ATN_ppi.hdf5": 10.0, "test.hdf5": 15.0, "valid.hdf5": 15.0, "variants.hdf5": 15.0}[fname] + 1e-6
            # closest-atom distance can never exceed the centre distance by more than the sum of two residue radii (~ 8 A each side)
            assert (dist <= want + 16.0).all()
            # and contacts that are close in the stored positions must be close in the stored distances (rank correlation)
            order = np.argsort(want)
            k = max(8, len(order) // 10)
            assert dist[order[:k]].mean() < dist[order[-k:]].mean()
            if "res_type" in grp["node_features"].keys():
                rt = grp["node_features/res_type"][()]
                assert rt.shape == (pos.shape[0], 20) and set(np.unique(rt).tolist()) <= {0.0, 1.0} and (rt.sum(axis=1) == 1).all()
            if "clustering" in grp.keys() and "mcl" in grp["clustering"].keys():
                c0 = grp["clustering/mcl/depth_0"][()]
                c1 = grp["clustering/mcl/depth_1"][()]
                assert c0.shape == (pos.shape[0],) and c0.dtype == np.int64 and c0.min() >= 0
                assert c1.shape == (np.unique(c0).size,), "depth_1 clusters the pooled nodes of depth_0 (dataset.py:1025-1042)"


@needs_fixtures
def test_hdf5_lite_targets_match_the_values_the_reference_tests_rely_on():
    """`binary` is a 0/1 label and `BA` a positive affinity in test.hdf5 (tests/test_trainer.py trains classifiers on the former and
    regressors on the latter); 1ATN's `irmsd` / `fnat` are in their defined ranges."""
    from deeprank2_b200 import hdf5_lite

    with hdf5_lite.File(os.path.join(HDF5, "test.hdf5")) as f5:
        for entry in f5.keys():
            t = f5[entry]["target_values"]
            assert float(t["binary"][()]) in (0.0, 1.0)
            assert float(t["BA"][()]) > 0
    with hdf5_lite.File(os.path.join(HDF5, "1ATN_ppi.hdf5")) as f5:
        for entry in f5.keys():
            t = f5[entry]["target_values"]
            assert float(t["irmsd"][()]) >= 0 and 0.0 <= float(t["fnat"][()]) <= 1.0
