"""CPU: host-side logic of GraphDataset / Trainer / BatchLoader (rows D and K of SURVEY.md section 8a) with a tiny
torch-only network standing in for the CUDA nets (which have no CPU path by design).

Reads the reference's HDF5 fixtures from /root/reference (build container only); skipped elsewhere.
"""
from __future__ import annotations

import os

import numpy as np
import pytest
import torch
from torch import nn

from conftest import load_golden

HDF5 = "/root/reference/tests/data/hdf5"
pytestmark = pytest.mark.skipif(not os.path.isdir(HDF5), reason="reference fixtures only exist in the build container")

DEFAULT_FEATURES = ["res_type", "polarity", "bsa", "res_depth", "hse", "info_content", "pssm"]  # tests/test_trainer.py:31-39


class TinyNet(nn.Module):
    """Net(input_shape, output_shape, input_shape_edge) / forward(data) contract of trainer.py:377 on plain torch."""

    def __init__(self, input_shape, output_shape, input_shape_edge):
        super().__init__()
        self.fc = nn.Linear(input_shape, output_shape)

    def forward(self, data):
        nb = int(data.ptr.numel()) - 1
        sums = torch.zeros(nb, data.x.shape[1], device=data.x.device).index_add_(0, data.batch, data.x)
        counts = torch.bincount(data.batch, minlength=nb).clamp(min=1).unsqueeze(1)
        return self.fc(sums / counts)


def _dataset(name="1ATN_ppi.hdf5", **kw):
    from deeprank2_b200.dataset import GraphDataset

    args = dict(node_features=DEFAULT_FEATURES, edge_features=["distance"], target="irmsd", clustering_method="mcl")
    args.update(kw)
    return GraphDataset(os.path.join(HDF5, name), **args)


def test_graph_tensors_match_load_one_graph_layout():
    from deeprank2_b200 import hdf5_lite
    from deeprank2_b200.data import Batch

    ds = _dataset()
    assert len(ds) == 4 and ds.task == "regress" and ds.classes is None
    with hdf5_lite.File(os.path.join(HDF5, "1ATN_ppi.hdf5")) as f5:
        for i, (_, entry) in enumerate(ds.index_entries):
            d = ds.get(i)
            n_half = f5[entry]["edge_features/_index"].shape[0]
            assert d.edge_index.shape == (2, 2 * n_half)  # reference pin: tests/test_query.py:73
            assert d.edge_attr.shape == (2 * n_half, 1)   # tests/test_query.py:76
            assert torch.equal(d.edge_index[:, :n_half], d.edge_index[:, n_half:].flip(0)), "(i,j) half then (j,i) half, same order"
            assert torch.equal(d.edge_attr[:n_half], d.edge_attr[n_half:])
            assert d.x.shape[1] == 50 and d.x.dtype == torch.float32 and d.pos.shape == (d.x.shape[0], 3)
            assert d.entry_names == entry and d.cluster0.shape[0] == d.x.shape[0]
    # identical to the batch the golden vectors were recorded on (oracle/make_golden.py:fixture_batch)
    g = load_golden("fixture_1ATN").inputs()
    b = Batch.from_data_list([ds.get(i) for i in range(4)])
    for k in ("x", "edge_index", "edge_attr", "y", "pos", "batch", "ptr", "cluster0", "cluster1"):
        assert torch.equal(getattr(b, k), getattr(g, k)), k
    assert b.entry_names == [e for _, e in ds.index_entries]


def test_all_features_classes_filters_and_errors():
    from deeprank2_b200.dataset import GraphDataset

    ds = GraphDataset(os.path.join(HDF5, "test.hdf5"), target="binary")
    assert ds.task == "classif" and ds.classes == [0, 1] and ds.classes_to_index == {0: 0, 1: 1}
    assert "_name" not in ds.node_features and len(ds.edge_features) == 5
    assert ds.get(0).x.shape[1] == 57  # 57 node feature columns available in test.hdf5
    filt = GraphDataset(os.path.join(HDF5, "test.hdf5"), target="binary", target_filter={"BA": "< 100"})
    assert len(filt) == 1
    with pytest.raises(ValueError, match="Missing node features"):
        GraphDataset(os.path.join(HDF5, "test.hdf5"), target="binary", node_features=["nope"])
    with pytest.raises(ValueError, match="not present"):
        GraphDataset(os.path.join(HDF5, "test.hdf5"), target="dockq")
    with pytest.raises(ValueError, match="set the target"):
        GraphDataset(os.path.join(HDF5, "test.hdf5"))
    with pytest.raises(TypeError):
        GraphDataset(42, target="binary")
    with pytest.raises(ValueError, match="must be 'classif' or 'regress'"):
        GraphDataset(os.path.join(HDF5, "test.hdf5"), target="BA")


def test_standardisation_and_inheritance():
    ds = _dataset("test.hdf5", node_features=["bsa", "hse"], target="BA", task="regress", clustering_method=None,
                  features_transform={"all": {"transform": None, "standardize": True}})
    assert set(ds.means) == {"bsa", "hse_0", "hse_1", "hse_2", "distance", "BA"}
    raw = _dataset("test.hdf5", node_features=["bsa", "hse"], target="BA", task="regress", clustering_method=None)
    x_raw = torch.cat([raw.get(i).x for i in range(len(raw))])
    assert ds.means["bsa"] == round(float(np.nanmean(x_raw[:, 0].numpy())), 1)  # rounded to one decimal like the reference
    x_std = torch.cat([ds.get(i).x for i in range(len(ds))])
    expect = (x_raw[:, 0].double() - ds.means["bsa"]) / ds.devs["bsa"]
    assert torch.allclose(x_std[:, 0].double(), expect, atol=1e-6)
    from deeprank2_b200.dataset import GraphDataset

    val = GraphDataset(os.path.join(HDF5, "valid.hdf5"), train_source=ds)
    assert val.node_features == ["bsa", "hse"] and val.target == "BA" and val.means == ds.means
    sig = _dataset(target_transform=True)
    y = sig.get(0).y
    assert torch.allclose(y, torch.sigmoid(torch.log(torch.tensor([14.919]))), atol=1e-6)


def _trainer(tmp_path, **kw):
    from deeprank2_b200.dataset import GraphDataset
    from deeprank2_b200.trainer import Trainer
    from deeprank2_b200.utils.exporters import HDF5OutputExporter

    train = _dataset("test.hdf5", target="BA", task="regress", clustering_method=None)
    val = GraphDataset(os.path.join(HDF5, "valid.hdf5"), train_source=train)
    test = GraphDataset(os.path.join(HDF5, "test.hdf5"), train_source=train)
    return Trainer(TinyNet, train, val, test, output_exporters=[HDF5OutputExporter(str(tmp_path / "out"))], **kw), train, test


def test_trainer_runs_saves_and_reloads(tmp_path):
    from deeprank2_b200.trainer import Trainer
    from deeprank2_b200.utils.exporters import HDF5OutputExporter

    trainer, train, test = _trainer(tmp_path)
    path = str(tmp_path / "model.pth.tar")
    trainer.train(nepoch=3, batch_size=2, validate=True, best_model=False, filename=path)
    assert os.path.exists(path) and os.listdir(tmp_path / "out")
    state = torch.load(path, weights_only=False)
    for key in ("data_type", "model_state", "optimizer", "optimizer_state", "lossfunction", "target", "task", "node_features", "means", "cuda", "ngpu", "features_transform"):
        assert key in state, key
    assert trainer.epoch_saved_model == 3
    trainer.test(batch_size=4)
    again = Trainer(TinyNet, dataset_test=test, pretrained_model=path, output_exporters=[HDF5OutputExporter(str(tmp_path / "out2"))])
    for a, b in zip(trainer.model.state_dict().values(), again.model.state_dict().values()):
        assert torch.equal(a, b)
    again.test()
    with pytest.raises(ValueError, match="No training dataset"):
        again.train()


def test_trainer_argument_errors(tmp_path):
    from deeprank2_b200.dataset import GraphDataset
    from deeprank2_b200.trainer import Trainer, _divide_dataset

    train = _dataset("test.hdf5", target="BA", task="regress", clustering_method=None)
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="CUDA not detected"):
            Trainer(TinyNet, train, cuda=True)
        with pytest.raises(ValueError, match="CUDA not detected"):
            Trainer(TinyNet, train, ngpu=2)  # reference: tests/test_trainer.py:629-639
    with pytest.raises(ValueError, match="at least a train or test dataset"):
        Trainer(TinyNet)
    with pytest.raises(ValueError):
        Trainer(None, train)
    orphan = GraphDataset(os.path.join(HDF5, "valid.hdf5"), target="BA", task="regress")
    with pytest.raises(ValueError, match="train_source"):
        Trainer(TinyNet, train, orphan)
    main, split = _divide_dataset(train, 0.5)
    assert len(main) == 2 and len(split) == 2 and not set(main.index_entries) & set(split.index_entries)
    with pytest.raises(ValueError):
        _divide_dataset(train, 4)
    with pytest.raises(TypeError):
        _divide_dataset(train, "half")
    t = Trainer(TinyNet, train, val_size=1, output_exporters=[])
    assert len(t.dataset_train) == 3 and len(t.dataset_val) == 1
    with pytest.raises(ValueError, match="not appropriate"):
        t.set_lossfunction(nn.CrossEntropyLoss)
    t.set_lossfunction(nn.L1Loss)
    t.configure_optimizers(torch.optim.SGD, lr=0.1, weight_decay=0.0)
    assert isinstance(t.optimizer, torch.optim.SGD)
    with pytest.raises(ValueError, match="No pretrained model"):
        t.test()


def test_classification_targets_become_class_indices(tmp_path):
    from deeprank2_b200.dataset import GraphDataset
    from deeprank2_b200.trainer import Trainer

    train = GraphDataset(os.path.join(HDF5, "test.hdf5"), target="binary", node_features=["bsa"], edge_features=["distance"])
    t = Trainer(TinyNet, train, output_exporters=[], class_weights=True)
    assert t.output_shape == 2
    pred, y = t._format_output(torch.zeros(3, 2), torch.tensor([1.0, 0.0, 1.0]))
    assert y.dtype == torch.int64 and y.tolist() == [1, 0, 1]
    t.train(nepoch=1, batch_size=4, filename=None)
    assert t.weights is not None and abs(float(t.weights.sum()) - 1.0) < 1e-6


def test_batch_loader_shards_every_global_batch():
    from deeprank2_b200.trainer import BatchLoader

    ds = _dataset("test.hdf5", target="BA", task="regress", clustering_method=None)
    whole = [b for b, _ in BatchLoader(ds, batch_size=3)]
    assert [int(b.ptr.numel()) - 1 for b in whole] == [3, 1]
    r0 = list(BatchLoader(ds, batch_size=3, rank=0, world_size=2))
    r1 = list(BatchLoader(ds, batch_size=3, rank=1, world_size=2))
    assert [g for _, g in r0] == [3, 1] == [g for _, g in r1]
    assert r0[0][0].entry_names + r1[0][0].entry_names == whole[0].entry_names
    assert r1[1][0] is None and r0[1][0].entry_names == whole[1].entry_names
