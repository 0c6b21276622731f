"""GPU: the Trainer driving the CUDA nets end to end on in-memory synthetic graphs (no HDF5 on the GPU box):
a few epochs run, losses are finite and decrease, parameters live on the GPU, the step launches drk kernels,
and one Trainer epoch reproduces the CPU oracle's epoch (same batches, same weights) within tolerance."""
from __future__ import annotations

import pytest
import torch

from conftest import assert_adam_close, assert_close
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _graphs(n, clusters=False, seed=1000):
    from deeprank2_b200.synthetic import RESIDUE, make_graph

    level = dict(RESIDUE, n_lo=30, n_hi=60)
    return [make_graph(g, 50, 1, level=level, seed=seed, with_clusters=clusters) for g in range(n)]


class _Collect:
    """minimal exporter recording what the Trainer hands over"""

    def __init__(self):
        self.calls = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return None

    def process(self, pass_name, epoch, names, outputs, targets_, loss):
        self.calls.append((pass_name, epoch, list(names), list(outputs), list(targets_), loss))

    def is_compatible_with(self, *a):
        return True


@pytest.mark.parametrize("net_name", ["ginet_nocluster", "vanilla", "ginet", "foutnet", "sgat"])
def test_trainer_trains_every_net_on_gpu(net_name, tmp_path, monkeypatch):
    import numpy as np

    from deeprank2_b200 import _lib

    # the train/val/test split draws from an unseeded numpy generator (as the reference does): pin it, so that "the loss goes
    # down over three epochs" is one deterministic trajectory instead of a coin with a small failure probability
    seeded = np.random.Generator(np.random.PCG64(7))
    monkeypatch.setattr(np.random, "default_rng", lambda *a, **k: seeded)
    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.neuralnets.gnn import foutnet, ginet, ginet_nocluster, sgat, vanilla_gnn
    from deeprank2_b200.trainer import Trainer

    net = {"ginet_nocluster": ginet_nocluster.GINet, "vanilla": vanilla_gnn.VanillaNetwork, "ginet": ginet.GINet, "foutnet": foutnet.FoutNet, "sgat": sgat.SGAT}[net_name]
    clustered = net_name in ("ginet", "foutnet", "sgat")
    ds = InMemoryGraphDataset(_graphs(24, clusters=clustered), clustering_method="mcl" if clustered else None)
    sink = _Collect()
    torch.manual_seed(0)
    trainer = Trainer(net, ds, val_size=4, test_size=4, cuda=True, output_exporters=[sink])
    assert all(p.is_cuda for p in trainer.model.parameters())  # reference: tests/test_trainer.py:94-97
    before = _lib.launch_count()
    trainer.train(nepoch=3, batch_size=8, validate=True, filename=str(tmp_path / "m.pth.tar"))
    assert _lib.launch_count() > before, "the CUDA extension must be what runs"
    train_losses = [c[5] for c in sink.calls if c[0] == "training"]
    assert len(train_losses) == 4 and all(l == l for l in train_losses)  # epoch 0 eval + 3 epochs, no NaN
    assert train_losses[-1] < train_losses[0]
    names = sink.calls[0][2]
    assert len(names) == 16 and len(sink.calls[0][3]) == 16 and len(sink.calls[0][4]) == 16
    trainer.test()
    assert sink.calls[-1][0] == "testing" and len(sink.calls[-1][2]) == 4
    if net_name == "ginet_nocluster":
        from deeprank2_b200.fused import GINetFusedStep

        from deeprank2_b200.trainer import ResidentBatches

        assert isinstance(trainer._fused, GINetFusedStep), "the Trainer must drive the reference GINet through the per-graph step kernels"
        assert isinstance(trainer.train_loader, ResidentBatches), "a dataset that fits HBM is collated once and stays resident"
        assert trainer._fused._adam is not None, "default optimizer: the Adam update runs in the finalize kernel"
    else:
        from deeprank2_b200.trainer import ResidentBatches

        assert trainer._fused is False
        assert isinstance(trainer.train_loader, ResidentBatches) and trainer.train_loader.collate, "other networks: batches are cut out of the resident set on the device"


def test_device_collate_equals_host_collate():
    """ResidentGraphSet.collate(ids) on the GPU == Batch.from_data_list([...]).to(device), tensor by tensor and bit for bit
    (clustered graphs: node-, edge-, graph- and cluster-aligned attributes), and the nets give identical outputs on both."""
    from deeprank2_b200.data import Batch
    from deeprank2_b200.fused import ResidentGraphSet
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaNetwork

    graphs = _graphs(16, clusters=True)
    gset = ResidentGraphSet(graphs, "cuda")
    for ids in ([5, 2, 2, 15, 0, 9], [7], list(range(16))):
        got = gset.collate(ids)
        ref = Batch.from_data_list([graphs[i] for i in ids]).to("cuda")
        for k in ("x", "edge_index", "edge_attr", "y", "pos", "cluster0", "cluster1", "batch", "ptr", "_node_ptr32", "_edge_ptr32"):
            a, b = got.__dict__[k], ref.__dict__[k]
            assert a.is_cuda and a.dtype == b.dtype and torch.equal(a, b), k
        assert got.entry_names == ref.entry_names and got.__dict__[Batch._META_KEY] == ref.__dict__[Batch._META_KEY]
        assert got.num_graphs == len(ids) and got.num_nodes == ref.num_nodes and got.num_edges == ref.num_edges
    torch.manual_seed(0)
    net = VanillaNetwork(50, 1, 1).to("cuda").eval()
    with torch.no_grad():
        assert torch.equal(net(gset.collate([3, 1, 4])), net(Batch.from_data_list([graphs[i] for i in (3, 1, 4)]).to("cuda")))


@pytest.mark.parametrize("net_name", ["vanilla", "ginet"])
def test_device_collated_and_streamed_training_are_identical(net_name, monkeypatch):
    """Trainer.train of a network outside the per-graph step kernel: batches gathered on the device out of the resident set vs
    batches collated on the host and copied (DRK_NO_RESIDENT=1).  The batches are bit-identical, so are the trained weights."""
    import numpy as np

    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.neuralnets.gnn import ginet, vanilla_gnn
    from deeprank2_b200.trainer import BatchLoader, ResidentBatches, Trainer

    net = {"vanilla": vanilla_gnn.VanillaNetwork, "ginet": ginet.GINet}[net_name]
    clustered = net_name == "ginet"
    from deeprank2_b200 import ops

    monkeypatch.setattr(ops, "VANILLA_FUSED_MIN_GRAPHS", 1)  # the per-graph Vanilla kernels: their sums must not depend on the CTA schedule
    results = []
    for resident in (True, False):
        if resident:
            monkeypatch.delenv("DRK_NO_RESIDENT", raising=False)
        else:
            monkeypatch.setenv("DRK_NO_RESIDENT", "1")
        seeded = np.random.Generator(np.random.PCG64(3))
        monkeypatch.setattr(np.random, "default_rng", lambda *a, **k: seeded)
        ds = InMemoryGraphDataset(_graphs(20, clusters=clustered), clustering_method="mcl" if clustered else None)
        torch.manual_seed(1)
        trainer = Trainer(net, ds, cuda=True, output_exporters=[_Collect()])
        if hasattr(trainer.model, "dropout"):
            trainer.model.dropout = 0.0  # torch's dropout RNG advances identically anyway; keep the comparison about the batches
        trainer.train(nepoch=2, batch_size=8, shuffle=True, validate=False, filename=None)
        assert isinstance(trainer.train_loader, ResidentBatches if resident else BatchLoader)
        results.append(torch.cat([p.detach().reshape(-1) for p in trainer.model.parameters()]).cpu())
    assert torch.equal(results[0], results[1])


def test_trainer_epoch_matches_cpu_oracle_epoch():
    """Trainer._epoch on the GPU == the reference's loop body (oracle port) over the same three mini-batches."""
    from deeprank2_b200.data import Batch
    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
    from deeprank2_b200.trainer import Trainer

    graphs = _graphs(12)
    ds = InMemoryGraphDataset([g.clone() for g in graphs])
    torch.manual_seed(3)
    trainer = Trainer(GINet, ds, cuda=True, output_exporters=[])
    params = R.as_parameters({k: v.cpu() for k, v in trainer.model.state_dict().items()})
    opt = R.make_adam(params)
    trainer.model.eval()  # dropout off on both sides: it draws from different RNG streams on CPU and GPU
    trainer.train_loader = trainer._loader(ds, 4, False)
    loss_gpu = trainer._run_pass(trainer.train_loader, 1, "training", train=True)
    losses = []
    for start in range(0, 12, 4):
        batch = Batch.from_data_list([g.clone() for g in graphs[start : start + 4]])
        _, l = R.train_step(R.ginet_nocluster_forward, params, opt, batch, training=False)
        losses.append(l * 4)
    assert abs(loss_gpu - sum(losses) / 12) <= 1e-5 * abs(sum(losses) / 12) + 1e-7
    for (k, p_ref), p in zip(params.items(), trainer.model.parameters()):
        assert_adam_close(p, p_ref, f"weights after 3 steps: {k}")


def test_resident_and_streamed_training_agree(monkeypatch):
    """Trainer.train on a device-resident graph set (batches = id lists) vs the streamed BatchLoader path (DRK_NO_RESIDENT=1): same
    shuffles, same weights up to the order of the per-graph sums."""
    import numpy as np

    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
    from deeprank2_b200.trainer import BatchLoader, ResidentBatches, Trainer

    results = []
    for resident in (True, False):
        if resident:
            monkeypatch.delenv("DRK_NO_RESIDENT", raising=False)
        else:
            monkeypatch.setenv("DRK_NO_RESIDENT", "1")
        seeded = np.random.Generator(np.random.PCG64(3))
        monkeypatch.setattr(np.random, "default_rng", lambda *a, **k: seeded)
        ds = InMemoryGraphDataset(_graphs(20))
        torch.manual_seed(1)
        trainer = Trainer(GINet, ds, cuda=True, output_exporters=[_Collect()])
        trainer.model.dropout = 0.0  # the two paths number the graphs of a batch differently, so they would draw different masks
        trainer.train(nepoch=2, batch_size=8, shuffle=False, validate=False, filename=None)
        assert isinstance(trainer.train_loader, ResidentBatches if resident else BatchLoader)
        results.append(torch.cat([p.detach().reshape(-1) for p in trainer.model.parameters()]).cpu())
    assert float((results[0] - results[1]).abs().max()) <= 2e-6


@pytest.mark.parametrize("net_name", ["ginet_nocluster", "vanilla"])
def test_pretrained_model_reloads_on_cuda_and_tests(net_name, tmp_path, monkeypatch):
    """Train on the GPU, save, build a second Trainer from the checkpoint with ``cuda=True`` and run ``test()``: the loader of the
    pretrained path must not ask the (not yet built) model which step kernels apply, and the reloaded weights must reproduce the
    first Trainer's test predictions."""
    import numpy as np

    seeded = np.random.Generator(np.random.PCG64(11))
    monkeypatch.setattr(np.random, "default_rng", lambda *a, **k: seeded)
    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.neuralnets.gnn import ginet_nocluster, vanilla_gnn
    from deeprank2_b200.trainer import Trainer

    net = {"ginet_nocluster": ginet_nocluster.GINet, "vanilla": vanilla_gnn.VanillaNetwork}[net_name]
    graphs = _graphs(20)
    train = InMemoryGraphDataset(graphs[:14])
    test = InMemoryGraphDataset(graphs[14:], train_source=train)
    sink = _Collect()
    torch.manual_seed(0)
    trainer = Trainer(net, train, dataset_test=test, cuda=True, output_exporters=[sink])
    path = str(tmp_path / "model.pth.tar")
    trainer.train(nepoch=2, batch_size=7, validate=False, filename=path)
    trainer.test(batch_size=3)
    first = [c for c in sink.calls if c[0] == "testing"][-1]

    sink2 = _Collect()
    again = Trainer(net, dataset_test=InMemoryGraphDataset(graphs[14:], train_source=train), pretrained_model=path, cuda=True, output_exporters=[sink2])
    assert next(again.model.parameters()).is_cuda
    again.test(batch_size=3)
    second = [c for c in sink2.calls if c[0] == "testing"][-1]
    assert first[2] == second[2]
    assert_close(torch.tensor(second[3]), torch.tensor(first[3]), "predictions of the reloaded model")
    for (k, p), q in zip(trainer.model.state_dict().items(), again.model.state_dict().values()):
        assert torch.equal(p, q), k


def test_classification_target_outside_the_classes_is_reported():
    """The reference looks every target up in ``classes_to_index`` and raises for an unknown label (trainer.py:812); here the check
    rides on the pass's single read-back."""
    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.domain import targetstorage as targets
    from deeprank2_b200.neuralnets.gnn import ginet_nocluster
    from deeprank2_b200.trainer import Trainer

    graphs = _graphs(8)
    for i, g in enumerate(graphs):
        g.y = torch.tensor([float(i % 2)])
    graphs[5].y = torch.tensor([7.0])  # not a class
    ds = InMemoryGraphDataset(graphs, task=targets.CLASSIF, classes=[0, 1])
    trainer = Trainer(ginet_nocluster.GINet, ds, cuda=True, output_exporters=[_Collect()])
    with pytest.raises(ValueError, match="not one of the dataset's classes"):
        trainer.train(nepoch=1, batch_size=4, validate=False, filename=None)


def test_pinned_slab_travels_in_one_copy_and_matches_per_tensor_copies():
    """`Batch.pin_memory(only=...)` packs the tensors a step reads into one page-locked slab; `to(device, only=...)` moves the slab with a
    single copy and re-creates the views.  Same tensors, bit for bit, as the per-tensor route; the host batch stays usable."""
    from deeprank2_b200.data import Batch
    from deeprank2_b200.fused import GINetFusedStep
    from deeprank2_b200.pipeline import shallow_host_view

    plain = Batch.from_data_list(_graphs(9))
    packed = plain.clone().pin_memory(only=GINetFusedStep.FIELDS)
    slab, layout = packed.__dict__["_slab"]
    assert slab.is_pinned() and {k for k, *_ in layout} == {k for k in GINetFusedStep.FIELDS if isinstance(plain.__dict__.get(k), torch.Tensor)}
    for _ in range(2):  # the host batch can be sent again and again
        dev = shallow_host_view(packed).to("cuda", non_blocking=True, only=GINetFusedStep.FIELDS)
        torch.cuda.synchronize()
        for k in GINetFusedStep.FIELDS:
            v = plain.__dict__.get(k)
            if isinstance(v, torch.Tensor):
                got = dev.__dict__[k]
                assert got.is_cuda and got.dtype == v.dtype and tuple(got.shape) == tuple(v.shape) and torch.equal(got.cpu(), v), k
        assert torch.equal(dev.edge_attr.cpu(), plain.edge_attr), "tensors outside the slab still travel on first access"
    assert "_slab" in packed.__dict__ and not packed.x.is_cuda
