"""CPU, world_size 2, gloo: the data-parallel host logic (graph sharding + one flat gradient all-reduce) reproduces the
single-process gradient of the global mini-batch -- SURVEY.md 8e."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Tiny(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.a = nn.Linear(f, 8)
        self.b = nn.Linear(8, 1)
        self.unused = nn.Parameter(torch.zeros(3))  # a parameter without gradient must not break the flat buffer

    def forward(self, batch):
        nb = int(batch.ptr.numel()) - 1
        pooled = torch.zeros(nb, batch.x.shape[1]).index_add_(0, batch.batch, batch.x) / torch.bincount(batch.batch, minlength=nb).unsqueeze(1)
        return self.b(torch.relu(self.a(pooled))).reshape(-1)


def _graphs(n_graphs):
    from deeprank2_b200.synthetic import RESIDUE, make_graph

    return [make_graph(g, 6, 1, level=dict(RESIDUE, n_lo=8, n_hi=14)) for g in range(n_graphs)]


def _worker(rank, world, port, n_graphs, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeprank2_b200.data import Batch
        from deeprank2_b200.parallel import GradAllReduce, broadcast_parameters, shard_indices

        torch.manual_seed(100 + rank)  # different initial weights per rank: broadcast must fix that
        model = _Tiny(6)
        broadcast_parameters(model)
        graphs = _graphs(n_graphs)
        mine = [graphs[i] for i in shard_indices(n_graphs, rank, world)]
        sync = GradAllReduce(model)
        batch = Batch.from_data_list(mine)
        loss = torch.nn.functional.mse_loss(model(batch), batch.y)
        loss.backward()
        sync(local_weight=len(mine) / n_graphs)
        if rank == 0:
            torch.save({"grads": [None if p.grad is None else p.grad.clone() for p in model.parameters()], "weights": [p.detach().clone() for p in model.parameters()]}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_graphs", [6, 5])  # even split and ragged split (3 + 2)
def test_two_rank_gradient_equals_single_process(tmp_path, n_graphs):
    from deeprank2_b200.data import Batch

    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_graphs, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    model = _Tiny(6)
    for p, w in zip(model.parameters(), got["weights"]):
        p.data.copy_(w)
    batch = Batch.from_data_list(_graphs(n_graphs))
    torch.nn.functional.mse_loss(model(batch), batch.y).backward()
    for p, g in zip(model.parameters(), got["grads"]):
        if p.grad is None:
            assert g is None or float(g.abs().max()) == 0.0
        else:
            assert torch.allclose(p.grad, g, rtol=1e-5, atol=1e-7), (p.grad - g).abs().max()


def test_shard_indices_cover_everything_once():
    from deeprank2_b200.parallel import shard_indices

    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in shard_indices(n, r, world)]
            assert seen == list(range(n))
            sizes = [len(shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


class _TinyNet(nn.Module):
    """A CPU network with the constructor the Trainer calls (`Net(input_shape, output_shape, input_shape_edge)`)."""

    def __init__(self, input_shape, output_shape, input_shape_edge):  # noqa: ARG002
        super().__init__()
        self.a = nn.Linear(input_shape, 8)
        self.b = nn.Linear(8, output_shape)

    def forward(self, batch):
        nb = int(batch.ptr.numel()) - 1
        pooled = torch.zeros(nb, batch.x.shape[1]).index_add_(0, batch.batch, batch.x) / torch.bincount(batch.batch, minlength=nb).unsqueeze(1)
        return self.b(torch.relu(self.a(pooled)))


class _Sink:
    def __init__(self):
        self.calls = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return None

    def process(self, pass_name, epoch, names, outputs, targets_, loss):
        self.calls.append((pass_name, epoch, list(names), loss))

    def is_compatible_with(self, *a):
        return True


def _trainer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeprank2_b200.dataset import InMemoryGraphDataset
        from deeprank2_b200.trainer import Trainer

        torch.manual_seed(7 + rank)  # ranks start from different RNG states: the split and the weights must still agree
        ds = InMemoryGraphDataset(_graphs(15))
        sink = _Sink()
        trainer = Trainer(_TinyNet, ds, val_size=4, cuda=False, output_exporters=[sink])
        trainer.train(nepoch=3, batch_size=4, validate=True, filename=None, num_workers=0)
        torch.save({
            "val_entries": [str(e) for e in trainer.dataset_val.index_entries],
            "train_entries": [str(e) for e in trainer.dataset_train.index_entries],
            "weights": [p.detach().clone() for p in trainer.model.parameters()],
            "saved_epoch": trainer.epoch_saved_model,
            "calls": sink.calls,
        }, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_trainer_ranks_stay_consistent(tmp_path):
    """Two ranks (gloo): the train / validation split is the same on both, the epoch losses that drive best-model selection are the
    GLOBAL ones (identical on both ranks), the weights end bit-identical, and only rank 0 feeds the exporters -- with every graph of
    the pass, not just its own slices."""
    mp.spawn(_trainer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(str(tmp_path / "rank0.pt"), weights_only=False)
    r1 = torch.load(str(tmp_path / "rank1.pt"), weights_only=False)
    assert r0["val_entries"] == r1["val_entries"] and r0["train_entries"] == r1["train_entries"]
    assert len(r0["val_entries"]) == 4 and len(r0["train_entries"]) == 11
    assert r0["saved_epoch"] == r1["saved_epoch"]
    for a, b in zip(r0["weights"], r1["weights"]):
        assert torch.equal(a, b)
    assert r1["calls"] == [], "only rank 0 exports"
    by_pass = {}
    for pass_name, epoch, names, loss in r0["calls"]:
        by_pass.setdefault(pass_name, []).append((epoch, names, loss))
    assert all(len(names) == 11 for _, names, _ in by_pass["training"]), "the whole pass reaches rank 0's exporter"
    assert all(len(names) == 4 for _, names, _ in by_pass["validation"])
    assert all(loss is not None and loss == loss for _, _, loss in by_pass["training"])
