"""GPU parity tests of the opt-in segment-softmax attention (``GINetConvLayer(attention="segment_softmax")``,
``csrc/drk_attention.cu``) against the CPU oracle ``oracle/restate.py:ginet_conv_segment_softmax`` + torch autograd.

The reference never computes this operator (its softmax runs over a singleton axis, ``ginet.py:54``), so the oracle is a
restatement of the *intended* arithmetic only -- parity unpinned, as its header says; ``tests/test_attention_oracle.py`` pins
that restatement against a float64 loop.  Tolerance: the path's fp32 bar (conftest.assert_close).
"""
from __future__ import annotations

import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close
from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _random_graph(n, e, seed, hub=None):
    """Directed multigraph with duplicates and self loops, isolated nodes (the last 3 ids never appear as a destination) and
    optionally one hub destination of large degree."""
    gen = torch.Generator().manual_seed(seed)
    row = torch.randint(0, max(n - 3, 1), (e,), generator=gen)
    col = torch.randint(0, n, (e,), generator=gen)
    if e > 4:
        row[1], col[1] = row[0], col[0]  # duplicate edge
        col[2] = row[2]  # self loop
    if hub is not None:
        row[e // 2 : e // 2 + hub] = 1
    return torch.stack([row, col])


def _layer_case(fi, fo, fe, n, e, seed, relu, hub=None):
    from deeprank2_b200.neuralnets.gnn._common import GINetConvLayer

    gen = torch.Generator().manual_seed(seed + 100)
    ei = _random_graph(n, e, seed, hub)
    x = torch.randn(n, fi, generator=gen)
    ea = torch.rand(e, fe, generator=gen) * 8.0
    proj = torch.randn(n, fo, generator=gen)  # the loss is <z, proj>: a generic cotangent
    torch.manual_seed(seed)
    layer = GINetConvLayer(fi, fo, fe, attention="segment_softmax")
    with torch.no_grad():
        layer.fc_attention.weight.mul_(3.0)  # spread the logits so the softmax is far from uniform
    p = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    z_ref = R.ginet_conv_segment_softmax(xr, ei, ea, p)
    if relu:
        z_ref = F.relu(z_ref)
    (z_ref * proj).sum().backward()

    layer = layer.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    z = layer(xg, ei.to(DEV), ea.to(DEV), relu=relu)
    (z * proj.to(DEV)).sum().backward()
    assert_close(z, z_ref, "z")
    assert_close(xg.grad, xr.grad, "dx")
    for k, v in layer.named_parameters():
        assert v.grad is not None, k
        assert_close(v.grad, p[k].grad, f"grad {k}")
    return layer, xg, z


@pytest.mark.parametrize(
    "fi,fo,fe,relu",
    [(50, 16, 1, False), (50, 16, 1, True), (16, 32, 1, True), (38, 16, 6, False), (8, 64, 2, True), (12, 24, 3, False), (20, 128, 1, False)],
)
def test_attention_layer_vs_oracle(fi, fo, fe, relu):
    _layer_case(fi, fo, fe, n=157, e=2200, seed=fi + fo, relu=relu, hub=150)


def test_attention_layer_edge_cases():
    from deeprank2_b200.neuralnets.gnn._common import GINetConvLayer

    # no edges at all: z = 0, every gradient exists and is 0
    torch.manual_seed(1)
    layer = GINetConvLayer(6, 16, 2, attention="segment_softmax").to(DEV)
    x = torch.randn(5, 6, device=DEV, requires_grad=True)
    z = layer(x, torch.zeros(2, 0, dtype=torch.long, device=DEV), torch.zeros(0, 2, device=DEV))
    assert z.shape == (5, 16) and not bool(z.any())
    z.sum().backward()
    assert not bool(x.grad.any())
    for p in layer.parameters():
        assert p.grad is not None and not bool(p.grad.any())
    # one node with one self loop: alpha = 1, z = P[0]
    _layer_case(4, 16, 1, n=1, e=1, seed=3, relu=False)
    # 1-D edge_attr is unsqueezed like ginet.py:43
    torch.manual_seed(2)
    layer = GINetConvLayer(6, 16, 1, attention="segment_softmax").to(DEV)
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 0, 0]], device=DEV)
    ea = torch.rand(4, device=DEV)
    xs = torch.randn(3, 6, device=DEV)
    assert torch.equal(layer(xs, ei, ea), layer(xs, ei, ea.unsqueeze(1)))


def test_attention_rejects_unsupported():
    from deeprank2_b200.neuralnets.gnn._common import GINetConvLayer

    with pytest.raises(ValueError):
        GINetConvLayer(4, 16, 1, attention="softmax")
    with pytest.raises(NotImplementedError):
        GINetConvLayer(4, 16, 1, bias=True, attention="segment_softmax")
    layer = GINetConvLayer(4, 6, 1, attention="segment_softmax").to(DEV)  # 6 channels: not a multiple of 4
    with pytest.raises(NotImplementedError):
        layer(torch.randn(3, 4, device=DEV), torch.tensor([[0], [1]], device=DEV), torch.rand(1, 1, device=DEV))


def test_attention_full_size_properties():
    """C2 batch (256 graphs, ~1.5 M directed edges): with a zero attention vector every coefficient is 1/deg, so the layer
    must equal the mean aggregation of the projected rows; coefficients sum to 1 per destination for any weights; two runs are
    bit-identical (no float atomics)."""
    from deeprank2_b200 import ops
    from deeprank2_b200.graph import graph_index
    from deeprank2_b200.neuralnets.gnn._common import GINetConvLayer
    from deeprank2_b200.synthetic import make_batch

    b = make_batch(256).to(DEV)
    g = graph_index(b)
    n = b.num_nodes
    torch.manual_seed(0)
    layer = GINetConvLayer(50, 16, 1, attention="segment_softmax").to(DEV)
    with torch.no_grad():
        saved = layer.fc_attention.weight.clone()
        layer.fc_attention.weight.zero_()
        z = layer(b.x, b.edge_index, b.edge_attr, graph=g)
        p = ops.node_linear(b.x, layer.fc.weight, True)
        assert_close(z, ops.spmm(g.rowptr, g.colidx, p, n, reduce=ops.REDUCE_MEAN_CLAMP), "uniform attention == mean aggregation")
        layer.fc_attention.weight.copy_(saved * 4)
    outs = []
    for _ in range(2):
        layer.zero_grad()
        x = b.x.clone().requires_grad_(True)
        z = layer(x, b.edge_index, b.edge_attr, graph=g, relu=True)
        z.square().sum().backward()
        outs.append([z.detach().clone(), x.grad.clone()] + [q.grad.clone() for q in layer.parameters()])
    for a, c in zip(*outs):
        assert torch.equal(a, c), "attention forward/backward must be bit-reproducible"
    # alpha sums to 1 over every non-empty destination: run the kernel with P = ones through the C ABI wrapper
    fn = ops.GINetAttentionConvFunction
    ones_w = torch.zeros(16, 50, device=DEV)
    xs = torch.zeros(n, 50, device=DEV)
    xs[:, 0] = 1.0
    ones_w[:, 0] = 1.0  # P = 1 everywhere
    with torch.no_grad():
        z1 = fn.apply(xs, b.edge_attr, ones_w, layer.fc_edge_attr.weight, layer.fc_attention.weight, g, False)
    deg = g.degree()
    expect = (deg > 0).float().unsqueeze(1).expand(n, 16)
    assert_close(z1, expect, "sum of coefficients per destination")


def test_ginet_with_segment_softmax_vs_oracle():
    """The whole no-cluster GINet with the opt-in attention: prediction, loss gradients of every parameter."""
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
    from deeprank2_b200.fused import step_supported
    from deeprank2_b200.synthetic import make_batch

    batch = make_batch(6, first=300)
    torch.manual_seed(0)
    net = GINet(50, 1, 1, attention="segment_softmax").eval()
    p = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    pred_ref = R.ginet_nocluster_forward(p, batch, conv=R.ginet_conv_segment_softmax)
    loss_ref = F.mse_loss(pred_ref.reshape(-1), batch.y)
    loss_ref.backward()

    net = net.to(DEV)
    gb = copy.copy(batch).clone().to(DEV)
    assert not step_supported(net, gb)  # the per-graph step kernel implements the reference's arithmetic only
    pred = net(gb)
    loss = F.mse_loss(pred.reshape(-1), gb.y)
    loss.backward()
    assert_close(pred, pred_ref, "prediction")
    assert_close(loss, loss_ref, "loss")
    for k, v in net.named_parameters():
        assert_close(v.grad, p[k].grad, f"grad {k}")


def test_clustered_ginet_with_segment_softmax_runs_and_matches_layerwise_oracle():
    """Clustered GINet with the opt-in attention on pooled graphs (pooled edge attributes, fewer nodes): forward vs the oracle's
    clustered forward with the convolution swapped."""
    from deeprank2_b200.neuralnets.gnn.ginet import GINet
    from deeprank2_b200.synthetic import make_batch

    batch = make_batch(4, first=40, with_clusters=True)
    torch.manual_seed(0)
    net = GINet(50, 1, 1, attention="segment_softmax").eval()
    params = R.as_parameters(net.state_dict())
    with torch.no_grad():
        ref = R.ginet_forward(params, copy.copy(batch).clone(), conv=R.ginet_conv_segment_softmax)
        out = net.to(DEV)(copy.copy(batch).clone().to(DEV))
    assert_close(out, ref, "clustered prediction")


def test_trainer_trains_ginet_with_segment_softmax(tmp_path, monkeypatch):
    """The Trainer API with the opt-in attention (``functools.partial`` as ``neuralnet``): runs on the layer kernels, the loss goes
    down, and the attention weights -- dead parameters in the reference arithmetic -- move."""
    import functools

    import numpy as np

    seeded = np.random.Generator(np.random.PCG64(7))
    monkeypatch.setattr(np.random, "default_rng", lambda *a, **k: seeded)
    from deeprank2_b200.dataset import InMemoryGraphDataset
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
    from deeprank2_b200.synthetic import RESIDUE, make_graph
    from deeprank2_b200.trainer import Trainer

    level = dict(RESIDUE, n_lo=30, n_hi=60)
    ds = InMemoryGraphDataset([make_graph(g, 50, 1, level=level) for g in range(24)])
    torch.manual_seed(0)
    trainer = Trainer(functools.partial(GINet, attention="segment_softmax"), ds, val_size=4, test_size=4, cuda=True, output_exporters=[])
    assert trainer.model.attention == "segment_softmax"
    w0 = {k: v.detach().clone() for k, v in trainer.model.state_dict().items()}
    trainer.train(nepoch=3, batch_size=8, validate=True, filename=str(tmp_path / "m.pth.tar"))
    assert trainer._fused is None or not trainer._fused, "the per-graph step kernel implements the reference arithmetic only"
    w1 = trainer.model.state_dict()
    for k in ("conv1.fc_attention.weight", "conv2.fc_attention.weight", "conv1_ext.fc_edge_attr.weight"):
        assert not torch.equal(w0[k], w1[k]), f"{k} must receive a gradient in segment_softmax mode"
    assert all(bool(torch.isfinite(v).all()) for v in w1.values())


@pytest.mark.parametrize("case", ["a", "b"])
def test_attention_layer_vs_golden(case):
    """CUDA layer against ``tests/golden/attention_segment_softmax.npz``: logits from the reference's own module (``ginet.py:45-52``),
    per-destination normalisation and gradients recorded in float64 (``oracle/make_golden_attention.py``)."""
    import os

    import numpy as np

    from deeprank2_b200.neuralnets.gnn._common import GINetConvLayer

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attention_segment_softmax.npz"))
    g = {k[len(case) + 1:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(case + "/")}
    fo, fi = g["w/fc.weight"].shape
    fe = g["in/edge_attr"].shape[1]
    layer = GINetConvLayer(fi, fo, fe, attention="segment_softmax")
    layer.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("w/")})
    layer = layer.to(DEV)
    x = g["in/x"].to(DEV).requires_grad_(True)
    out = layer(x, g["in/edge_index"].to(DEV), g["in/edge_attr"].to(DEV))
    assert_close(out, g["out/z"], "z")
    (out * g["gout/z"].to(DEV)).sum().backward()
    assert_close(x.grad, g["grad/x"], "dx")
    for k, v in layer.named_parameters():
        assert_close(v.grad, g["grad/" + k], f"grad {k}")
