"""GPU parity of the drop-in modules against the golden vectors recorded from the reference
(tests/golden, oracle/make_golden.py) and against the CPU oracle on seeded inputs."""
from __future__ import annotations

import copy

import pytest
import torch

from conftest import GOLDEN_CASES, assert_adam_close, assert_close, load_golden
from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _batch_from_golden(g):
    from deeprank2_b200.data import Batch

    d = g.inputs()
    b = Batch()
    for k, v in vars(d).items():
        setattr(b, k, v.clone())
    return b.to(DEV)


def _load(module, weights):
    module.load_state_dict({k: v.clone() for k, v in weights.items()})
    return module.to(DEV)


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("tag,fo", [("ginet_conv", 16), ("ginet_conv_nc", 32)])
def test_ginet_conv_layer_vs_reference(case, tag, fo):
    from deeprank2_b200.neuralnets.gnn.ginet import GINetConvLayer

    g = load_golden(case)
    d = g.inputs()
    fi, fe = d.x.shape[1], d.edge_attr.shape[1]
    layer = _load(GINetConvLayer(fi, fo, fe), g.group(f"{tag}/w"))
    x = d.x.to(DEV).requires_grad_(True)
    z = layer(x, d.edge_index.to(DEV), d.edge_attr.to(DEV))
    assert_close(z, g.t(f"{tag}/out/z"), f"{case}:{tag}:z")
    z.backward(g.t(f"{tag}/gout/z").to(DEV))
    assert_close(x.grad, g.t(f"{tag}/grad/x"), f"{case}:{tag}:dx")
    for k, p in layer.named_parameters():
        assert p.grad is not None, f"{k} must get a gradient tensor (zeros for the dead attention weights)"
        assert_close(p.grad, g.t(f"{tag}/grad/{k}"), f"{case}:{tag}:d{k}")
    assert torch.count_nonzero(layer.fc_attention.weight.grad) == 0
    assert torch.count_nonzero(layer.fc_edge_attr.weight.grad) == 0


def _train_step_vs_golden(g, tag, module):
    module.eval()  # dropout off, as in the golden run
    opt = torch.optim.Adam(module.parameters(), lr=1e-3, weight_decay=1e-5)
    batch = _batch_from_golden(g)
    opt.zero_grad()
    pred = module(batch)
    loss = torch.nn.functional.mse_loss(pred.reshape(-1), batch.y)
    loss.backward()
    assert_close(pred, g.t(f"{tag}/out/pred"), f"{g.case}:{tag}:pred")
    assert_close(loss, g.t(f"{tag}/out/loss"), f"{g.case}:{tag}:loss")
    for k, p in module.named_parameters():
        assert_close(p.grad, g.t(f"{tag}/grad/{k}"), f"{g.case}:{tag}:grad:{k}")
    opt.step()
    for k, v in module.state_dict().items():
        gk = f"{tag}/grad/{k}"
        assert_adam_close(v, g.t(f"{tag}/adam/{k}"), f"{g.case}:{tag}:adam:{k}", g.t(gk) if g.has(gk) else None, g.t(f"{tag}/w/{k}"))


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("fused", [True, False], ids=["fused-per-graph", "stacked-layer-kernels"])
def test_ginet_nocluster_train_step_vs_reference(case, fused):
    from deeprank2_b200 import _lib
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet

    g = load_golden(case)
    d = g.inputs()
    net = _load(GINet(d.x.shape[1], 1, d.edge_attr.shape[1]), g.group("ginet_nocluster/w"))
    net.fused = fused
    before = _lib.launch_count()
    _train_step_vs_golden(g, "ginet_nocluster", net)
    launched = _lib.launch_count() - before
    # index build (5) + offsets (1) + [fused: fwd 1 + bwd 2 | stacked: 6 + 11]
    assert launched == (9 if fused else 23), launched


@pytest.mark.parametrize("case", ["toy_edgecases", "fixture_1ATN"])
def test_ginet_nocluster_layerwise_fallback_vs_reference(case, monkeypatch):
    """A user-modified net (biases, unequal branches) takes the layer-by-layer path: same parity bar."""
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet

    monkeypatch.setattr(GINet, "_stackable", lambda self: False)
    g = load_golden(case)
    d = g.inputs()
    net = _load(GINet(d.x.shape[1], 1, d.edge_attr.shape[1]), g.group("ginet_nocluster/w"))
    _train_step_vs_golden(g, "ginet_nocluster", net)


def test_ginet_nocluster_c2_batch_vs_oracle_and_deterministic():
    """Config C2 at reduced batch (32 graphs x ~300 nodes, degree ~20): CUDA path vs the CPU oracle on
    identical inputs and weights; two CUDA runs must agree bit for bit (no float atomics)."""
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
    from deeprank2_b200.synthetic import make_batch

    batch = make_batch(32)
    torch.manual_seed(0)
    net = GINet(50, 1, 1)
    params = R.as_parameters(net.state_dict())
    opt = R.make_adam(params)
    pred_ref, loss_ref = R.train_step(R.ginet_nocluster_forward, params, opt, batch)

    net = net.to(DEV).eval()
    assert net.fused
    gb = copy.copy(batch).clone().to(DEV)
    outs = []
    for _ in range(2):
        net.zero_grad()
        pred = net(gb)
        loss = torch.nn.functional.mse_loss(pred.reshape(-1), gb.y)
        loss.backward()
        outs.append((pred.detach().clone(), [p.grad.clone() for p in net.parameters()]))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b), "gradients must be bit-reproducible"
    assert_close(outs[0][0], pred_ref, "pred")
    assert_close(loss, torch.tensor(loss_ref), "loss")
    for (k, p), gcuda in zip(params.items(), outs[0][1]):
        assert_close(gcuda, p.grad, f"grad:{k}")


def test_module_on_cpu_raises():
    from deeprank2_b200.data import Batch
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet

    g = load_golden("toy_edgecases")
    d = g.inputs()
    b = Batch()
    for k, v in vars(d).items():
        setattr(b, k, v)
    with pytest.raises(RuntimeError):
        GINet(d.x.shape[1], 1, 1)(b)


# ------------------------------------------------------------------ Vanilla / Fout / SGAT layers and nets
def _layer_vs_golden(g, tag, layer, call):
    d = g.inputs()
    layer = _load(layer, g.group(f"{tag}/w"))
    x = d.x.to(DEV).requires_grad_(True)
    z = call(layer, x, d)
    assert_close(z, g.t(f"{tag}/out/z"), f"{g.case}:{tag}:z")
    z.backward(g.t(f"{tag}/gout/z").to(DEV))
    assert_close(x.grad, g.t(f"{tag}/grad/x"), f"{g.case}:{tag}:dx")
    for k, p in layer.named_parameters():
        assert p.grad is not None, k
        assert_close(p.grad, g.t(f"{tag}/grad/{k}"), f"{g.case}:{tag}:d{k}")


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_vanilla_conv_layer_vs_reference(case):
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaConvolutionalLayer

    g = load_golden(case)
    d = g.inputs()
    layer = VanillaConvolutionalLayer(d.x.shape[1], d.edge_attr.shape[1])
    _layer_vs_golden(g, "vanilla_conv", layer, lambda m, x, d: m(x, d.edge_index.to(DEV), d.edge_attr.to(DEV)))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fout_conv_layer_vs_reference(case):
    """includes the NaN rows of nodes without neighbours (toy cases) -- NaN positions must coincide"""
    from deeprank2_b200.neuralnets.gnn.foutnet import FoutLayer

    g = load_golden(case)
    d = g.inputs()
    _layer_vs_golden(g, "fout_conv", FoutLayer(d.x.shape[1], 16), lambda m, x, d: m(x, d.edge_index.to(DEV)))


@pytest.mark.parametrize("case", ["toy_edgecases", "synthetic_small", "fixture_1ATN"])
def test_sgat_conv_layer_vs_reference(case):
    from deeprank2_b200.neuralnets.gnn.sgat import SGraphAttentionLayer

    g = load_golden(case)
    d = g.inputs()
    _layer_vs_golden(g, "sgat_conv", SGraphAttentionLayer(d.x.shape[1], 16), lambda m, x, d: m(x, d.edge_index.to(DEV), d.edge_attr.to(DEV)))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_vanilla_train_step_vs_reference(case):
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import NaiveNetwork, VanillaNetwork

    assert NaiveNetwork is VanillaNetwork
    g = load_golden(case)
    d = g.inputs()
    net = _load(VanillaNetwork(d.x.shape[1], 1, d.edge_attr.shape[1]), g.group("vanilla/w"))
    _train_step_vs_golden(g, "vanilla", net)


@pytest.mark.parametrize("case", ["synthetic_small", "fixture_1ATN", "fixture_variants_fe5"])
def test_clustered_ginet_train_step_vs_reference(case):
    from deeprank2_b200.neuralnets.gnn.ginet import GINet

    g = load_golden(case)
    d = g.inputs()
    net = _load(GINet(d.x.shape[1], 1, d.edge_attr.shape[1]), g.group("ginet/w"))
    _train_step_vs_golden(g, "ginet", net)


@pytest.mark.parametrize("case", ["synthetic_small", "fixture_1ATN", "fixture_variants_fe5"])
def test_foutnet_train_step_vs_reference(case):
    from deeprank2_b200.neuralnets.gnn.foutnet import FoutNet

    g = load_golden(case)
    d = g.inputs()
    net = _load(FoutNet(d.x.shape[1], 1, d.edge_attr.shape[1]), g.group("foutnet/w"))
    _train_step_vs_golden(g, "foutnet", net)


@pytest.mark.parametrize("case", ["synthetic_small", "fixture_1ATN"])
def test_sgat_train_step_vs_reference(case):
    from deeprank2_b200.neuralnets.gnn.sgat import SGAT

    g = load_golden(case)
    d = g.inputs()
    net = _load(SGAT(d.x.shape[1], 1, d.edge_attr.shape[1]), g.group("sgat/w"))
    _train_step_vs_golden(g, "sgat", net)


@pytest.mark.parametrize("route", ["per-graph kernels", "batch-level kernels"])
def test_vanilla_c2_batch_vs_oracle(route, monkeypatch):
    """Config C4: VanillaNetwork on C2-style batches (16 graphs), CUDA vs CPU oracle incl. all gradients, through the per-graph layer
    kernels (the default only for batches of >= 80 graphs) and through the batch-level kernels."""
    from deeprank2_b200 import ops
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaNetwork
    from deeprank2_b200.synthetic import make_batch

    monkeypatch.setattr(ops, "VANILLA_FUSED_MIN_GRAPHS", 1 if route == "per-graph kernels" else 10**9)

    batch = make_batch(16)
    torch.manual_seed(1)
    net = VanillaNetwork(50, 1, 1)
    params = R.as_parameters(net.state_dict())
    pred_ref, loss_ref = R.train_step(R.vanilla_forward, params, R.make_adam(params), batch)
    net = net.to(DEV)
    gb = batch.clone().to(DEV)
    pred = net(gb)
    loss = torch.nn.functional.mse_loss(pred.reshape(-1), gb.y)
    loss.backward()
    assert_close(pred, pred_ref, "pred")
    for (k, p_ref), p in zip(params.items(), net.parameters()):
        assert_close(p.grad, p_ref.grad, f"grad:{k}")


def test_atom_level_inference_vs_oracle_c3():
    """Config C3: atom-level graphs (~3 k nodes, ~60 k directed edges each, 38 node features) are far beyond one CTA's shared
    memory, so inference runs on the layer kernels (blocked index build, projection, aggregation, readout): predictions against
    the CPU oracle, bit-reproducible, and the per-graph step kernel must report itself as not applicable."""
    from deeprank2_b200.fused import step_supported
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
    from deeprank2_b200.synthetic import ATOM, make_batch

    batch = make_batch(3, first=50, n_node_features=38, n_edge_features=1, level=ATOM)
    assert batch.num_nodes > 8000 and batch.num_edges > 150000
    torch.manual_seed(0)
    net = GINet(38, 1, 1).eval()
    params = R.as_parameters(net.state_dict())
    with torch.no_grad():
        ref = R.ginet_nocluster_forward(params, batch)
    net = net.to(DEV)
    gb = batch.clone().to(DEV)
    assert not step_supported(net, gb)
    with torch.no_grad():
        p1 = net(gb)
        p2 = net(gb)
    assert torch.equal(p1, p2)
    assert_close(p1, ref, "atom-level inference")
