"""GPU: the per-graph fused VanillaConvolutionalLayer kernels (csrc/drk_vanilla.cu: one CTA per graph, forward and backward one
launch each) against the CPU oracle of ``vanilla_gnn.py:26-38`` and against the batch-level kernels they replace, over feature widths,
edge-feature counts and graph shapes that exercise every branch: ragged last tiles, rows that are not 16-byte aligned (odd F),
isolated nodes, single-node graphs, rows of more than 32 edges, more graphs than CTAs."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _graph(rng, n, f, fe, p_edge):
    from deeprank2_b200.data import Data

    iu = np.triu_indices(n, 1)
    keep = rng.random(len(iu[0])) < p_edge
    pairs = np.stack((iu[0][keep], iu[1][keep]), axis=1).astype(np.int64)
    both = np.vstack((pairs, np.flip(pairs, 1))).T if len(pairs) else np.zeros((2, 0), dtype=np.int64)
    half = rng.random((len(pairs), fe)).astype(np.float32)
    return Data(x=torch.tensor(rng.standard_normal((n, f)), dtype=torch.float), edge_index=torch.tensor(np.ascontiguousarray(both), dtype=torch.long).reshape(2, -1),
                edge_attr=torch.tensor(np.vstack((half, half)), dtype=torch.float).reshape(2 * len(pairs), fe), y=torch.tensor(rng.random(1), dtype=torch.float))


def _batch(seed, sizes, f, fe, p_edge):
    from deeprank2_b200.data import Batch

    rng = np.random.default_rng(seed)
    return Batch.from_data_list([_graph(rng, n, f, fe, p_edge) for n in sizes])


def _layer_run(layer, batch, gout, fused: bool, monkeypatch):
    from deeprank2_b200 import ops
    from deeprank2_b200.graph import graph_index

    monkeypatch.setattr(ops, "VANILLA_FUSED", fused)
    monkeypatch.setattr(ops, "VANILLA_FUSED_MIN_GRAPHS", 1)  # (the default sends batches of fewer than 80 graphs to the batch-level kernels)
    gb = batch.clone().to(DEV)
    g = graph_index(gb)
    assert ops._vanilla_fused_ok(gb.x, layer._edge_mlp[0].weight, layer._node_mlp[0].weight, g, gb.x.shape[1], gb.edge_attr.shape[1]) == fused
    x = gb.x.clone().requires_grad_(True)
    layer.zero_grad()
    z = layer(x, gb.edge_index, gb.edge_attr, graph=g)
    z.backward(gout)
    assert int(g.status.item()) == 0
    return z.detach().clone(), x.grad.clone(), {k: p.grad.clone() for k, p in layer.named_parameters()}


CASES = [
    # (sizes, F, Fe, edge probability)
    ([70, 64, 1, 130, 65, 2, 200], 50, 1, 0.08),      # ragged tiles, a single-node graph, a two-node graph
    ([40, 90, 33], 7, 0, 0.2),                        # odd F (rows not 16-byte aligned, scalar stores), no edge features
    ([100, 37, 64], 64, 3, 0.15),                     # widest F, several edge features
    ([60, 60], 16, 1, 0.9),                           # rows of more than 32 edges
    ([12] * 331, 24, 2, 0.3),                         # more graphs than CTAs: several graphs per CTA, accumulators carried across them
    ([300, 345, 256], 50, 1, 0.07),                   # C2-sized graphs
]


@pytest.mark.parametrize("sizes,f,fe,p_edge", CASES)
def test_fused_layer_vs_oracle_and_batch_level_kernels(sizes, f, fe, p_edge, monkeypatch):
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaConvolutionalLayer

    batch = _batch(7, sizes, f, fe, p_edge)
    torch.manual_seed(3)
    layer = VanillaConvolutionalLayer(f, fe)
    gout = torch.randn(batch.x.shape[0], f)
    # CPU oracle
    p = R.as_parameters(layer.state_dict())
    xr = batch.x.clone().requires_grad_(True)
    zr = R.vanilla_conv(xr, batch.edge_index, batch.edge_attr, p)
    zr.backward(gout)
    layer = layer.to(DEV)
    z, dx, grads = _layer_run(layer, batch, gout.to(DEV), True, monkeypatch)
    z2, dx2, grads2 = _layer_run(layer, batch, gout.to(DEV), True, monkeypatch)
    assert torch.equal(z, z2) and torch.equal(dx, dx2) and all(torch.equal(grads[k], grads2[k]) for k in grads), "the fused layer must be bit-reproducible"
    assert_close(z, zr, "z")
    assert_close(dx, xr.grad, "dx")
    for k, v in grads.items():
        assert_close(v, p[k].grad, f"d{k}")
    # the batch-level route computes the same function
    zu, dxu, gradsu = _layer_run(layer, batch, gout.to(DEV), False, monkeypatch)
    assert_close(z, zu, "z vs batch-level")
    assert_close(dx, dxu, "dx vs batch-level")
    for k, v in grads.items():
        assert_close(v, gradsu[k], f"d{k} vs batch-level")


def test_fused_layer_without_input_gradient_and_without_bias(monkeypatch):
    """First layer of the network: x needs no gradient (dx is not computed); a layer without biases leaves their gradients out."""
    from deeprank2_b200 import ops
    from deeprank2_b200.graph import graph_index

    batch = _batch(11, [80, 50, 129], 20, 1, 0.1)
    gb = batch.clone().to(DEV)
    g = graph_index(gb)
    torch.manual_seed(5)
    we = (torch.randn(32, 41) * 0.2).to(DEV).requires_grad_(True)
    wn = (torch.randn(20, 52) * 0.2).to(DEV).requires_grad_(True)
    gout = torch.randn(gb.x.shape[0], 20, device=DEV)
    outs = []
    monkeypatch.setattr(ops, "VANILLA_FUSED_MIN_GRAPHS", 1)
    for fused in (True, False):
        monkeypatch.setattr(ops, "VANILLA_FUSED", fused)
        we.grad = wn.grad = None
        z = ops.VanillaConvFunction.apply(gb.x, gb.edge_attr, we, None, wn, None, g)
        z.backward(gout)
        outs.append((z.detach().clone(), we.grad.clone(), wn.grad.clone()))
    for a, b, what in zip(outs[0], outs[1], ("z", "dwe", "dwn")):
        assert_close(a, b, what)


def test_fused_network_train_step_is_capturable_and_few_launches(monkeypatch):
    """VanillaNetwork train step on a collated batch: every conv layer is one launch per direction (+ one fixed-order reduction)."""
    from deeprank2_b200 import _lib, ops
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaNetwork
    from deeprank2_b200.synthetic import make_batch

    monkeypatch.setattr(ops, "VANILLA_FUSED_MIN_GRAPHS", 1)
    batch = make_batch(6)
    torch.manual_seed(1)
    net = VanillaNetwork(50, 1, 1).to(DEV)
    gb = batch.clone().to(DEV)
    pred = net(gb)  # builds and caches the graph index
    loss = torch.nn.functional.mse_loss(pred.reshape(-1), gb.y)
    before = _lib.launch_count()
    loss.backward()
    torch.cuda.synchronize()
    # backward of two fused layers = 2 x (layer kernel + reduction) + the readout's backward + the slot map of the graph index
    assert _lib.launch_count() - before <= 8


def test_fused_kernels_stay_inside_their_output_buffers():
    """Every output of the two kernels is a window of a larger sentinel-filled buffer: the guard bands on both sides must survive
    (compute-sanitizer is not available on the GPU pool, so the C ABI is called directly with guarded windows)."""
    from deeprank2_b200 import _lib
    from deeprank2_b200.graph import graph_index, stream_ptr, workspace

    lib = _lib.load()
    f, fe, guard = 50, 1, 1024
    batch = _batch(5, [70, 1, 200, 33, 129, 64], f, fe, 0.08).to(DEV)
    g = graph_index(batch)
    n, e, ngraphs = batch.x.shape[0], batch.edge_index.shape[1], g.num_graphs
    torch.manual_seed(0)
    we = (torch.randn(32, 2 * f + fe, device=DEV) * 0.2).contiguous()
    wn = (torch.randn(f, f + 32, device=DEV) * 0.2).contiguous()
    be, bn = torch.randn(32, device=DEV), torch.randn(f, device=DEV)
    attr = g.attr_in_slot_order(batch.edge_attr).contiguous()

    def window(numel, dtype, fill):
        buf = torch.full((numel + 2 * guard,), fill, dtype=dtype, device=DEV)
        return buf, buf[guard : guard + numel]

    def intact(buf, numel, fill):
        lo, hi = buf[:guard], buf[guard + numel :]
        return bool(((lo == fill) | (lo != lo)).all()) and bool(((hi == fill) | (hi != hi)).all()) if buf.dtype.is_floating_point else bool((lo == fill).all() and (hi == fill).all())

    nan = float("nan")
    bufs = {}
    for name, numel, dtype, fill in (("out", n * f, torch.float32, nan), ("s", n * 32, torch.float32, nan), ("cnt", n * 32, torch.float32, nan),
                                     ("tf", n * fe * 32, torch.float32, nan), ("mask", e, torch.int32, -7), ("dx", n * f, torch.float32, nan),
                                     ("dwe", we.numel(), torch.float32, nan), ("dbe", 32, torch.float32, nan), ("dwn", wn.numel(), torch.float32, nan),
                                     ("dbn", f, torch.float32, nan)):
        bufs[name] = window(numel, dtype, fill) + (numel, fill)
    x = batch.x.contiguous()
    p = lambda k: bufs[k][1].data_ptr()  # noqa: E731
    rc = lib.drk_vanilla_layer_fwd(x.data_ptr(), f, g.rowptr.data_ptr(), g.colidx.data_ptr(), attr.data_ptr(), fe, g.graph_ptr.data_ptr(), None, ngraphs,
                                   int(g.max_graph_nodes), we.data_ptr(), we.stride(0), be.data_ptr(), wn.data_ptr(), wn.stride(0), bn.data_ptr(), p("out"), p("s"),
                                   p("cnt"), p("tf"), p("mask"), g.status.data_ptr(), stream_ptr())
    _lib.check(rc, "drk_vanilla_layer_fwd")
    dout = torch.randn(n, f, device=DEV)
    ws = workspace(lib.drk_vanilla_layer_bwd_workspace_bytes(f, ngraphs), torch.device(DEV))
    rc = lib.drk_vanilla_layer_bwd(x.data_ptr(), p("s"), p("out"), dout.data_ptr(), p("cnt"), p("tf"), f, fe, p("mask"), g.colptr.data_ptr(), g.rowidx.data_ptr(),
                                   g.slot_map().data_ptr(), g.graph_ptr.data_ptr(), None, ngraphs, int(g.max_graph_nodes), we.data_ptr(), we.stride(0),
                                   wn.data_ptr(), wn.stride(0), p("dx"), p("dwe"), we.stride(0), p("dbe"), p("dwn"), wn.stride(0), p("dbn"),
                                   g.status.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr())
    _lib.check(rc, "drk_vanilla_layer_bwd")
    torch.cuda.synchronize()
    assert int(g.status.item()) == 0
    for name, (buf, view, numel, fill) in bufs.items():
        assert intact(buf, numel, fill), f"{name}: a guard band was overwritten"
        if view.dtype.is_floating_point:
            assert bool(torch.isfinite(view).all()), f"{name}: not every element was written"
    assert bool((bufs["mask"][1][: g.num_edges] != -7).any())


def test_fused_network_inference_equals_training_forward(monkeypatch):
    """Under ``no_grad`` the layer kernel skips the backward pass's side outputs (cnt, tf): same predictions, bit for bit."""
    from deeprank2_b200 import ops
    from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaNetwork
    from deeprank2_b200.synthetic import make_batch

    monkeypatch.setattr(ops, "VANILLA_FUSED_MIN_GRAPHS", 1)

    gb = make_batch(5, first=11).to(DEV)
    torch.manual_seed(2)
    net = VanillaNetwork(50, 1, 1).to(DEV).eval()
    pred_train = net(gb.clone())
    with torch.no_grad():
        pred_eval = net(gb.clone())
    assert torch.equal(pred_train.detach(), pred_eval)
