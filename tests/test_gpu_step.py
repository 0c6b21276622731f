"""GPU parity of the per-graph kernels (csrc/drk_ginet_step.cu): the blocked graph-index build (bit-exact against the oracle
and against drk_graph_index_build), and the whole-step kernel (prediction, loss, every gradient, weights after Adam)
against the golden vectors recorded from the reference and against the CPU oracle on seeded batches."""
from __future__ import annotations

import copy

import pytest
import torch

from conftest import GOLDEN_CASES, assert_adam_close, assert_close, assert_equal_int, load_golden
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


def _synthetic(n_graphs, first=0, **kw):
    from deeprank2_b200.synthetic import make_batch

    return make_batch(n_graphs, first=first, **kw)


def _net(fi, out, fe, weights=None, seed=0):
    from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet

    torch.manual_seed(seed)
    net = GINet(fi, out, fe)
    if weights is not None:
        net.load_state_dict({k: v.clone() for k, v in weights.items()})
    return net.to(DEV)


# ------------------------------------------------------------------------------------------------ blocked index build
def _blocked_index(batch):
    from deeprank2_b200 import _lib
    from deeprank2_b200.fused import block_info
    from deeprank2_b200.graph import stream_ptr

    lib = _lib.load()
    batch.__dict__.pop("_pairs", None)  # the blocked index build takes the directed edge list: use its offsets, not the pairs'
    batch.__dict__.pop("_pairs16", None)
    info = block_info(batch)
    n, e = batch.num_nodes, batch.num_edges
    out = {k: torch.full((n + 1 if k.endswith("ptr") else e,), -7, dtype=torch.int32, device=DEV) for k in ("rowptr", "colidx", "perm", "colptr", "rowidx", "permT")}
    assert lib.drk_graph_index_blocked_supported(info.max_nodes, info.max_edges) == 1
    rc = lib.drk_graph_index_build_blocked(
        batch.edge_index.data_ptr(), e, n, info.node_ptr.data_ptr(), info.edge_ptr.data_ptr(), info.num_graphs, info.max_nodes, info.max_edges,
        *[out[k].data_ptr() for k in ("rowptr", "colidx", "perm", "colptr", "rowidx", "permT")], info.status.data_ptr(), stream_ptr())
    _lib.check(rc, "drk_graph_index_build_blocked")
    torch.cuda.synchronize()
    return out, info


def _check_blocked_index(host_batch):
    batch = copy.copy(host_batch)
    batch.__dict__ = dict(host_batch.__dict__)
    batch = batch.clone().to(DEV)
    out, info = _blocked_index(batch)
    assert int(info.status.item()) == 0
    rowptr, colidx, perm = R.graph_csr(host_batch.edge_index, host_batch.num_nodes)
    colptr, rowidx, permT = R.graph_csc(host_batch.edge_index, host_batch.num_nodes)
    for k, ref in (("rowptr", rowptr), ("colidx", colidx), ("perm", perm), ("colptr", colptr), ("rowidx", rowidx), ("permT", permT)):
        assert_equal_int(out[k], ref, f"blocked index {k}")


def test_blocked_index_matches_oracle_synthetic():
    _check_blocked_index(_synthetic(24, first=3))


def test_blocked_index_full_c2_batch_matches_global_build():
    from deeprank2_b200.graph import GraphIndex

    host = _synthetic(256)
    batch = host.clone().to(DEV)
    out, info = _blocked_index(batch)
    assert int(info.status.item()) == 0
    ref = GraphIndex.build(batch.edge_index, batch.num_nodes, batch=batch.batch, num_graphs=256)
    for k in ("rowptr", "colidx", "perm", "colptr", "rowidx", "permT"):
        assert_equal_int(out[k], getattr(ref, k), f"blocked vs global {k}")
    assert_equal_int(info.node_ptr, ref.graph_ptr, "node offsets")


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("with_csc", [False, True])
def test_blocked_index_of_atom_level_graphs_matches_global_build(with_csc, pair, monkeypatch):
    """Graphs of ~3 k nodes / ~60 k directed edges: the edge slice does not fit shared memory, so the per-graph builder streams the
    edges twice and handles one key at a time -- one CTA per graph (k_index_blocked_large, DRK_INDEX_PAIR=0 or many graphs) or a
    cluster of two CTAs per graph, half of the edges each (k_index_blocked_pair, batches of few graphs) -- bit for bit the global
    counting sort, with a malformed edge dropped from both orders and flagged."""
    from deeprank2_b200 import _lib

    monkeypatch.setenv("DRK_INDEX_PAIR", pair)
    from deeprank2_b200.graph import GraphIndex, graph_index
    from deeprank2_b200.synthetic import ATOM, make_batch

    host = make_batch(5, first=90, n_node_features=4, n_edge_features=1, level=ATOM)
    assert host.meta("max_graph_edges") > 65535 // 2 and _lib.load().drk_graph_index_blocked_supported(host.meta("max_graph_nodes"), host.meta("max_graph_edges"))
    batch = host.clone().to(DEV)
    before = _lib.launch_count()
    gi = graph_index(batch, with_csc=with_csc)  # collated batch: the blocked builder
    assert _lib.launch_count() - before <= 2, "one per-graph kernel (+ batch offsets), not the five-kernel global sort"
    assert int(gi.status.item()) == 0
    ref = GraphIndex.build(batch.edge_index, batch.num_nodes, batch=batch.batch, num_graphs=5, with_csc=with_csc)
    keys = ("rowptr", "colidx", "perm") + (("colptr", "rowidx", "permT") if with_csc else ())
    for k in keys:
        assert_equal_int(getattr(gi, k), getattr(ref, k), f"blocked (large) vs global {k}")
    # an edge that leaves its graph is dropped and flagged
    bad = host.clone().to(DEV)
    bad.edge_index[1, 7] = bad.num_nodes - 1
    gi_bad = graph_index(bad, with_csc=with_csc)
    assert int(gi_bad.status.item()) & _lib.STATUS_CROSS_GRAPH


def test_blocked_index_edge_cases():
    """isolated nodes, duplicate edges, self loops, a single-node graph, a graph without edges, non-symmetric edges, and
    32 copies of the same edge in one warp batch (ranks inside a match group)."""
    from deeprank2_b200.data import Batch, Data

    def graph(n, edges):
        ei = torch.tensor(edges, dtype=torch.int64).reshape(-1, 2).t().contiguous()
        return Data(x=torch.zeros(n, 4), edge_index=ei, edge_attr=torch.zeros(ei.shape[1], 1), y=torch.zeros(1))

    graphs = [
        graph(5, [(0, 1), (1, 0), (0, 1), (2, 2), (3, 0), (0, 3), (0, 1)]),
        graph(1, []),
        graph(3, []),
        graph(1, [(0, 0), (0, 0)]),
        graph(4, [(1, 2)] * 40 + [(2, 1)] * 33 + [(3, 1), (1, 3)]),
        graph(7, [(i, (i * 3 + 1) % 7) for i in range(7)] + [((i * 5) % 7, i) for i in range(7)]),
    ]
    _check_blocked_index(Batch.from_data_list(graphs))


def test_blocked_index_flags_cross_graph_edges():
    host = _synthetic(4)
    host.edge_index[1, 5] = host.num_nodes - 1  # an edge from graph 0 to the last graph
    batch = host.clone().to(DEV)
    _, info = _blocked_index(batch)
    from deeprank2_b200 import _lib

    assert int(info.status.item()) & _lib.STATUS_CROSS_GRAPH


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_blocked_index_on_golden_batches(case):
    """golden batches are hand-assembled (no collate metadata): offsets are derived on the device (drk_edge_ptr)."""
    g = load_golden(case)
    batch = _batch_from_golden(g)
    out, info = _blocked_index(batch)
    assert int(info.status.item()) == 0
    d = g.inputs()
    refs = dict(zip(("rowptr", "colidx", "perm"), R.graph_csr(d.edge_index, d.x.shape[0])))
    refs.update(zip(("colptr", "rowidx", "permT"), R.graph_csc(d.edge_index, d.x.shape[0])))
    for k, ref in refs.items():
        assert_equal_int(out[k], ref, f"{case}:{k}")


# ------------------------------------------------------------------------------------------------ whole-step kernel
def _fused_step(net, batch, loss_fn=None, target_fn=None, train_mode=False, seed=None):
    from deeprank2_b200.fused import GINetFusedStep

    net.train(train_mode)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
    step = GINetFusedStep(net, opt, loss_fn or torch.nn.MSELoss(), target_fn=target_fn, seed=seed)
    return step, opt


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fused_step_vs_reference_golden(case):
    from deeprank2_b200.fused import check_status, block_info, step_supported

    g = load_golden(case)
    d = g.inputs()
    net = _net(d.x.shape[1], 1, d.edge_attr.shape[1], g.group("ginet_nocluster/w"))
    batch = _batch_from_golden(g)
    if not step_supported(net, batch):
        pytest.skip("graphs of this case do not fit the per-graph plan (covered by the layer kernels)")
    step, opt = _fused_step(net, batch)
    loss, pred = step.forward_backward(batch)
    check_status(block_info(batch))
    tag = "ginet_nocluster"
    assert_close(pred, g.t(f"{tag}/out/pred"), f"{case}:pred")
    assert_close(loss, g.t(f"{tag}/out/loss"), f"{case}:loss")
    for k, p in net.named_parameters():
        assert p.grad is not None
        assert_close(p.grad, g.t(f"{tag}/grad/{k}"), f"{case}:grad:{k}")
    opt.step()
    for k, v in net.state_dict().items():
        gk = f"{tag}/grad/{k}"
        assert_adam_close(v, g.t(f"{tag}/adam/{k}"), f"{case}:adam:{k}", g.t(gk) if g.has(gk) else None, g.t(f"{tag}/w/{k}"))


def _oracle_step(net, host_batch, loss_fn=R.regression_loss, **kw):
    params = R.as_parameters({k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    opt = R.make_adam(params)
    pred, loss = R.train_step(R.ginet_nocluster_forward, params, opt, host_batch, loss_fn=loss_fn, **kw)
    return params, pred, loss


@pytest.mark.parametrize("n_graphs,fi,fe,n", [(12, 50, 1, None), (5, 38, 3, None), (3, 7, 1, None), (9, 64, 1, 230), (4, 50, 1, 1)])
def test_fused_step_vs_oracle_mse(n_graphs, fi, fe, n):
    from deeprank2_b200.fused import block_info, check_status

    host = _synthetic(n_graphs, first=40, n_node_features=fi, n_edge_features=fe, n=n)
    net = _net(fi, 1, fe, seed=1)
    params, pred_ref, loss_ref = _oracle_step(net, host)
    batch = host.clone().to(DEV)
    step, opt = _fused_step(net, batch)
    loss, pred = step(batch)
    check_status(block_info(batch))
    assert_close(pred, pred_ref, "pred")
    assert_close(loss, torch.tensor(loss_ref), "loss")
    w_before = {k: v.detach().cpu() for k, v in _net(fi, 1, fe, seed=1).state_dict().items()}
    for (k, p_ref), p in zip(params.items(), net.parameters()):
        assert_close(p.grad, p_ref.grad, f"grad {k}")
        assert_adam_close(p, p_ref, f"adam {k}", p_ref.grad, w_before[k])


def test_fused_step_cross_entropy_two_classes():
    from deeprank2_b200.fused import block_info, check_status

    host = _synthetic(10, first=7)
    host.y = (host.y > 0.5).to(torch.float32)
    net = _net(50, 2, 1, seed=2)

    def ce(pred, y):
        return torch.nn.functional.cross_entropy(pred, y.to(torch.int64))

    params, pred_ref, loss_ref = _oracle_step(net, host, loss_fn=ce)
    batch = host.clone().to(DEV)
    step, opt = _fused_step(net, batch, loss_fn=torch.nn.CrossEntropyLoss(), target_fn=lambda b: b.y.to(torch.int64))
    loss, pred = step.forward_backward(batch)
    check_status(block_info(batch))
    assert_close(pred, pred_ref, "pred")
    assert_close(loss, torch.tensor(loss_ref), "loss")
    for (k, p_ref), p in zip(params.items(), net.parameters()):
        assert_close(p.grad, p_ref.grad, f"grad {k}")


def test_fused_inference_matches_oracle_and_module_forward():
    from deeprank2_b200.fused import ginet_infer

    host = _synthetic(16, first=100)
    net = _net(50, 1, 1, seed=3).eval()
    params = R.as_parameters({k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    with torch.no_grad():
        ref = R.ginet_nocluster_forward(params, host)
    batch = host.clone().to(DEV)
    assert_close(ginet_infer(net, batch), ref, "inference")
    with torch.no_grad():
        assert_close(net(batch), ref, "module forward under no_grad")


def test_fused_step_full_batch_properties():
    """C2 size (256 graphs): bit-reproducible, independent of the graph -> CTA schedule, and equal to the layer-kernel path."""
    from deeprank2_b200.fused import GINetFusedStep, block_info, check_status

    host = _synthetic(256)
    net = _net(50, 1, 1, seed=4).eval()
    batch = host.clone().to(DEV)
    opt = torch.optim.Adam(net.parameters(), lr=0.0)
    step = GINetFusedStep(net, opt, torch.nn.MSELoss())
    loss1, pred1 = step.forward_backward(batch)
    g1, l1, p1 = step.flat_grad.clone(), loss1.clone(), pred1.clone()
    loss2, pred2 = step.forward_backward(batch)
    assert torch.equal(g1, step.flat_grad) and torch.equal(l1, loss2) and torch.equal(p1, pred2)
    # collate order instead of the size-sorted snake schedule
    info = block_info(batch)
    info.order = None
    loss3, pred3 = step.forward_backward(batch)
    assert torch.equal(g1, step.flat_grad) and torch.equal(l1, loss3) and torch.equal(p1, pred3)
    check_status(info)
    # the autograd path through the layer kernels
    net2 = copy.deepcopy(net)
    net2.fused = False
    for p in net2.parameters():
        p.grad = None
    pred_l = net2(host.clone().to(DEV))
    loss_l = torch.nn.functional.mse_loss(pred_l.reshape(-1), batch.y)
    loss_l.backward()
    assert_close(p1, pred_l, "pred vs layer kernels")
    assert_close(l1, loss_l, "loss vs layer kernels")
    for (k, p), v in zip(net2.named_parameters(), step.views):
        assert_close(v, p.grad, f"grad {k} vs layer kernels")


def test_fused_step_dropout_mask_replayed_through_oracle():
    """Training mode: the kernel draws its own Philox keep-mask; read it back from the workspace, check its statistics and
    replay the oracle with the same mask."""
    from deeprank2_b200 import _lib
    from deeprank2_b200.fused import block_info
    from deeprank2_b200.graph import workspace

    n_graphs, fi = 64, 50
    host = _synthetic(n_graphs, first=300)
    net = _net(fi, 1, 1, seed=5)
    batch = host.clone().to(DEV)
    step, opt = _fused_step(net, batch, train_mode=True, seed=1234)
    w0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    loss, pred = step.forward_backward(batch)
    torch.cuda.synchronize()
    info = block_info(batch)
    kp = (fi + 3) // 4 * 4
    kp += 4 if (kp // 4) % 2 == 0 else 0
    ws = workspace(1, batch.x.device).view(torch.float32)
    off = n_graphs * (32 * kp + 64 * 16)
    gvec = ws[off : off + n_graphs * 64].view(n_graphs, 64).cpu()
    hvec = ws[off + n_graphs * 64 : off + n_graphs * 192].view(n_graphs, 128).cpu()
    pre = torch.nn.functional.linear(gvec, w0["fc1.weight"], w0["fc1.bias"])
    active = pre > 1e-6
    keep = torch.where(hvec != 0, torch.full_like(hvec, 1.0 / 0.6), torch.zeros_like(hvec))
    frac = float((hvec[active] != 0).float().mean())
    assert 0.55 < frac < 0.65, f"keep fraction {frac} for p = 0.4"
    keep = torch.where(active, keep, torch.full_like(keep, 1.0 / 0.6))  # inactive units: mask irrelevant
    params = R.as_parameters(w0)
    opt_ref = R.make_adam(params)
    pred_ref, loss_ref = R.train_step(R.ginet_nocluster_forward, params, opt_ref, host, training=True, keep=keep)
    assert_close(pred, pred_ref, "pred (dropout replay)")
    assert_close(loss, torch.tensor(loss_ref), "loss (dropout replay)")
    for (k, p_ref), p in zip(params.items(), net.parameters()):
        assert_close(p.grad, p_ref.grad, f"grad {k} (dropout replay)")
    # a second step draws a different mask (the step counter advanced on the device)
    h_first = hvec.clone()
    step.forward_backward(batch)
    torch.cuda.synchronize()
    hvec2 = ws[off + n_graphs * 64 : off + n_graphs * 192].view(n_graphs, 128).cpu()
    assert not torch.equal(h_first != 0, hvec2 != 0)
    assert int(step.state[0].item()) == 2
    assert _lib.launch_count() > 0


def test_trainstep_graph_capture_of_fused_step():
    """The fused step is capturable: replaying the CUDA graph gives the same loss sequence as eager steps."""
    from deeprank2_b200.fused import GINetFusedStep

    host = _synthetic(32, first=500)
    results = []
    for graphed in (False, True):
        net = _net(50, 1, 1, seed=6).eval()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
        step = GINetFusedStep(net, opt, torch.nn.MSELoss())
        batch = host.clone().to(DEV)
        losses = []
        if graphed:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(batch)
                losses.append(float(step.loss))
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step(batch)
            for _ in range(3):
                graph.replay()
                losses.append(float(step.loss))
        else:
            for _ in range(4):
                step(batch)
                losses.append(float(step.loss))
        results.append(losses)
    assert results[0] == pytest.approx(results[1], rel=1e-6)
    assert results[0][-1] < results[0][0]


def test_fused_adam_matches_torch_adam_on_its_own_state():
    """The finalize kernel's Adam update (2 launches per step) against torch.optim.Adam fed with the same gradients: parameters,
    exp_avg, exp_avg_sq and step after several steps, dead parameters (zero gradient, weight decay only) included."""
    from deeprank2_b200.fused import GINetFusedStep

    host = _synthetic(24, first=700)
    batch = host.clone().to(DEV)
    nets, opts, steps = [], [], []
    for kind in ("kernel", "torch"):
        net = _net(50, 1, 1, seed=8).eval()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True, fused=(kind == "torch"))
        step = GINetFusedStep(net, opt, torch.nn.MSELoss())
        if kind == "torch":
            step._adam = None  # gradients from the kernels, update by torch
        else:
            assert step._adam is not None
        nets.append(net), opts.append(opt), steps.append(step)
    from deeprank2_b200 import _lib

    for it in range(4):
        before = _lib.launch_count()
        l0, _ = steps[0](batch)
        assert _lib.launch_count() - before == 2
        l1, _ = steps[1](batch)
        assert_close(l0, l1, f"loss at step {it}", rtol=1e-6, atol_scale=1e-6)
    for (k, p0), p1 in zip(nets[0].named_parameters(), nets[1].parameters()):
        s0, s1 = opts[0].state[p0], opts[1].state[p1]
        assert float(s0["step"]) == float(s1["step"]) == 4.0
        assert_close(p0, p1, f"param {k}", rtol=1e-6, atol_scale=1e-6)
        assert_close(s0["exp_avg"], s1["exp_avg"], f"exp_avg {k}", rtol=1e-5, atol_scale=1e-6)
        assert_close(s0["exp_avg_sq"], s1["exp_avg_sq"], f"exp_avg_sq {k}", rtol=1e-5, atol_scale=1e-6)
    assert float(nets[0].conv1.fc_attention.weight.grad.abs().max()) == 0.0
    w0 = _net(50, 1, 1, seed=8).conv1.fc_attention.weight
    assert not torch.equal(nets[0].conv1.fc_attention.weight.detach().cpu(), w0.detach().cpu()), "weight decay moves the dead parameters"


def test_fused_adam_one_step_vs_oracle():
    from deeprank2_b200.fused import GINetFusedStep

    host = _synthetic(6, first=900)
    net = _net(50, 1, 1, seed=9).eval()
    w_before = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    params, pred_ref, loss_ref = _oracle_step(net, host)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
    step = GINetFusedStep(net, opt, torch.nn.MSELoss())
    assert step._adam is not None
    loss, pred = step(host.clone().to(DEV))
    assert_close(pred, pred_ref, "pred")
    for (k, p_ref), p in zip(params.items(), net.parameters()):
        assert_adam_close(p, p_ref, f"adam {k}", p_ref.grad, w_before[k])


def test_undirected_pairs_layout_is_bitwise_the_doubled_edge_list():
    """Collated batches of doubled graphs also carry every contact once (`_pairs`); the step kernel rebuilds the doubled list on
    the fly.  Same index, same sums: predictions, loss and gradients are bit-identical to reading the full `edge_index`."""
    from deeprank2_b200 import _lib
    from deeprank2_b200.fused import GINetFusedStep, block_info, check_status

    host = _synthetic(40, first=1200)
    assert host._pairs.shape[1] * 2 == host.edge_index.shape[1]
    assert host._pairs16.numel() == host._pairs.shape[1] and host._pairs16.dtype == torch.int32
    results = []
    for layout in (_lib.EDGES_LOCAL_PAIRS16, _lib.EDGES_UNDIRECTED_PAIRS, _lib.EDGES_DIRECTED):
        batch = host.clone().to(DEV)
        if layout != _lib.EDGES_LOCAL_PAIRS16:
            del batch.__dict__["_pairs16"]  # packed graph-local words (4 bytes per contact) are what the collate ships by default
        if layout == _lib.EDGES_DIRECTED:
            del batch.__dict__["_pairs"]
        net = _net(50, 1, 1, seed=11).eval()
        step = GINetFusedStep(net, torch.optim.SGD(net.parameters(), lr=0.0), torch.nn.MSELoss())
        loss, pred = step.forward_backward(batch)
        info = block_info(batch)
        check_status(info)
        assert info.layout == layout
        results.append((loss.clone(), pred.clone(), step.flat_grad.clone()))
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert torch.equal(a, b)


def test_collate_keeps_pairs_only_for_doubled_graphs():
    from deeprank2_b200.data import Batch, Data

    doubled = _synthetic(3, first=5)
    assert "_pairs" in doubled.__dict__ and "_pairs16" in doubled.__dict__
    odd = Data(x=torch.zeros(3, 4), edge_index=torch.tensor([[0, 1, 2], [1, 2, 0]]), edge_attr=torch.zeros(3, 1), y=torch.zeros(1))
    assert "_pairs" not in Batch.from_data_list([odd, odd]).__dict__ and "_pairs16" not in Batch.from_data_list([odd, odd]).__dict__


def test_resident_graph_set_selection_matches_collated_batch():
    """A mini-batch as a list of ids into a device-resident graph set (no collate, no copy): per-graph predictions are bit-identical
    to the same graphs collated into a batch, loss and gradients agree up to the order of the per-graph sums."""
    from deeprank2_b200.data import Batch
    from deeprank2_b200.fused import GINetFusedStep, ResidentGraphSet, check_status
    from deeprank2_b200.synthetic import make_graph

    graphs = [make_graph(g) for g in range(60, 100)]
    rset = ResidentGraphSet(graphs, DEV)
    ids = [3, 17, 5, 22, 39, 0, 8, 31, 17]  # any order, repeats allowed
    batch = Batch.from_data_list([graphs[i].clone() for i in ids]).to(DEV)
    out = []
    for resident in (False, True):
        net = _net(50, 1, 1, seed=12).eval()
        step = GINetFusedStep(net, torch.optim.SGD(net.parameters(), lr=0.0), torch.nn.MSELoss())
        if resident:
            sel, slot_ids = rset.select(ids)
            loss, pred = step.forward_backward(rset.batch, selection=sel)
            check_status(sel)
            out.append((loss.clone(), pred.clone(), step.flat_grad.clone(), slot_ids))
        else:
            loss, pred = step.forward_backward(batch)
            out.append((loss.clone(), pred.clone(), step.flat_grad.clone(), None))
    (loss_b, pred_b, grad_b, _), (loss_r, pred_r, grad_r, slot_ids) = out
    assert sorted(slot_ids) == sorted(ids)
    pool = {}
    for pos, gid in enumerate(ids):
        pool.setdefault(gid, []).append(pred_b[pos])
    for s, gid in enumerate(slot_ids):
        assert any(torch.equal(pred_r[s], p) for p in pool[gid]), f"prediction of graph {gid} differs between the two paths"
    assert_close(loss_r, loss_b, "loss", rtol=1e-6, atol_scale=1e-6)
    assert_close(grad_r, grad_b, "gradients", rtol=1e-5, atol_scale=1e-6)
    # and a full optimizer step through the public entry point
    net = _net(50, 1, 1, seed=12).eval()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
    step = GINetFusedStep(net, opt, torch.nn.MSELoss())
    l0, _, _ = step.step_selection(rset, ids)
    l0 = float(l0)
    for _ in range(5):
        l1, _, _ = step.step_selection(rset, ids)
    assert float(l1) < l0


def test_captured_selection_step_equals_eager_selection_steps():
    """The CUDA-graph replay of the step on a resident graph set: same loss sequence and weights as the eager id-list steps, and the
    capture itself leaves parameters and optimizer state untouched."""
    from deeprank2_b200.fused import CapturedSelectionStep, GINetFusedStep, ResidentGraphSet
    from deeprank2_b200.synthetic import make_graph

    rset = ResidentGraphSet([make_graph(g) for g in range(200, 248)], DEV)
    batches = [list(range(0, 16)), list(range(16, 32)), [47, 3, 9, 21, 33, 40, 41, 42, 5, 6, 7, 8, 30, 31, 32, 2], list(range(32, 48))]
    outs = []
    for captured in (False, True):
        net = _net(50, 1, 1, seed=13).eval()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
        step = GINetFusedStep(net, opt, torch.nn.MSELoss())
        before = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
        run = CapturedSelectionStep(step, rset, 16) if captured else (lambda ids: step.step_selection(rset, ids))
        if captured:
            assert torch.equal(before, torch.cat([p.detach().reshape(-1) for p in net.parameters()])), "capture must not train"
            assert all(float(opt.state[p]["step"]) == 0.0 for p in net.parameters())
        losses = [float(run(ids)[0]) for ids in batches]
        outs.append((losses, torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()))
    assert outs[0][0] == pytest.approx(outs[1][0], rel=1e-6)
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 1e-7


def test_blocked_index_without_perm_gives_the_same_csr():
    """Inference without edge attributes skips `perm` (the edge id of every CSR slot): rowptr / colidx must not change, for the
    residue-level kernel and for the atom-level one (no edge stash)."""
    from deeprank2_b200.graph import graph_index
    from deeprank2_b200.synthetic import ATOM, make_batch

    for kwargs in (dict(n_graphs=6), dict(n_graphs=2, first=50, n_node_features=38, n_edge_features=1, level=ATOM)):
        host = make_batch(**kwargs)
        full = graph_index(host.clone().to(DEV), with_csc=False)
        lean = graph_index(host.clone().to(DEV), with_csc=False, with_perm=False)
        assert full.perm is not None and lean.perm is None
        assert torch.equal(full.rowptr, lean.rowptr) and torch.equal(full.colidx, lean.colidx)
        assert int(lean.status.item()) == 0
