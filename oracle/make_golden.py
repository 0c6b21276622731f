"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by EXECUTING THE REFERENCE.

Run in the build container only (needs ``/root/reference``):

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

Every vector stored here comes from the unmodified reference model files
(``deeprank2/neuralnets/gnn/{ginet,ginet_nocluster,vanilla_gnn,foutnet,sgat}.py`` and
``deeprank2/utils/community_pooling.py``) imported by ``oracle/reference_loader.py``; the
third-party primitives underneath them are the pure-torch restatements of
``oracle/thirdparty.py`` (the wheels are absent from this image).  The reference's own
test-suite pins no numeric output of these layers, so these
vectors are the pin for both ``oracle/restate.py`` and the CUDA path.

npz key scheme:  ``in/<tensor>``   inputs (batch fields)
                 ``w/<state_dict key>``  weights the reference module was run with
                 ``out/<name>``   forward outputs (``pred``, ``loss``, layer ``z`` ...)
                 ``gout/<name>``  upstream gradient fed to ``backward`` (layer cases)
                 ``grad/<key>``   parameter / input gradients
                 ``adam/<key>``   weights after ONE Adam(lr 1e-3, wd 1e-5) step
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import thirdparty as tp  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
HDF5_DIR = "/root/reference/tests/data/hdf5"

DEFAULT_NODE_FEATURES = ["res_type", "polarity", "bsa", "res_depth", "hse", "info_content", "pssm"]  # tests/test_trainer.py:31-39


# --------------------------------------------------------------------------- inputs
def toy_batch(f_node=5, f_edge=1, seed=7):
    """Edge cases the path must survive: isolated node, duplicate edges,
    self loop, a single-node graph without edges, a graph whose edges are not symmetric."""
    gen = torch.Generator().manual_seed(seed)
    graphs = []
    # g0: 6 nodes, node 5 isolated, duplicate edge (0,1) x2, self loop (2,2)
    ei = torch.tensor([[0, 0, 1, 2, 2, 3, 4, 1, 3], [1, 1, 0, 2, 3, 2, 0, 4, 4]])
    graphs.append((6, ei))
    # g1: single node, no edges
    graphs.append((1, torch.zeros(2, 0, dtype=torch.long)))
    # g2: 4 nodes, directed (non symmetric) ring + chord, node 3 has in-degree 0 as destination
    graphs.append((4, torch.tensor([[0, 1, 2, 0, 2], [1, 2, 0, 3, 3]])))
    # g3: 2 nodes fully connected both ways
    graphs.append((2, torch.tensor([[0, 1], [1, 0]])))
    out = []
    for n, ei in graphs:
        e = ei.shape[1]
        d = tp.Data(
            x=torch.randn(n, f_node, generator=gen),
            edge_index=ei.long(),
            edge_attr=torch.rand(e, f_edge, generator=gen),
            y=torch.rand(1, generator=gen),
            pos=torch.randn(n, 3, generator=gen),
        )
        out.append(d)
    return tp.Batch.from_data_list(out)


def synthetic_batch(n_graphs=6, n_lo=40, n_hi=64, f_node=50, f_edge=1, seed=1000, clusters=False):
    from deeprank2_b200.synthetic import RESIDUE, make_graph

    level = dict(RESIDUE, n_lo=n_lo, n_hi=n_hi)
    gs = []
    for g in range(n_graphs):
        d = make_graph(g, f_node, f_edge, level=level, seed=seed, with_clusters=clusters)
        fields = dict(x=d.x, edge_index=d.edge_index, edge_attr=d.edge_attr, y=d.y, pos=d.pos)
        od = tp.Data(**fields)
        if clusters:
            od.cluster0, od.cluster1 = d.cluster0, d.cluster1
        gs.append(od)
    return tp.Batch.from_data_list(gs)


def fixture_batch(fname="1ATN_ppi.hdf5", node_features=DEFAULT_NODE_FEATURES, edge_features=("distance",), target="irmsd", clustering="mcl"):
    """The real residue-level PPI graphs of ``tests/data/hdf5`` loaded the way
    ``GraphDataset.load_one_graph`` does (``dataset.py:883-1052``)."""
    from deeprank2_b200 import hdf5_lite

    gs = []
    with hdf5_lite.File(os.path.join(HDF5_DIR, fname)) as f5:
        for entry in f5.keys():
            grp = f5[entry]
            cols = []
            for feat in node_features:
                v = grp[f"node_features/{feat}"][()]
                cols.append(v.reshape(-1, 1) if v.ndim == 1 else v)
            x = torch.tensor(np.hstack(cols), dtype=torch.float)
            ind = grp["edge_features/_index"][()]
            ind = np.vstack((ind, np.flip(ind, 1))).T
            edge_index = torch.tensor(np.ascontiguousarray(ind), dtype=torch.long).contiguous()
            ecols = []
            for feat in edge_features:
                v = grp[f"edge_features/{feat}"][()]
                ecols.append(v.reshape(-1, 1) if v.ndim == 1 else v)
            e = np.hstack(ecols)
            edge_attr = torch.tensor(np.vstack((e, e)), dtype=torch.float).contiguous()
            y = torch.tensor([grp[f"target_values/{target}"][()]], dtype=torch.float)
            pos = torch.tensor(grp["node_features/_position"][()], dtype=torch.float)
            d = tp.Data(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos)
            if clustering is not None:
                d.cluster0 = torch.tensor(grp[f"clustering/{clustering}/depth_0"][()], dtype=torch.long)
                d.cluster1 = torch.tensor(grp[f"clustering/{clustering}/depth_1"][()], dtype=torch.long)
            d.entry_names = entry
            gs.append(d)
    return tp.Batch.from_data_list(gs)


# --------------------------------------------------------------------------- recording
def _np(t):
    return t.detach().cpu().numpy().copy()  # copy: state_dict tensors alias the live parameters


def record_inputs(store, batch):
    for k in ("x", "edge_index", "edge_attr", "y", "pos", "batch", "ptr", "cluster0", "cluster1"):
        v = getattr(batch, k, None)
        if isinstance(v, torch.Tensor):
            store[f"in/{k}"] = _np(v)


def fresh(batch):
    return batch.clone()


def run_net(store, tag, module, batch, regress=True):
    """forward / MSE loss / backward / one Adam step with the reference module."""
    module.eval()  # dropout off: the only stochastic op on the path (ginet.py:122)
    for k, v in module.state_dict().items():
        store[f"{tag}/w/{k}"] = _np(v)
    opt = torch.optim.Adam(module.parameters(), lr=1e-3, weight_decay=1e-5)  # trainer.py:404-419
    opt.zero_grad()
    data = fresh(batch)
    pred = module(data)
    loss = torch.nn.functional.mse_loss(pred.reshape(-1), batch.y)  # trainer.py:828-831 + MSELoss
    loss.backward()
    store[f"{tag}/out/pred"] = _np(pred)
    store[f"{tag}/out/loss"] = _np(loss)
    for k, p in module.named_parameters():
        store[f"{tag}/grad/{k}"] = _np(p.grad if p.grad is not None else torch.full_like(p, float("nan")))
    opt.step()
    for k, v in module.state_dict().items():
        store[f"{tag}/adam/{k}"] = _np(v)


def run_layer(store, tag, layer, args, seed=3):
    """layer forward + backward w.r.t. its parameters and node input, random upstream grad."""
    for k, v in layer.state_dict().items():
        store[f"{tag}/w/{k}"] = _np(v)
    x = args[0].clone().requires_grad_(True)
    z = layer(x, *[a.clone() for a in args[1:]])
    gout = torch.randn(z.shape, generator=torch.Generator().manual_seed(seed))
    # NaN rows (FoutLayer on an empty neighbourhood) poison every gradient: record them as they are
    z.backward(gout)
    store[f"{tag}/out/z"] = _np(z)
    store[f"{tag}/gout/z"] = _np(gout)
    store[f"{tag}/grad/x"] = _np(x.grad)
    for k, p in layer.named_parameters():
        store[f"{tag}/grad/{k}"] = _np(p.grad)


def pooling_case(store, tag, ref, batch):
    """Integer-exact goldens for rows I/J: get_preloaded_cluster, community_pooling, max_pool_x."""
    cp = ref.community_pooling
    data = fresh(batch)
    c0 = cp.get_preloaded_cluster(data.cluster0, data.batch)
    store[f"{tag}/out/cluster0_global"] = _np(c0)
    pooled = cp.community_pooling(c0, data)
    store[f"{tag}/out/pool_x"] = _np(pooled.x)
    store[f"{tag}/out/pool_edge_index"] = _np(pooled.edge_index)
    store[f"{tag}/out/pool_edge_attr"] = _np(pooled.edge_attr)
    store[f"{tag}/out/pool_batch"] = _np(pooled.batch)
    store[f"{tag}/out/pool_pos"] = _np(pooled.pos)
    c1 = cp.get_preloaded_cluster(pooled.cluster1, pooled.batch)
    store[f"{tag}/out/cluster1_global"] = _np(c1)
    x2, b2 = tp.max_pool_x(c1, pooled.x, pooled.batch)
    store[f"{tag}/out/pool2_x"] = _np(x2)
    store[f"{tag}/out/pool2_batch"] = _np(b2)


# --------------------------------------------------------------------------- cases
def build_case(ref, name, batch, f_node, f_edge, clustered, nets=("ginet_nocluster", "vanilla", "ginet", "foutnet", "sgat")):
    store = {}
    record_inputs(store, batch)
    torch.manual_seed(0)
    ei, ea, x = batch.edge_index, batch.edge_attr, batch.x

    # layers (rows B, F, H + SGAT)
    run_layer(store, "ginet_conv", ref.ginet.GINetConvLayer(f_node, 16, f_edge), (x, ei, ea))
    run_layer(store, "ginet_conv_nc", ref.ginet_nocluster.GINetConvLayer(f_node, 32, f_edge), (x, ei, ea))
    run_layer(store, "vanilla_conv", ref.vanilla_gnn.VanillaConvolutionalLayer(f_node, f_edge), (x, ei, ea))
    run_layer(store, "fout_conv", ref.foutnet.FoutLayer(f_node, 16), (x, ei))
    if f_edge == 1:  # sgat.py:68 broadcasts edge_attr [E,Fe] against [E,Fo]: only Fe == 1 runs
        run_layer(store, "sgat_conv", ref.sgat.SGraphAttentionLayer(f_node, 16), (x, ei, ea))

    # nets (rows C, G, H)
    if "ginet_nocluster" in nets:
        run_net(store, "ginet_nocluster", ref.ginet_nocluster.GINet(f_node, 1, f_edge), batch)
    if "vanilla" in nets:
        run_net(store, "vanilla", ref.vanilla_gnn.VanillaNetwork(f_node, 1, f_edge), batch)
    if clustered:
        pooling_case(store, "pooling", ref, batch)
        if "ginet" in nets:
            run_net(store, "ginet", ref.ginet.GINet(f_node, 1, f_edge), batch)
        if "foutnet" in nets:
            run_net(store, "foutnet", ref.foutnet.FoutNet(f_node, 1, f_edge), batch)
        if "sgat" in nets and f_edge == 1:
            run_net(store, "sgat", ref.sgat.SGAT(f_node, 1, f_edge), batch)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **store)
    print(f"{name}: {len(store)} arrays -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(1)  # deterministic summation order on CPU
    build_case(ref, "toy_edgecases", toy_batch(5, 1), 5, 1, clustered=False)
    build_case(ref, "toy_edgecases_fe3", toy_batch(7, 3, seed=11), 7, 3, clustered=False)
    build_case(ref, "synthetic_small", synthetic_batch(clusters=True), 50, 1, clustered=True)
    build_case(ref, "fixture_1ATN", fixture_batch(), 50, 1, clustered=True)
    build_case(
        ref,
        "fixture_variants_fe5",
        fixture_batch("variants.hdf5", edge_features=("distance", "same_chain", "covalent", "electrostatic", "vanderwaals"), target="binary"),
        50,
        5,
        clustered=True,
    )


if __name__ == "__main__":
    main()
