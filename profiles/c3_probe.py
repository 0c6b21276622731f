"""Config C3: GINet inference on synthetic atom-level graphs (~3 k nodes, ~60 k directed edges each, 38 node features) -- the
graphs that do not fit one CTA's shared memory and therefore run on the layer kernels -- and the aggregation kernel (drk_spmm)
on its own against the measured HBM peak.
usage: python profiles/c3_probe.py [graphs=64] [reps=20]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200 import _lib, ops
from deeprank2_b200.graph import GraphIndex
from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
from deeprank2_b200.synthetic import ATOM, make_batch

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6544.3
host = make_batch(graphs, n_node_features=38, n_edge_features=1, level=ATOM)
batches = [host.clone().to(dev) for _ in range(3)]
n, e = host.num_nodes, host.num_edges
print(f"C3 batch: {graphs} graphs, {n} nodes, {e} directed edges (degree {e / n:.1f}); HBM peak {peak:.0f} GB/s", flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)
net = GINet(38, 1, 1).to(dev).eval()


def event_time(fn, reps, flush_l2=True):
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) * 1e3)
    return sorted(ts)[len(ts) // 2]


with torch.no_grad():
    for b in batches:
        net(b)
    torch.cuda.synchronize()
    c0 = _lib.launch_count()
    net(batches[0])
    launches = _lib.launch_count() - c0
    i = [0]

    def infer():
        i[0] += 1
        return net(batches[i[0] % len(batches)])

    t = event_time(infer, reps)
    print(f"(index cached on the batch) GINet inference, eager, L2 flushed: {t:8.1f} us per batch  {graphs / t * 1e6:9.0f} graphs/s  {e / t * 1e-3:7.2f} G edges/s  ({launches} launches of ours)", flush=True)

    def infer_fresh():  # what a loader delivers: a batch nobody has indexed yet
        i[0] += 1
        bt = batches[i[0] % len(batches)]
        bt.__dict__.pop("_graph_index", None)
        return net(bt)

    t = event_time(infer_fresh, reps)
    print(f"GINet inference incl. index build (CSR only), eager, L2 flushed: {t:8.1f} us per batch  {graphs / t * 1e6:9.0f} graphs/s  {e / t * 1e-3:7.2f} G edges/s", flush=True)
    # the same forward replayed from a CUDA graph (no launch gaps)
    static = batches[0]
    static.__dict__.pop("_graph_index", None)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        net(static)
    torch.cuda.current_stream().wait_stream(s)
    static.__dict__.pop("_graph_index", None)
    with torch.cuda.graph(g):
        out = net(static)  # index build + forward
    t = event_time(g.replay, reps)
    fwd_bytes = 4 * n * 38 + 16 * e + 980 * n  # x + int64 contacts + the layer path's forward intermediates (F_in=38 -> close)
    print(f"GINet inference incl. index build, graph replay, L2 flushed: {t:8.1f} us per batch  {graphs / t * 1e6:9.0f} graphs/s  {e / t * 1e-3:7.2f} G edges/s", flush=True)

gi = GraphIndex.build(static.edge_index, n, batch=static.batch, num_graphs=graphs)
for w in (16, 32, 64):
    src = torch.randn(n, w, device=dev)
    out = torch.empty_like(src)
    by = 8 * n * w + 4 * e + 4 * (n + 1)
    t = event_time(lambda: ops.spmm(gi.rowptr, gi.colidx, src, n, act=ops.ACT_RELU, out=out), reps)
    print(f"drk_spmm width {w:2d} (C3 adjacency): {t:7.1f} us  algorithmic {by / 1e6:6.1f} MB -> {by / t / 1e3:7.1f} GB/s = {by / t / 1e3 / peak:5.3f} of peak;"
          f"  gathered {(4 * e * w + by) / t / 1e3:7.1f} GB/s", flush=True)
blocks = (static._node_ptr32, static._edge_ptr32, static.meta("max_graph_nodes"), static.meta("max_graph_edges"))
by = 16 * e + 24 * e + 8 * (n + 1)
try:
    t = event_time(lambda: GraphIndex.build(static.edge_index, n, batch=static.batch, num_graphs=graphs, blocks=blocks), reps)
    print(f"graph index build (blocked): {t:7.1f} us  {by / t / 1e3:7.1f} GB/s = {by / t / 1e3 / peak:5.3f} of peak", flush=True)
except Exception as exc:  # noqa: BLE001
    print("blocked index build not applicable:", str(exc)[:100])
t = event_time(lambda: GraphIndex.build(static.edge_index, n, batch=static.batch, num_graphs=graphs), reps)
print(f"graph index build (general): {t:7.1f} us  {by / t / 1e3:7.1f} GB/s = {by / t / 1e3 / peak:5.3f} of peak", flush=True)
by = 16 * e + 12 * e + 4 * (n + 1)
t = event_time(lambda: GraphIndex.build(static.edge_index, n, batch=static.batch, num_graphs=graphs, with_csc=False), reps)
print(f"graph index build (general, CSR only): {t:7.1f} us  {by / t / 1e3:7.1f} GB/s = {by / t / 1e3 / peak:5.3f} of peak", flush=True)
