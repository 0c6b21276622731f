"""Micro-driver for profiling single kernels on the C2 batch (used under ncu).
usage: python profiles/micro_kernels.py [spmm|linear|wgrad|index|all] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200 import ops
from deeprank2_b200.graph import GraphIndex
from deeprank2_b200.synthetic import make_batch

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda")
b = make_batch(256).to(dev)
n = b.num_nodes
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, bytes_):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()  # > L2; no sync: the kernel is enqueued while the flush runs, so the interval has no launch latency
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:40s} {t:8.1f} us   {bytes_ / t / 1e3:8.1f} GB/s algorithmic ({bytes_ / 1e6:.1f} MB)", flush=True)


g = GraphIndex.build(b.edge_index, n, batch=b.batch, num_graphs=256)
e = g.num_edges
if what in ("index", "all"):
    timed("graph_index_build (CSR+CSC+offsets)", lambda: GraphIndex.build(b.edge_index, n, batch=b.batch, num_graphs=256), 16 * e + 16 * e + 8 * (n + 1))
    blocks = (b._node_ptr32, b._edge_ptr32, b.meta("max_graph_nodes"), b.meta("max_graph_edges"))
    timed("graph_index_build_blocked (same output)", lambda: GraphIndex.build(b.edge_index, n, batch=b.batch, num_graphs=256, blocks=blocks), 16 * e + 16 * e + 8 * (n + 1))
if what in ("spmm", "all"):
    for w in (16, 32, 64):
        src = torch.randn(n, w, device=dev)
        out = torch.empty_like(src)
        timed(f"spmm width {w}", lambda: ops.spmm(g.rowptr, g.colidx, src, n, act=ops.ACT_RELU, out=out), 8 * n * w + 4 * e + 4 * (n + 1))
if what in ("linear", "all"):
    for k, m in ((50, 32), (16, 32), (32, 16), (64, 32)):
        a = torch.randn(n, k, device=dev)
        wgt = torch.randn(m, k, device=dev)
        out = torch.empty(n, m, device=dev)
        timed(f"node_linear K={k} M={m}", lambda: ops.node_linear(a, wgt, True, out=out), 4 * n * (k + m))
if what in ("wgrad", "all"):
    for k, m in ((50, 32), (16, 32)):
        x = torch.randn(n, k, device=dev)
        dy = torch.randn(n, m, device=dev)
        timed(f"weight_grad K={k} M={m}", lambda: ops.weight_grad(dy, x), 4 * n * (k + m))
if what in ("mean", "all"):
    x = torch.randn(n, 64, device=dev)
    timed("segment_mean width 64", lambda: ops.segment_mean(x, g.graph_ptr, 256), 4 * n * 64)
    dg = torch.randn(256, 64, device=dev)
    timed("segment_mean_bwd width 64 + mask", lambda: ops.segment_mean_bwd(dg, g.graph_ptr, g.batch32, n, mask=x), 8 * n * 64)
