"""drk_spmm vs drk_spmm_tiled on the C3 adjacency (64 atom-level graphs), widths 16/32, L2 flushed and L2 warm.
    gpurun -- python profiles/spmm_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200 import ops

ops.SPMM_TILED = True  # the tiled kernel is opt-in
from deeprank2_b200.graph import graph_index
from deeprank2_b200.synthetic import ATOM, make_batch

dev = torch.device("cuda", 0)
host = make_batch(64, n_node_features=38, n_edge_features=1, level=ATOM)
b = host.clone().to(dev)
gi = graph_index(b)
n, e = b.num_nodes, b.num_edges
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"N {n} E {e} max graph nodes {gi.max_graph_nodes}")


def t_us(fn, reps=20, cold=True):
    tot = 0.0
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(c)
    return 1e3 * tot / reps


for width in (16, 32, 64):
    src = torch.randn(n, width, device=dev)
    out = torch.empty_like(src)
    by = 8 * n * width + 4 * e + 4 * (n + 1)
    for name, kw in (("generic", {}), ("tiled", {"graph": gi})):
        f = lambda: ops.spmm(gi.rowptr, gi.colidx, src, n, act=ops.ACT_RELU, out=out, **kw)  # noqa: E731
        f()
        cold, warm = t_us(f, cold=True), t_us(f, cold=False)
        print(f"width {width:2d} {name:8s}: cold {cold:7.1f} us ({by / cold / 1e3 / 6544.3:5.3f} of HBM peak)   warm {warm:7.1f} us")
