"""Segment-softmax attention kernels (csrc/drk_attention.cu) on the C2 batch: per-kernel event timing after an L2 flush against the
measured HBM peak, and the GINet(attention="segment_softmax") train step (layer path) eager and replayed from a CUDA graph.
usage: python profiles/attention_probe.py [reps=20] [width=16]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200 import _lib, ops
from deeprank2_b200.graph import graph_index, stream_ptr
from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
from deeprank2_b200.step import GraphedTrainStep, TrainStep
from deeprank2_b200.synthetic import make_batch

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fo = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6544.3
b = make_batch(256).to(dev)
g = graph_index(b)
n, e, fe = b.num_nodes, b.num_edges, b.edge_attr.shape[1]
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"C2 batch: {n} nodes, {e} directed edges, width {fo}, {fe} edge feature(s); HBM peak {peak:.0f} GB/s", flush=True)


def timed(name, fn, nbytes):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} {t:7.1f} us   algorithmic {nbytes / 1e6:6.1f} MB -> {nbytes / t / 1e3:7.1f} GB/s = {nbytes / t / 1e3 / peak:5.3f} of peak   ({e / t * 1e-3:5.1f} G edges/s)", flush=True)
    return t


torch.manual_seed(0)
p = torch.randn(n, fo, device=dev)
s = torch.randn(n, 2, device=dev)
u = torch.randn(fe, device=dev)
att = torch.randn(2 * fo, device=dev)
z = torch.empty(n, fo, device=dev)
adq = torch.empty(e, 2, device=dev)
lgs = torch.empty(e, device=dev)
dy = torch.randn(n, fo, device=dev)
ds = torch.empty(n, 2, device=dev)
dz = torch.empty(n, fo, device=dev)
dp = torch.empty(n, fo, device=dev)
P = lambda t: t.data_ptr()  # noqa: E731
ea = g.attr_in_slot_order(b.edge_attr)
smap = g.slot_map()


def fwd():
    _lib.check(lib.drk_attn_fwd(P(g.rowptr), P(g.colidx), P(p), fo, P(s), P(ea), fe, fe, P(u), 0.01, P(z), fo, P(adq), P(lgs), n, fo, 1, stream_ptr()), "fwd")


def bwd_dst():
    _lib.check(lib.drk_attn_bwd_dst(P(g.rowptr), P(g.colidx), P(p), fo, P(dy), fo, P(z), fo, P(adq), 0.01, P(ds), P(dz), fo, n, fo, 1, stream_ptr()), "bwd_dst")


def bwd_src():
    _lib.check(lib.drk_attn_bwd_src(P(g.colptr), P(g.rowidx), P(smap), P(dz), fo, P(adq), P(ds), P(att), P(dp), fo, n, fo, stream_ptr()), "bwd_src")


def per_batch():
    g._slot_map = None
    g._attr_csr = None
    g.attr_in_slot_order(b.edge_attr)
    g.slot_map()


# algorithmic bytes: every distinct input byte read once, every output byte written once (SURVEY 8d's model)
timed("drk_attn_fwd", fwd, 8 * n * fo + 12 * n + (4 + 4 * fe + 4) * e)            # P, z, s, rowptr | colidx, attr, alpha
timed("drk_attn_bwd_dst", bwd_dst, 16 * n * fo + 8 * n + (4 + 4 + 4) * e)         # P, dy, y, dz, rowptr, ds | colidx, alpha, dq
timed("drk_attn_bwd_src", bwd_src, 8 * n * fo + 16 * n + (4 + 4 + 8) * e)         # dz, dp, colptr, ds | rowidx, slot map, (alpha, dq)
timed("per batch: attr gather + slot map", per_batch, (8 * fe + 16) * e)
timed("drk_spmm (alpha == 1) same width", lambda: ops.spmm(g.rowptr, g.colidx, p, n, act=ops.ACT_RELU, out=z), 8 * n * fo + 4 * e + 4 * (n + 1))

for mode in ("segment_softmax", "reference"):
    torch.manual_seed(0)
    net = GINet(50, 1, 1, attention=mode).to(dev).train()
    net.fused = False  # layer kernels in both modes
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
    inner = TrainStep(net, opt, torch.nn.MSELoss())
    for _ in range(3):
        inner(b)
    torch.cuda.synchronize()
    c0 = _lib.launch_count()
    inner(b)
    launches = _lib.launch_count() - c0
    gs = GraphedTrainStep(inner, b, warmup=1)
    for _ in range(3):
        gs.replay()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        gs.replay()
    c.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(c) / 50
    print(f"GINet train step, layer kernels, attention={mode:16s}: {t:6.3f} ms/step in graph replay  ({256 / t * 1e3:8.0f} graphs/s, {e / t * 1e-6:5.2f} G edges/s)  {launches} launches of ours", flush=True)
