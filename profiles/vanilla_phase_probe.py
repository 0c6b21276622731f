"""SM clocks per phase of k_vanilla_fwd / k_vanilla_bwd (probe build: -DDRK_VANILLA_PROBE, see profiles/vanilla_phase_probe.sh).
Thread 0 of every CTA accumulates clock64() deltas per phase over all its graphs; this script runs VanillaNetwork train steps
(forward + backward, no optimizer) on the C2 batch and prints the per-CTA mean / max of every phase, per launch."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deeprank2_b200 import _lib
from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaNetwork
from deeprank2_b200.synthetic import make_batch

lib = _lib.load()
read = lib.drk_vanilla_probe_read
read.argtypes = [ctypes.c_void_p]
buf = np.zeros((148, 16), dtype=np.int64)
batch = make_batch(256).to("cuda")
net = VanillaNetwork(50, 1, 1).to("cuda").train()


def step():
    net.zero_grad()
    torch.nn.functional.mse_loss(net(batch).reshape(-1), batch.y).backward()


for _ in range(3):
    step()
read(buf.ctypes.data)
reps = 10
for _ in range(reps):
    step()
read(buf.ctypes.data)
per_launch = buf / (2.0 * reps)  # two layers per step
fwd = ["graph setup", "pass 0 (V, all tiles)", "prefetch + stage wait", "U product + barrier", "edge pass (warp 0)", "barrier after edges", "node-MLP product + barrier", "-"]
bwd = ["gradient partial of the graph", "graph setup", "pass 0 (dS product)", "pass 1: prefetch, wait, dZ", "walks dU / dV (warp 0)", "barrier after walks", "dx product", "dWn / dWab products + barrier"]
for title, names, lo in (("k_vanilla_fwd", fwd, 0), ("k_vanilla_bwd", bwd, 8)):
    part = per_launch[:, lo : lo + 8]
    tot = part.sum(1)
    print(f"{title}: cycles per CTA and launch: mean {tot.mean():.0f}  max {tot.max():.0f}  min {tot.min():.0f}")
    for i, nme in enumerate(names):
        if nme == "-":
            continue
        col = part[:, i]
        print(f"  {nme:30s} mean {col.mean():9.0f} ({100 * col.mean() / tot.mean():5.1f} %)   max {col.max():9.0f}")
