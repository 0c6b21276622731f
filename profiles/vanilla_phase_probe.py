"""SM clocks per phase of k_vanilla_fwd (probe build: -DDRK_VANILLA_PROBE, see profiles/vanilla_phase_probe.sh).
Thread 0 of every CTA accumulates clock64() deltas per phase over all its graphs; this script runs VanillaNetwork forward passes
on the C2 batch and prints the per-CTA mean / max of every phase."""
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
buf = np.zeros((148, 8), dtype=np.int64)
batch = make_batch(256).to("cuda")
net = VanillaNetwork(50, 1, 1).to("cuda").eval()
with torch.no_grad():
    for _ in range(3):
        net(batch)
    read(buf.ctypes.data)
    reps = 10
    for _ in range(reps):
        net(batch)
    read(buf.ctypes.data)
per_launch = buf / (2.0 * reps)  # two layers per forward pass
names = ["graph setup", "pass 0 (V, all tiles)", "prefetch + stage wait", "U GEMM + barrier", "edge pass (warp 0)", "barrier after edges", "out GEMM + barrier", "-"]
tot = per_launch.sum(1)
print(f"cycles per CTA and launch: mean {tot.mean():.0f}  max {tot.max():.0f}  min {tot.min():.0f}")
for i, nme in enumerate(names[:7]):
    col = per_launch[:, i]
    print(f"  {nme:26s} mean {col.mean():9.0f} ({100 * col.mean() / tot.mean():5.1f} %)   max {col.max():9.0f}")
