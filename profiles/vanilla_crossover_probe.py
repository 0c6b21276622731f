"""Per-graph fused Vanilla layer kernels vs the batch-level kernels as a function of the number of graphs in the batch
(one CTA per graph: a batch of few graphs leaves most SMs idle).  Train step of VanillaNetwork, CUDA-graph replay."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank2_b200 import ops
from deeprank2_b200.neuralnets.gnn.vanilla_gnn import VanillaNetwork
from deeprank2_b200.step import GraphedTrainStep, TrainStep
from deeprank2_b200.synthetic import make_batch

for g in (4, 16, 32, 64, 96, 128, 192, 256):
    host = make_batch(g)
    res = []
    for fused in (True, False):
        ops.VANILLA_FUSED = fused
        if hasattr(ops, "VANILLA_FUSED_MIN_GRAPHS"):
            ops.VANILLA_FUSED_MIN_GRAPHS = 1
        batch = host.clone().to("cuda")
        torch.manual_seed(0)
        net = VanillaNetwork(50, 1, 1).to("cuda").train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
        inner = TrainStep(net, opt, torch.nn.MSELoss())
        for _ in range(3):
            inner(batch)
        gs = GraphedTrainStep(inner, batch, warmup=1)
        for _ in range(3):
            gs.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30):
            gs.replay()
        b.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 30)
    print(f"graphs {g:4d}: fused {res[0]:.3f} ms   batch-level {res[1]:.3f} ms", flush=True)
