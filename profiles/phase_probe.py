"""Where the cycles of k_ginet_step go: SM clock at the phase boundaries of every graph (drk_ginet_step_set_phase_clocks),
C2 batches (256 graphs).  Prints mean cycles per phase per graph and the CTAs' spans in cycles.

    gpurun -- python profiles/phase_probe.py > gpurun_out/phase_probe.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200 import _lib
from deeprank2_b200.fused import GINetFusedStep
from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
from deeprank2_b200.synthetic import make_batch

NAMES = ["index", "x stage", "project", "agg H1", "agg A2", "conv2+readout", "head+loss", "dW2+dA2", "agg dZ1", "agg Q", "dW1", "dW1 reduce"]


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    batches = [make_batch(256, first=b * 256, n_node_features=50, n_edge_features=1).to(dev) for b in range(3)]
    torch.manual_seed(0)
    model = GINet(50, 1, 1).to(dev).train(os.environ.get("DRK_PROBE_EVAL") is None)
    step = GINetFusedStep(model, torch.optim.SGD(model.parameters(), lr=0.0), torch.nn.MSELoss())
    for b in batches:
        step.forward_backward(b)
    torch.cuda.synchronize()
    clk = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
    lib.drk_ginet_step_set_phase_clocks(clk.data_ptr(), 256)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    acc = torch.zeros(256, 12, dtype=torch.float64)
    span = []
    reps = 6
    for r in range(reps):
        if os.environ.get("DRK_PROBE_NOFLUSH") is None:
            flush.zero_()  # inputs AND the kernel's code come from HBM; without the flush three 19 MB batches rotate inside L2
        step.forward_backward(batches[r % 3])
        torch.cuda.synchronize()
        c = clk.view(256, 16).cpu()
        acc += (c[:, 1:13] - c[:, 0:12]).double()
        # slots s and s + 148 ran on the same CTA: the CTA's span is first start -> last end (clocks are per SM, only compare within a CTA)
        per_cta = []
        for s in range(148):
            end = c[s + 148, 12] if s + 148 < 256 else c[s, 12]
            per_cta.append(int(end - c[s, 0]))
        span.append((max(per_cta), sum(per_cta) / len(per_cta), min(per_cta)))
    lib.drk_ginet_step_set_phase_clocks(None, 0)
    acc /= reps
    mean = acc.mean(0)
    total = float(mean.sum())
    print(f"mean cycles per graph: {total:.0f}")
    for i, nm in enumerate(NAMES):
        print(f"  {nm:16s} {float(mean[i]):9.0f}  {100 * float(mean[i]) / total:5.1f} %")
    c = clk.view(256, 16).cpu().double()
    sub = [("  project: warp 0's two tiles", c[:, 13] - c[:, 2]), ("  project: barrier", c[:, 14] - c[:, 13]), ("  project: L2 prefetch of the next graph", c[:, 3] - c[:, 14])]
    for nm, d in sub:
        print(f"{nm:24s} {float(d.mean()):9.0f}   (first-round graphs {float(d[:148].mean()):9.0f}, second-round {float(d[148:].mean()):9.0f})")
    print("CTA span cycles (max, mean, min) per rep:", span)
    ptr = batches[0]._node_ptr32.cpu()
    n = ptr[1:] - ptr[:-1]
    print("graph sizes: nodes min/mean/max", int(n.min()), float(n.float().mean()), int(n.max()))


if __name__ == "__main__":
    main()
