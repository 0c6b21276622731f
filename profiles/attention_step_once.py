"""Two eager GINet(attention="segment_softmax") train steps on the C2 batch (layer kernels) -- the command profiled for the launch list
profiles/r01_launches_attention_step.csv (ncu --metrics gpu__time_duration.sum)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
from deeprank2_b200.step import TrainStep
from deeprank2_b200.synthetic import make_batch

dev = torch.device("cuda", 0)
b = make_batch(256).to(dev)
torch.manual_seed(0)
net = GINet(50, 1, 1, attention="segment_softmax").to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True, fused=True)
step = TrainStep(net, opt, torch.nn.MSELoss())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    loss, _ = step(b)
torch.cuda.synchronize()
print("loss", float(loss))
