import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank2_b200.neuralnets.gnn import sgat
from deeprank2_b200.synthetic import make_batch
from deeprank2_b200.utils import community_pooling as cp
import copy
host = make_batch(256, with_clusters=True)
b = host.clone().to("cuda")
net = sgat.SGAT(50, 1, 1).to("cuda").train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
def step():
    v = copy.copy(b); v.__dict__ = dict(b.__dict__)
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(net(v).reshape(-1), b.y)
    loss.backward(); opt.step()
    return loss
for env in ("1", "0"):
    cp.POOL_BLOCKED = env == "1"
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): step()
    torch.cuda.synchronize()
    print("POOL_BLOCKED", env, (time.perf_counter() - t0) / 10 * 1e3, "ms/step", flush=True)
cp.POOL_BLOCKED = True
torch.cuda.set_sync_debug_mode("warn")
step()
torch.cuda.set_sync_debug_mode("default")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12))
