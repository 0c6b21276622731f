"""Where the end-to-end step time goes: raw H2D rate of a pinned C2 batch, then the prefetch loop with and without the step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank2_b200.fused import GINetFusedStep
from deeprank2_b200.neuralnets.gnn.ginet_nocluster import GINet
from deeprank2_b200.pipeline import DevicePrefetcher, batch_nbytes
from deeprank2_b200.synthetic import make_batch

dev = torch.device("cuda", 0)
hbs = [make_batch(256, first=256 * b).pin_memory() for b in range(4)]
fields = GINetFusedStep.FIELDS
nb = batch_nbytes(hbs[0], fields)
x = hbs[0].x
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dbig = torch.empty_like(big, device=dev)
for _ in range(3):
    dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
print(f"raw pinned H2D: {5 * big.numel() / (time.perf_counter() - t0) / 1e9:.1f} GB/s")

def loop(n, work):
    feed = DevicePrefetcher((hbs[i % 4] for i in range(n)), dev, depth=2, only=fields)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in feed:
        work(b)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

torch.manual_seed(0)
model = GINet(50, 1, 1).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True, fused=True)
step = GINetFusedStep(model, opt, torch.nn.MSELoss())
for name, work in (("copy only", lambda b: None), ("copy + step", lambda b: step(b)), ("copy + step + item", lambda b: step(b)[0].item())):
    loop(6, work)
    dt = loop(40, work)
    print(f"{name}: {dt * 1e3:.3f} ms/step, {nb / dt / 1e9:.1f} GB/s H2D, {256 / dt:.0f} graphs/s")
# host-side cost of issuing one batch copy
t0 = time.perf_counter()
pf = DevicePrefetcher([], dev, only=fields)
for i in range(20):
    pf._issue(hbs[i % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host time to issue one batch copy: {(t1 - t0) / 20 * 1e3:.3f} ms")
