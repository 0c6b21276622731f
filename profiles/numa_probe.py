"""Host topology of the GPU box and pinned H2D bandwidth as a function of the CPU set the process (and its pinned allocation) runs on.
usage: python profiles/numa_probe.py"""
import glob
import os
import subprocess
import sys
import time

import torch

print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^cpu\\(s\\)'; nvidia-smi topo -m | head -8; nproc", shell=True, capture_output=True, text=True).stdout)
dev = torch.device("cuda", 0)
props = torch.cuda.get_device_properties(0)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0" if hasattr(props, "pci_bus_id") else None
node = None
if bus:
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except OSError as exc:
        print("numa_node unreadable:", exc)
print("gpu pci", bus, "numa node", node, "affinity now", len(os.sched_getaffinity(0)), "cpus")
nodes = {}
for p in glob.glob("/sys/devices/system/node/node*/cpulist"):
    nid = int(p.split("node")[-1].split("/")[0])
    cpus = set()
    for part in open(p).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus |= set(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    nodes[nid] = cpus
print({k: (min(v), max(v), len(v)) for k, v in nodes.items() if v})
allowed = os.sched_getaffinity(0)


def h2d_rate(nbytes=64 << 20, reps=20):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


print(f"H2D pinned, default affinity: {h2d_rate():.1f} GB/s")
for nid, cpus in sorted(nodes.items()):
    use = cpus & allowed
    if not use:
        continue
    os.sched_setaffinity(0, use)
    print(f"H2D pinned, process on node {nid} ({len(use)} cpus): {h2d_rate():.1f} GB/s", flush=True)
os.sched_setaffinity(0, allowed)
