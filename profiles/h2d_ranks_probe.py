"""Aggregate pinned host->device bandwidth of N concurrent ranks on one box (the wall of the host-fed `e2e` loop at N GPUs): every rank
copies the byte count of one C2 step (18.6 MB, four tensors) back to back from its own pinned buffers, all ranks at once.

    torchrun --nproc-per-node N profiles/h2d_ranks_probe.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sizes = [15_493_800, 3_133_620, 1_024, 2_056]  # x, packed contacts, targets, offsets of one C2 batch
host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
devb = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
stream = torch.cuda.Stream()


def run(steps):
    with torch.cuda.stream(stream):
        for _ in range(steps):
            for h, d in zip(host, devb):
                d.copy_(h, non_blocking=True)
    stream.synchronize()


run(20)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
steps = 400
run(steps)
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    per_step = sum(sizes)
    agg = per_step * steps * world / float(t.item()) / 1e9
    print(f"{world} rank(s): {agg:.1f} GB/s aggregate pinned H2D ({agg / world:.1f} GB/s per rank) = {steps * world / float(t.item()) * 256:.0f} graphs/s worth of C2 batches; "
          f"cpus {os.cpu_count()}, affinity {len(os.sched_getaffinity(0))}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
