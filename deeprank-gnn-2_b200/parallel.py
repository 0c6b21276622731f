"""Data parallelism over graph mini-batches, one process per GPU: the NCCL gradient all-reduce of the AUTOGRAD path (every network
but the benchmark model, and the fallback of the fused step).  The benchmark model's step does not come through here: its gradient
exchange is fused into the finalize kernel over NVLink peer memory (``fused.GINetFusedStep._peer_exchange``, ``k_step_finalize``).

The reference's only multi-GPU mechanism is ``nn.DataParallel`` (``deeprank2/trainer.py:387-389``), a
single-process replicate/scatter/gather that cannot split a PyG ``Batch``.  Graphs are independent (a
batch is a block-diagonal union), so the path shards by graph: every rank builds its own batches and
graph index, weights are replicated, and the only exchange is ONE all-reduce of the flat fp32
gradient (11 273 parameters = 45 KB for GINet at F_in = 50) per step over NVLink/NVSwitch.  The
message is latency bound, so on this path everything is packed into a single NCCL call.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradAllReduce:
    """Averages ``p.grad`` over ranks with one flat all-reduce (call between backward and optimizer.step).

    With per-rank mean losses over equal local batch sizes, the average of the local gradients is the
    gradient of the global mean loss; ``weights`` lets the caller pass ``B_local / B_global`` for ragged
    last batches (the sum is then not divided by the world size).
    """

    def __init__(self, model: torch.nn.Module, world_size: int | None = None, group=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.group = group
        self.world = world_size if world_size is not None else dist.get_world_size(group)
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def __call__(self, local_weight: float | None = None):
        if self.world == 1:
            return
        p0 = self.params[0]
        if self._flat is None or self._flat.device != p0.device:
            self._flat = torch.empty(self.numel, dtype=torch.float32, device=p0.device)
        flat = self._flat
        off = 0
        views = []
        for p in self.params:
            n = p.numel()
            v = flat[off : off + n].view_as(p)
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
            views.append(v)
            off += n
        if local_weight is not None:
            flat.mul_(local_weight)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        if local_weight is None:
            flat.div_(self.world)
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s weights (replaces DataParallel's per-step replicate)."""
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_indices(n_items: int, rank: int, world: int) -> range:
    """Contiguous, near-equal split of ``n_items`` graphs over ``world`` ranks (rank r gets the r-th slice)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
