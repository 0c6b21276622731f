"""One training step of ``Trainer._epoch`` (reference ``deeprank2/trainer.py:682-694``) as a unit that can be
run eagerly or captured once into a CUDA graph and replayed.

A step is: (collate-side) build the batch's graph index on the device -> ``zero_grad`` -> ``model(batch)``
-> ``loss`` -> ``backward`` -> ``optimizer.step``.  The reference synchronises three times per step
(``loss.item()``, ``y.cpu()``, ``pred.cpu()``, ``trainer.py:694-703``); here the loss and predictions stay
on the device and are read back when the caller asks (end of epoch), so a captured step has no host
round trip at all -- at ~60 us of device work per step that is the difference between being
launch/sync bound and bandwidth bound.
"""
from __future__ import annotations

import torch


class TrainStep:
    """``step = TrainStep(model, optimizer, loss_fn); loss, pred = step(batch)`` (eager)."""

    def __init__(self, model, optimizer, loss_fn, format_output=None, rebuild_index: bool = True):
        self.model = model
        self.optimizer = optimizer
        self.loss_fn = loss_fn
        self.format_output = format_output or (lambda pred, y: (pred.reshape(-1), y))
        self.rebuild_index = rebuild_index

    def __call__(self, batch):
        if self.rebuild_index:
            batch.__dict__.pop("_graph_index", None)  # the index is part of every step, like PyG's collate
        self.optimizer.zero_grad(set_to_none=True)
        pred = self.model(batch)
        out, target = self.format_output(pred, batch.y)
        loss = self.loss_fn(out, target)
        loss.backward()
        self.optimizer.step()
        return loss.detach(), pred.detach()


class GraphedTrainStep:
    """The same step captured into a CUDA graph for ONE device-resident, pre-collated batch.

    ``replay()`` re-runs index build + forward + backward + optimizer on the batch's static tensors; new
    input values can be copied into ``batch.x`` etc. beforehand (shapes are fixed by the capture).
    Several batches share one memory pool (``pool=``) because their replays never overlap.
    """

    def __init__(self, step: TrainStep, batch, pool=None, warmup: int = 2):
        if not batch.x.is_cuda:
            raise RuntimeError("GraphedTrainStep needs a device-resident batch")
        self.step = step
        self.batch = batch
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step(batch)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.loss, self.pred = step(batch)
        self.pool = self.graph.pool()

    def replay(self):
        self.graph.replay()
        return self.loss, self.pred
