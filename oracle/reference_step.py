"""TEST INFRASTRUCTURE ONLY -- the reference's own train step, for ``bench.py``'s reference legs.

``reference_train_steps`` builds the UNMODIFIED ``deeprank2.neuralnets.gnn.ginet_nocluster.GINet`` (or ``VanillaNetwork`` /
``FoutNet`` / clustered ``GINet``) through ``oracle/reference_loader.py`` (from ``/root/reference`` in the build container,
from the installed copy ``oracle/_ref`` on the GPU box) and runs the loop body of ``Trainer._epoch`` (reference
``deeprank2/trainer.py:682-694``): ``zero_grad -> pred = model(batch) -> loss = MSELoss(pred.reshape(-1), y) -> backward ->
Adam(lr 1e-3, weight_decay 1e-5).step()`` -- on the CPU with all host threads (``cpu_baseline`` / ``--impl reference``) or on
a CUDA device with stock eager ATen kernels (``gpu_torch_baseline``).  The third-party primitives under the reference
modules are the pure-torch restatements of ``oracle/thirdparty.py`` (the wheels are absent from this image).

When the reference files are not available at all the same loop runs on the oracle port (``oracle/restate.py``) and the
result says ``kind: "port"``.
"""
from __future__ import annotations

import os
import time

import torch

NETS = {"ginet": ("ginet_nocluster", "GINet"), "vanilla": ("vanilla_gnn", "VanillaNetwork"), "fout": ("foutnet", "FoutNet"), "ginet_clustered": ("ginet", "GINet")}


def to_reference_batch(batch):
    """A ``deeprank2_b200`` host ``Batch`` as the attribute bag the reference modules take (``oracle.thirdparty.Batch``)."""
    from oracle import thirdparty as tp

    fields = {}
    for k in ("x", "edge_index", "edge_attr", "y", "pos", "cluster0", "cluster1", "ptr"):
        v = batch.__dict__.get(k)
        if v is None and k in ("x", "edge_index", "edge_attr", "y", "pos"):
            v = getattr(batch, k, None)
        if isinstance(v, torch.Tensor):
            fields[k] = v
    out = tp.Batch(batch=batch.batch, **{k: v for k, v in fields.items() if k != "ptr"})
    if "ptr" in fields:
        out.ptr = fields["ptr"]
    return out


def build_reference(net: str, f_node: int, f_edge: int, seed: int = 0):
    """(module, kind): the reference network with ``torch.manual_seed(seed)`` weights, or (None, "port")."""
    from oracle import reference_loader as rl

    if not rl.reference_available():
        return None, "port"
    ref = rl.load_reference()
    mod_name, cls_name = NETS[net]
    torch.manual_seed(seed)
    return getattr(getattr(ref, mod_name), cls_name)(f_node, 1, f_edge), "reference"


def reference_train_steps(batch, net: str = "ginet", steps: int = 5, warmup: int = 1, device: str = "cpu", budget_s: float | None = None,
                          f_node: int | None = None, f_edge: int | None = None, train: bool = True):
    """Time ``steps`` train (or ``train=False``: inference) steps of the reference on ``batch`` (a host ``Batch`` of this repo)."""
    f_node = int(batch.x.shape[1]) if f_node is None else f_node
    f_edge = int(batch.edge_attr.shape[1]) if f_edge is None else f_edge
    dev = torch.device(device)
    if dev.type == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    module, kind = build_reference(net, f_node, f_edge)
    if module is None:
        if net != "ginet" or dev.type != "cpu":
            return None
        return _port_steps(batch, steps, warmup, budget_s)
    module = module.to(dev)
    module.train(train)
    data = to_reference_batch(batch).to(dev)
    opt = torch.optim.Adam(module.parameters(), lr=1e-3, weight_decay=1e-5)  # trainer.py:404-419 defaults
    loss_fn = torch.nn.MSELoss()

    def one():
        if not train:
            with torch.no_grad():
                return module(data.clone())
        opt.zero_grad()
        pred = module(data.clone())  # the nets overwrite data.x / pool the batch in place: a loader hands out a fresh batch
        loss = loss_fn(pred.reshape(-1), data.y)
        loss.backward()
        opt.step()
        return loss

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    t0 = time.perf_counter()
    for _ in range(warmup):
        one()
    sync()
    if warmup > 0 and budget_s is not None:
        per_step = (time.perf_counter() - t0) / warmup
        steps = max(2, min(steps, int(budget_s / max(per_step, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    sync()
    dt = time.perf_counter() - t0
    return dict(seconds=dt, steps=steps, graphs=int(batch.num_graphs), nodes=int(batch.num_nodes), edges=int(batch.num_edges),
                threads=torch.get_num_threads() if dev.type == "cpu" else 0, kind=kind, device=str(dev), net=net)


def _port_steps(batch, steps, warmup, budget_s):
    from oracle import restate as R

    torch.manual_seed(0)
    params = R.as_parameters(R.ginet_nocluster_init(int(batch.x.shape[1]), 1, int(batch.edge_attr.shape[1])))
    opt = R.make_adam(params)
    t0 = time.perf_counter()
    for _ in range(warmup):
        R.train_step(R.ginet_nocluster_forward, params, opt, batch, training=True)
    if warmup > 0 and budget_s is not None:
        per_step = (time.perf_counter() - t0) / warmup
        steps = max(2, min(steps, int(budget_s / max(per_step, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        R.train_step(R.ginet_nocluster_forward, params, opt, batch, training=True)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, steps=steps, graphs=int(batch.num_graphs), nodes=int(batch.num_nodes), edges=int(batch.num_edges),
                threads=torch.get_num_threads(), kind="port", device="cpu", net="ginet")
