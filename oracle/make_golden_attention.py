"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/attention_segment_softmax.npz``.

Run in the build container only (needs ``/root/reference``):

    python -m oracle.make_golden_attention

The opt-in segment-softmax attention has no reference output to record: the reference normalises its attention logit over a
singleton axis (``ginet.py:54``).  What CAN be taken from the reference is everything up to that line, and this script does:
the logit ``leaky_relu(fc_attention(cat[fc(x)[row], fc(x)[col], fc_edge_attr(edge_attr)]))`` is computed BY THE REFERENCE'S OWN
MODULE (the sub-modules of an unmodified ``GINetConvLayer`` executed in the order of ``ginet.py:45-52``).  The normalisation over
the edges of each destination and the weighted sum are then formed in float64 by a plain loop written from the formula, and the
gradients by float64 autograd of the same formula.  Key scheme as in ``oracle/make_golden.py``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch.nn.functional import leaky_relu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "attention_segment_softmax.npz")


def graph(n, e, gen):
    row = torch.randint(0, n - 3, (e,), generator=gen)  # the last three nodes never receive an edge
    col = torch.randint(0, n, (e,), generator=gen)
    row[1], col[1] = row[0], col[0]  # duplicate edge
    col[2] = row[2]  # self loop
    row[e // 2 : e // 2 + 70] = 1  # a hub destination
    return torch.stack([row, col])


def segment_softmax_float64(logit, row, proj, n):
    z = torch.zeros(n, proj.shape[1], dtype=torch.float64)
    for i in range(n):
        edges = (row == i).nonzero().reshape(-1)
        if edges.numel() == 0:
            continue
        lg = logit[edges]
        ex = torch.exp(lg - lg.max())
        alpha = ex / ex.sum()
        z[i] = (alpha.unsqueeze(1) * proj[edges]).sum(0)
    return z


def main():
    ref = load_reference()
    store = {}
    for case, (fi, fo, fe, n, e, seed) in {"a": (10, 16, 2, 60, 700, 3), "b": (50, 32, 1, 90, 1500, 4)}.items():
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        layer = ref.ginet.GINetConvLayer(fi, fo, fe)  # the reference's module: its parameters, its initialisation
        with torch.no_grad():
            layer.fc_attention.weight.mul_(3.0)  # spread the logits
        ei = graph(n, e, gen)
        x = torch.randn(n, fi, generator=gen)
        ea = torch.rand(e, fe, generator=gen) * 8.0
        gout = torch.randn(n, fo, generator=gen)
        row, col = ei
        with torch.no_grad():  # ginet.py:45-52, executed by the reference's sub-modules in fp32
            xcol = layer.fc(x[col])
            xrow = layer.fc(x[row])
            ed = layer.fc_edge_attr(ea)
            logit32 = leaky_relu(layer.fc_attention(torch.cat([xrow, xcol, ed], dim=1))).squeeze(1)
        # the same formula in float64 with autograd, for z and every gradient
        w = {k: v.detach().double().requires_grad_(True) for k, v in layer.state_dict().items()}
        xd = x.double().requires_grad_(True)
        proj = xd @ w["fc.weight"].T
        logit = leaky_relu(torch.cat([proj[row], proj[col], ea.double() @ w["fc_edge_attr.weight"].T], dim=1) @ w["fc_attention.weight"].T).squeeze(1)
        assert torch.allclose(logit.detach().float(), logit32, rtol=1e-5, atol=1e-5), "float64 formula and the reference's fp32 modules disagree"
        mx = torch.full((n,), -float("inf"), dtype=torch.float64).scatter_reduce(0, row, logit.detach(), reduce="amax")
        ex = torch.exp(logit - mx[row])
        den = torch.zeros(n, dtype=torch.float64).index_add(0, row, ex)
        z = torch.zeros(n, fo, dtype=torch.float64).index_add(0, row, (ex / den[row]).unsqueeze(1) * proj[col])
        z_loop = segment_softmax_float64(logit.detach(), row, proj.detach()[col], n)
        assert torch.allclose(z.detach(), z_loop, rtol=1e-12, atol=1e-12)
        (z * gout.double()).sum().backward()
        p = f"{case}/"
        store[p + "in/x"], store[p + "in/edge_index"], store[p + "in/edge_attr"] = x.numpy(), ei.numpy(), ea.numpy()
        for k, v in layer.state_dict().items():
            store[p + "w/" + k] = v.detach().numpy()
        store[p + "out/logit_reference_fp32"] = logit32.numpy()
        store[p + "out/z"] = z.detach().float().numpy()
        store[p + "gout/z"] = gout.numpy()
        store[p + "grad/x"] = xd.grad.float().numpy()
        for k, v in w.items():
            store[p + "grad/" + k] = v.grad.float().numpy()
    np.savez_compressed(GOLDEN, **store)
    print("wrote", GOLDEN, f"{os.path.getsize(GOLDEN) / 1e3:.0f} kB", sorted(store)[:6], "...")


if __name__ == "__main__":
    main()
