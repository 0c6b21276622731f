"""CPU: pins ``oracle/restate.py:ginet_conv_segment_softmax`` (the restatement of the intended attention operator -- the reference
never computes it, ``ginet.py:54``) against an independent float64 loop written from the formula, and checks the host-side
argument handling of the layer's ``attention`` switch."""
from __future__ import annotations

import math

import pytest
import torch

from oracle import restate as R


def _loop_float64(x, edge_index, edge_attr, p):
    w = p["fc.weight"].double()
    we = p["fc_edge_attr.weight"].double()
    a = p["fc_attention.weight"].double()[0]
    proj = x.double() @ w.T
    n, fo = proj.shape
    row, col = edge_index.tolist()
    out = torch.zeros(n, fo, dtype=torch.float64)
    for i in range(n):
        edges = [e for e in range(len(row)) if row[e] == i]
        if not edges:
            continue
        logits = []
        for e in edges:
            cat = torch.cat([proj[i], proj[col[e]], we @ edge_attr[e].double()])
            q = float(a @ cat)
            logits.append(q if q > 0 else 0.01 * q)
        m = max(logits)
        ex = [math.exp(v - m) for v in logits]
        den = sum(ex)
        for v, e in zip(ex, edges):
            out[i] += (v / den) * proj[col[e]]
    return out


@pytest.mark.parametrize("fe", [1, 3])
def test_segment_softmax_restatement_vs_float64_loop(fe):
    gen = torch.Generator().manual_seed(7 + fe)
    n, e = 9, 40
    ei = torch.stack([torch.randint(0, n - 2, (e,), generator=gen), torch.randint(0, n, (e,), generator=gen)])
    ei[:, 1] = ei[:, 0]  # duplicate edge
    ei[1, 2] = ei[0, 2]  # self loop
    x = torch.randn(n, 5, generator=gen)
    ea = torch.rand(e, fe, generator=gen) * 5
    p = R.ginet_conv_init(5, 8, fe, generator=gen)
    p["fc_attention.weight"] = p["fc_attention.weight"] * 4
    z = R.ginet_conv_segment_softmax(x, ei, ea, p)
    ref = _loop_float64(x, ei, ea, p)
    assert z.dtype == torch.float32
    assert torch.allclose(z.double(), ref, rtol=1e-5, atol=1e-6)
    assert not bool(z[n - 2 :].any())  # destinations without edges stay 0 (scatter into zeros, ginet.py:57-58)
    # with a zero attention vector the operator is the mean of the projected neighbours
    p0 = dict(p)
    p0["fc_attention.weight"] = torch.zeros_like(p["fc_attention.weight"])
    z0 = R.ginet_conv_segment_softmax(x, ei, ea, p0)
    proj = x @ p["fc.weight"].T
    deg = torch.zeros(n).index_add_(0, ei[0], torch.ones(e))
    mean = torch.zeros(n, 8).index_add_(0, ei[0], proj[ei[1]]) / deg.clamp(min=1).unsqueeze(1)
    assert torch.allclose(z0, mean, rtol=1e-5, atol=1e-6)


def test_attention_switch_arguments():
    from deeprank2_b200.neuralnets.gnn import ginet, ginet_nocluster
    from deeprank2_b200.neuralnets.gnn._common import GINetConvLayer

    with pytest.raises(ValueError):
        GINetConvLayer(4, 16, 1, attention="nope")
    with pytest.raises(NotImplementedError):
        GINetConvLayer(4, 16, 1, bias=True, attention="segment_softmax")
    for mod in (ginet, ginet_nocluster):
        ref_keys = list(mod.GINet(7, 2, 3).state_dict())
        net = mod.GINet(7, 2, 3, attention="segment_softmax")
        assert list(net.state_dict()) == ref_keys  # same checkpoint layout in both modes
        assert all(c.attention == "segment_softmax" for c in (net.conv1, net.conv2, net.conv1_ext, net.conv2_ext))
    assert not ginet_nocluster.GINet(7, 2, 3, attention="segment_softmax")._stackable()
    assert ginet_nocluster.GINet(7, 2, 3)._stackable()


def _attention_golden(case):
    import os

    import numpy as np

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attention_segment_softmax.npz"))
    pre = case + "/"
    return {k[len(pre):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(pre)}


@pytest.mark.parametrize("case", ["a", "b"])
def test_segment_softmax_restatement_vs_golden(case):
    """``tests/golden/attention_segment_softmax.npz`` (``oracle/make_golden_attention.py``): logits computed by the REFERENCE'S OWN
    module up to ``ginet.py:52``, normalisation per destination and gradients in float64.  The fp32 restatement must reproduce
    the logits' consequences -- output and every gradient -- at the path's tolerance."""
    from conftest import assert_close

    g = _attention_golden(case)
    p = {k[2:]: v.clone().requires_grad_(True) for k, v in g.items() if k.startswith("w/")}
    x = g["in/x"].clone().requires_grad_(True)
    z = R.ginet_conv_segment_softmax(x, g["in/edge_index"], g["in/edge_attr"], p)
    assert_close(z, g["out/z"], "z")
    (z * g["gout/z"]).sum().backward()
    assert_close(x.grad, g["grad/x"], "dx")
    for k, v in p.items():
        assert_close(v.grad, g["grad/" + k], f"grad {k}")
