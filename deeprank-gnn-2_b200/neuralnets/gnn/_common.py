"""Shared pieces of the GNN modules."""
from __future__ import annotations

import math

import torch
from torch import nn

from ... import ops
from ...graph import graph_index


def uniform(size: int, tensor) -> None:
    """``torch_geometric.nn.inits.uniform``: U(-1/sqrt(size), 1/sqrt(size)); ``None`` is a no-op
    (used by ``GINetConvLayer.reset_parameters`` ``ginet.py:33-38``, ``FoutLayer`` ``foutnet.py:42-46``)."""
    if tensor is not None:
        bound = 1.0 / math.sqrt(size)
        tensor.data.uniform_(-bound, bound)


class GINetConvLayer(nn.Module):
    """Drop-in for ``deeprank2.neuralnets.gnn.ginet.GINetConvLayer`` (``ginet.py:13-63``, identical in
    ``ginet_nocluster.py:10-60``): same constructor, parameter names/shapes/initialisation, and
    ``forward(x, edge_index, edge_attr)``.

    ``attention="reference"`` reproduces the reference bit-for-bit in structure: its ``softmax(alpha,
    dim=1)`` runs over a singleton axis, so the layer is ``z = scatter_sum(fc(x[col]), row)`` and the
    two attention weights get exact-zero gradients.

    ``attention="segment_softmax"`` (opt-in, not the reference's arithmetic) normalises the same logit over the edges of
    each destination node -- the operator the layer's name and BASELINE.json's north_star describe
    (``ops.GINetAttentionConvFunction``, ``csrc/drk_attention.cu``); all three weights then receive real gradients.
    """

    ATTENTION_MODES = ("reference", "segment_softmax")

    def __init__(self, in_channels, out_channels, number_edge_features=1, bias=False, attention="reference"):
        super().__init__()
        if attention not in self.ATTENTION_MODES:
            raise ValueError(f"attention must be one of {self.ATTENTION_MODES}, got {attention!r}")
        if attention == "segment_softmax" and bias:
            raise NotImplementedError("attention='segment_softmax' is implemented for bias=False (the only setting the reference nets use)")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.attention = attention
        self.fc = nn.Linear(self.in_channels, self.out_channels, bias=bias)
        self.fc_edge_attr = nn.Linear(number_edge_features, number_edge_features, bias=bias)
        self.fc_attention = nn.Linear(2 * self.out_channels + number_edge_features, 1, bias=bias)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        size = self.in_channels
        uniform(size, self.fc.weight)
        uniform(size, self.fc_attention.weight)
        uniform(size, self.fc_edge_attr.weight)

    def forward(self, x, edge_index, edge_attr=None, graph=None, relu=False):
        """``graph``: a prebuilt :class:`GraphIndex` (the nets pass the batch's cached one); when omitted
        it is built from ``edge_index`` here, which is what a stand-alone call of the layer does."""
        if graph is None:
            from ...graph import GraphIndex

            graph = GraphIndex.build(edge_index, x.shape[0])
        if self.attention == "segment_softmax":
            if edge_attr is None:
                raise ValueError("attention='segment_softmax' needs edge_attr")
            return ops.ginet_attention_conv(x, edge_attr, self.fc.weight, self.fc_edge_attr.weight, self.fc_attention.weight, graph, relu=relu)
        # biases of the attention branch (bias=True) are dead parameters as well: they get no grad, like
        # any parameter torch autograd sees only through a softmax over one element ... the reference
        # gives them zeros too, but bias=True is never used by the reference nets.
        return ops.ginet_conv(x, self.fc.weight, graph, bias=self.fc.bias, relu=relu,
                              dead_params=(self.fc_edge_attr.weight, self.fc_attention.weight))

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


def mean_readout(x, data):
    return ops.mean_readout(x, graph_index(data, with_csc=False))  # any cached index serves: the readout only needs the graph offsets


def num_graphs_of(data):
    """Number of graphs of a batch without touching the device: ``ptr`` from the collate, or the hint pooled batches
    carry; ``None`` means "unknown" (callers then fall back to ``batch.max()+1`` like torch_scatter does)."""
    ptr = data.__dict__.get("ptr")
    if ptr is not None:
        return int(ptr.numel()) - 1
    return data.__dict__.get("_num_graphs")
