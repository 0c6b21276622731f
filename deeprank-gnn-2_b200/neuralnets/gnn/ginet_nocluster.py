"""``GINet`` without community pooling -- the primary benchmark model.

Mirror of ``deeprank2/neuralnets/gnn/ginet_nocluster.py`` (``GINetConvLayer`` ``:10-60``, ``GINet``
``:63-111``): same class names, constructor signatures, ``state_dict`` keys and shapes, so reference
checkpoints load unchanged.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.functional import dropout, relu

from ... import ops
from ...graph import graph_index, max_graph_nodes
from ._common import GINetConvLayer, mean_readout  # noqa: F401  (GINetConvLayer is part of this module's API)


class GINet(nn.Module):
    """Two branches ("external"/"internal") of conv(F->16) -> ReLU -> conv(16->32) -> ReLU on the same
    graph, per-graph mean readout, ``fc1`` 64->128, ReLU, dropout 0.4, ``fc2`` 128->out
    (``ginet_nocluster.py:72-111``)."""

    def __init__(self, input_shape, output_shape=1, input_shape_edge=1, attention="reference"):
        """``attention``: "reference" (the reference's arithmetic: every coefficient is 1) or "segment_softmax" (opt-in:
        the logit normalised over each destination's edges, see ``_common.GINetConvLayer``)."""
        super().__init__()
        self.attention = attention
        self.conv1 = GINetConvLayer(input_shape, 16, input_shape_edge, attention=attention)
        self.conv2 = GINetConvLayer(16, 32, input_shape_edge, attention=attention)

        self.conv1_ext = GINetConvLayer(input_shape, 16, input_shape_edge, attention=attention)
        self.conv2_ext = GINetConvLayer(16, 32, input_shape_edge, attention=attention)

        self.fc1 = nn.Linear(2 * 32, 128)
        self.fc2 = nn.Linear(128, output_shape)
        self.dropout = 0.4
        self.fused = True  # per-graph fused kernels when every graph fits in shared memory; False -> layer kernels

    def _stackable(self) -> bool:
        """The stacked path needs bias-free convolutions of equal shapes in both branches (always true
        for the reference architecture; a user-modified net falls back to layer-by-layer)."""
        convs = (self.conv1, self.conv1_ext, self.conv2, self.conv2_ext)
        return (
            all(c.fc.bias is None and c.attention == "reference" for c in convs)
            and self.conv1.fc.weight.shape == self.conv1_ext.fc.weight.shape
            and self.conv2.fc.weight.shape == self.conv2_ext.fc.weight.shape
            and self.conv2.fc.weight.shape[1] == self.conv1.fc.weight.shape[0]
            and self.conv1.fc.weight.shape[0] % 4 == 0
        )

    def forward(self, data):
        if self.fused and not torch.is_grad_enabled() and not (self.training and self.dropout > 0):
            from ... import fused as _fused

            if data.x.is_cuda and _fused.step_supported(self, data):
                return _fused.ginet_infer(self, data)  # inference: the whole forward pass as one per-graph kernel
        # CSR/CSC + graph offsets, built once on the device and shared by all layers; the CSC half only serves the backward pass
        # (and `perm`, the edge id of every CSR slot, only a pass that reads edge attributes: the reference-mode convolutions do not)
        g = graph_index(data, with_csc=torch.is_grad_enabled(), with_perm=torch.is_grad_enabled() or not self._stackable())
        # the reference deep-copies the batch (data.clone(), :86) and overwrites data.x in place (:90,:93);
        # neither has a numerical effect, so no copy is made here.
        if self._stackable():
            # both branches + readout as one autograd node: x -> [B, 64]
            fi = data.x.shape[1]
            fusable = (
                self.fused
                and self.conv1.fc.weight.shape[0] == 16
                and self.conv2.fc.weight.shape == (32, 16)
                and fi <= 64
                and not data.x.requires_grad
                and g.graph_ptr is not None
            )
            biggest = max_graph_nodes(data, g) if fusable else 0
            if fusable and biggest <= ops.ginet_fused_max_nodes(fi):
                # one CTA per graph, intermediates in shared memory (2 launches per train step)
                x = ops.ginet_fused(data.x, self.conv1, self.conv1_ext, self.conv2, self.conv2_ext, g, biggest, data.meta("max_graph_edges", 0) if hasattr(data, "meta") else 0)
            else:
                x = ops.ginet_stack(data.x, self.conv1, self.conv1_ext, self.conv2, self.conv2_ext, g)
        else:
            x0 = data.x
            x = self.conv1(x0, data.edge_index, data.edge_attr, graph=g, relu=True)
            x = self.conv2(x, data.edge_index, data.edge_attr, graph=g, relu=True)
            x_ext = self.conv1_ext(x0, data.edge_index, data.edge_attr, graph=g, relu=True)
            x_ext = self.conv2_ext(x_ext, data.edge_index, data.edge_attr, graph=g, relu=True)
            x = torch.cat([mean_readout(x, data), mean_readout(x_ext, data)], dim=1)
        x = relu(self.fc1(x))
        x = dropout(x, self.dropout, training=self.training)
        return self.fc2(x)
