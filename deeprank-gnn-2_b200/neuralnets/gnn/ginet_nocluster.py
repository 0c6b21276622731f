"""``GINet`` without community pooling -- the primary benchmark model.

Mirror of ``deeprank2/neuralnets/gnn/ginet_nocluster.py`` (``GINetConvLayer`` ``:10-60``, ``GINet``
``:63-111``): same class names, constructor signatures, ``state_dict`` keys and shapes, so reference
checkpoints load unchanged.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.functional import dropout, relu

from ...graph import graph_index
from ._common import GINetConvLayer, mean_readout  # noqa: F401  (GINetConvLayer is part of this module's API)


class GINet(nn.Module):
    """Two branches ("external"/"internal") of conv(F->16) -> ReLU -> conv(16->32) -> ReLU on the same
    graph, per-graph mean readout, ``fc1`` 64->128, ReLU, dropout 0.4, ``fc2`` 128->out
    (``ginet_nocluster.py:72-111``)."""

    def __init__(self, input_shape, output_shape=1, input_shape_edge=1):
        super().__init__()
        self.conv1 = GINetConvLayer(input_shape, 16, input_shape_edge)
        self.conv2 = GINetConvLayer(16, 32, input_shape_edge)

        self.conv1_ext = GINetConvLayer(input_shape, 16, input_shape_edge)
        self.conv2_ext = GINetConvLayer(16, 32, input_shape_edge)

        self.fc1 = nn.Linear(2 * 32, 128)
        self.fc2 = nn.Linear(128, output_shape)
        self.dropout = 0.4

    def forward(self, data):
        g = graph_index(data)  # CSR/CSC + graph offsets, built once on the device and shared by all layers
        x0 = data.x
        # the reference deep-copies the batch (data.clone(), :86) and overwrites data.x in place (:90,:93);
        # neither has a numerical effect, so no copy is made here.
        x = self.conv1(x0, data.edge_index, data.edge_attr, graph=g, relu=True)
        x = self.conv2(x, data.edge_index, data.edge_attr, graph=g, relu=True)
        x_ext = self.conv1_ext(x0, data.edge_index, data.edge_attr, graph=g, relu=True)
        x_ext = self.conv2_ext(x_ext, data.edge_index, data.edge_attr, graph=g, relu=True)

        x = mean_readout(x, data)
        x_ext = mean_readout(x_ext, data)

        x = torch.cat([x, x_ext], dim=1)
        x = relu(self.fc1(x))
        x = dropout(x, self.dropout, training=self.training)
        return self.fc2(x)
