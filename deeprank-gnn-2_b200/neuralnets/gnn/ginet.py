"""Clustered ``GINet`` (mirror of ``deeprank2/neuralnets/gnn/ginet.py``: ``GINetConvLayer`` ``:13-63``,
``GINet`` ``:66-125``): two branches of conv -> ReLU -> community pooling -> conv -> ReLU -> max pooling,
per-graph mean, MLP head.  ``GINetConvLayer`` is shared with the no-cluster variant."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.functional import dropout, relu

from ... import ops
from ...graph import graph_index
from ...utils.community_pooling import community_pooling, get_preloaded_cluster, max_pool_x, pool_meta
from ._common import GINetConvLayer, num_graphs_of  # noqa: F401


class GINet(nn.Module):
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
        self.clustering = "mcl"
        self.dropout = 0.4

    def _branch(self, data, conv1, conv2):
        ng = num_graphs_of(data)
        x = conv1(data.x, data.edge_index, data.edge_attr, graph=graph_index(data), relu=True)
        data.x = x
        # (the offsets are added in place: on a copy, so that a batch that is used again -- CUDA-graph replay, resident sets -- stays intact)
        cluster = get_preloaded_cluster(data.cluster0.clone(), data.batch, ng)
        data = community_pooling(cluster, data)

        data.x = conv2(data.x, data.edge_index, data.edge_attr, graph=graph_index(data), relu=True)
        cluster = get_preloaded_cluster(data.cluster1.clone(), data.batch, ng)
        x, batch = max_pool_x(cluster, data.x, data.batch, meta=pool_meta(data, 1))
        return ops.scatter_mean(x, batch, dim=0, dim_size=ng)

    def forward(self, data):
        # the reference clones the batch for the second branch (ginet.py:92) because get_preloaded_cluster edits
        # cluster0/cluster1 in place; the same is done here (x / edge tensors are only read, so a shallow copy of
        # everything but the two cluster vectors would do, but clone() keeps the semantics obvious).
        data_ext = data.clone()
        x = self._branch(data, self.conv1, self.conv2)
        x_ext = self._branch(data_ext, self.conv1_ext, self.conv2_ext)

        x = torch.cat([x, x_ext], dim=1)
        x = relu(self.fc1(x))
        x = dropout(x, self.dropout, training=self.training)
        return self.fc2(x)
