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

    def _branch(self, data, conv1, conv2, shared):
        """``shared``: the pooling structure of the batch (relabelled clusters, pooled edges / attributes / positions / batch vector,
        the pooled batch's graph index) -- identical for both branches, built by the first and reused by the second."""
        ng = num_graphs_of(data)
        x = conv1(data.x, data.edge_index, data.edge_attr, graph=graph_index(data), relu=True)
        data.x = x
        # (the offsets are added in place: on a copy, so that a batch that is used again -- CUDA-graph replay, resident sets -- stays intact)
        cluster = get_preloaded_cluster(data.cluster0.clone(), data.batch, ng) if "level0" not in shared else None
        data = community_pooling(cluster, data, shared=shared)

        data.x = conv2(data.x, data.edge_index, data.edge_attr, graph=graph_index(data), relu=True)
        cluster = get_preloaded_cluster(data.cluster1.clone(), data.batch, ng) if "level1" not in shared else None
        x, batch = max_pool_x(cluster, data.x, data.batch, meta=pool_meta(data, 1), shared=shared)
        return ops.scatter_mean(x, batch, dim=0, dim_size=ng)

    def forward(self, data):
        # The reference deep-copies the batch for the second branch (ginet.py:92) because get_preloaded_cluster edits cluster0 / cluster1
        # in place.  Here the cluster vectors are offset on copies (see _branch), nothing else of the batch is written but `x`: a shallow view
        # suffices, and both branches share the batch's graph index and its pooling structure.
        import copy

        graph_index(data)
        data_ext = copy.copy(data)
        data_ext.__dict__ = dict(data.__dict__)
        shared: dict = {}
        x = self._branch(data, self.conv1, self.conv2, shared)
        x_ext = self._branch(data_ext, self.conv1_ext, self.conv2_ext, shared)

        x = torch.cat([x, x_ext], dim=1)
        x = relu(self.fc1(x))
        x = dropout(x, self.dropout, training=self.training)
        return self.fc2(x)
