"""``SGraphAttentionLayer`` / ``SGAT`` (mirror of ``deeprank2/neuralnets/gnn/sgat.py:13-136``).

``z_i = 1/N_i sum_j a_ij [x_i || x_j] W + b`` with ``a_ij`` the (single) edge attribute.  Splitting ``W`` by rows into
the ``x_i`` half ``Wt`` and the ``x_j`` half ``Wb``:  ``z_i = ( (x Wt)_i * sum_j a_ij + sum_j a_ij (x Wb)_j ) / max(N_i,1) + b``,
i.e. one [N, 2Fo] projection and one weighted segmented mean instead of an [E, 2F] x [2F, Fo] GEMM.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.functional import relu

from ... import ops
from ...graph import GraphIndex, graph_index
from ...utils.community_pooling import community_pooling, get_preloaded_cluster, max_pool_x, pool_meta
from ._common import num_graphs_of, uniform


class SGraphAttentionLayer(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, undirected: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.undirected = undirected
        self.weight = nn.Parameter(torch.Tensor(2 * in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        size = 2 * self.in_channels
        uniform(size, self.weight)
        uniform(size, self.bias)

    def forward(self, x, edge_index, edge_attr, graph=None):
        if not self.undirected:
            raise NotImplementedError("undirected=False (second scatter over `col`, sgat.py:77-78) is never used by the reference nets")
        if edge_attr.dim() == 2:
            if edge_attr.shape[1] != 1:
                # sgat.py:68 broadcasts edge_attr [E,Fe] against [E,Fo]: only Fe == 1 works in the reference too
                raise ValueError("SGraphAttentionLayer needs exactly one edge feature")
            edge_attr = edge_attr[:, 0]
        if graph is None:
            graph = GraphIndex.build(edge_index, x.shape[0])
        return ops.sgat_conv(x, edge_attr, self.weight, self.bias, graph)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


class SGAT(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=None):  # noqa: ARG002
        super().__init__()
        self.conv1 = SGraphAttentionLayer(input_shape, 16)
        self.conv2 = SGraphAttentionLayer(16, 32)
        self.fc1 = nn.Linear(32, 64)
        self.fc2 = nn.Linear(64, output_shape)
        self.clustering = "mcl"

    def forward(self, data):
        ng = num_graphs_of(data)
        data.x = relu(self.conv1(data.x, data.edge_index, data.edge_attr, graph=graph_index(data)))
        # (the offsets are added in place: on a copy, so that a batch that is used again -- CUDA-graph replay, resident sets -- stays intact)
        cluster = get_preloaded_cluster(data.cluster0.clone(), data.batch, ng)
        data = community_pooling(cluster, data)

        data.x = relu(self.conv2(data.x, data.edge_index, data.edge_attr, graph=graph_index(data)))
        cluster = get_preloaded_cluster(data.cluster1.clone(), data.batch, ng)
        x, batch = max_pool_x(cluster, data.x, data.batch, meta=pool_meta(data, 1))

        x = ops.scatter_mean(x, batch, dim=0, dim_size=ng)
        x = relu(self.fc1(x))
        return self.fc2(x)
