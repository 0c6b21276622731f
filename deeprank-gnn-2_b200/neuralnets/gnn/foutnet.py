"""``FoutLayer`` / ``FoutNet`` (mirror of ``deeprank2/neuralnets/gnn/foutnet.py:13-118``):
eq. (1) of Fout et al., "Protein Interface Prediction using Graph Convolutional Networks", NIPS 2017."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.functional import relu

from ... import ops
from ...graph import GraphIndex, graph_index
from ...utils.community_pooling import community_pooling, get_preloaded_cluster, max_pool_x, pool_meta
from ._common import num_graphs_of, uniform


class FoutLayer(nn.Module):
    """``out = x Wc + mean_{j in N(i)} x_j Wn + b``; parameters ``wc, wn [Fi,Fo]``, ``bias [Fo]``, all
    U(+-1/sqrt(Fi)) (``foutnet.py:25-46``).  An empty neighbourhood yields a NaN row, as in the reference."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.wc = nn.Parameter(torch.Tensor(in_channels, out_channels))
        self.wn = nn.Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        size = self.in_channels
        uniform(size, self.wc)
        uniform(size, self.wn)
        uniform(size, self.bias)

    def forward(self, x, edge_index, graph=None, relu=False):
        if graph is None:
            graph = GraphIndex.build(edge_index, x.shape[0])
        return ops.fout_conv(x, self.wc, self.wn, self.bias, graph, relu=relu)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


class FoutNet(nn.Module):
    """conv1 -> ReLU -> community pooling (depth 0) -> conv2 -> ReLU -> max pooling (depth 1) -> mean readout
    -> ``fc1`` 32->64 -> ReLU -> ``fc2`` (``foutnet.py:83-118``)."""

    def __init__(self, input_shape, output_shape=1, input_shape_edge=None):  # noqa: ARG002
        super().__init__()
        self.conv1 = FoutLayer(input_shape, 16)
        self.conv2 = FoutLayer(16, 32)
        self.fc1 = nn.Linear(32, 64)
        self.fc2 = nn.Linear(64, output_shape)
        self.clustering = "mcl"

    def forward(self, data):
        ng = num_graphs_of(data)
        data.x = self.conv1(data.x, data.edge_index, graph=graph_index(data), relu=True)
        # (the offsets are added in place: on a copy, so that a batch that is used again -- CUDA-graph replay, resident sets -- stays intact)
        cluster = get_preloaded_cluster(data.cluster0.clone(), data.batch, ng)
        data = community_pooling(cluster, data)

        data.x = self.conv2(data.x, data.edge_index, graph=graph_index(data), relu=True)
        cluster = get_preloaded_cluster(data.cluster1.clone(), data.batch, ng)
        x, batch = max_pool_x(cluster, data.x, data.batch, meta=pool_meta(data, 1))

        x = ops.scatter_mean(x, batch, dim=0, dim_size=ng)
        x = relu(self.fc1(x))
        return self.fc2(x)
