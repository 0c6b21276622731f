"""``VanillaNetwork`` -- the "NaiveNetwork" of the north star (the old name survives only in a stale UML of
the reference, ``tests/utils/uml_training.svg``).  Mirror of ``deeprank2/neuralnets/gnn/vanilla_gnn.py``
(``VanillaConvolutionalLayer`` ``:10-38``, ``VanillaNetwork`` ``:41-65``): same constructor signatures and
``state_dict`` keys (``_external{1,2}._edge_mlp.0.*``, ``_external{1,2}._node_mlp.0.*``, ``_graph_mlp.{0,2}.*``),
verified against ``tests/data/pretrained/testing_graph_model.pth.tar``.
"""
from __future__ import annotations

from torch import nn

from ... import ops
from ...graph import GraphIndex, graph_index


class VanillaConvolutionalLayer(nn.Module):
    """Per-edge MLP on [x_i, x_j, e] -> ReLU -> sum per destination i -> node MLP on [x, sum] -> ReLU."""

    def __init__(self, count_node_features, count_edge_features):
        super().__init__()
        message_size = ops.MESSAGE_SIZE
        edge_input_size = 2 * count_node_features + count_edge_features
        self._edge_mlp = nn.Sequential(nn.Linear(edge_input_size, message_size), nn.ReLU())
        node_input_size = count_node_features + message_size
        self._node_mlp = nn.Sequential(nn.Linear(node_input_size, count_node_features), nn.ReLU())

    def forward(self, node_features, edge_node_indices, edge_features, graph=None):
        if graph is None:
            graph = GraphIndex.build(edge_node_indices, node_features.shape[0])
        return ops.vanilla_conv(node_features, edge_features, self._edge_mlp[0], self._node_mlp[0], graph)


class VanillaNetwork(nn.Module):
    """Two vanilla convolutions, per-graph mean readout, MLP ``F -> 128 -> output_shape``."""

    def __init__(self, input_shape: int, output_shape: int, input_shape_edge: int):
        super().__init__()
        self._external1 = VanillaConvolutionalLayer(input_shape, input_shape_edge)
        self._external2 = VanillaConvolutionalLayer(input_shape, input_shape_edge)
        hidden_size = 128
        self._graph_mlp = nn.Sequential(nn.Linear(input_shape, hidden_size), nn.ReLU(), nn.Linear(hidden_size, output_shape))

    def forward(self, data):
        g = graph_index(data)
        h = self._external1(data.x, data.edge_index, data.edge_attr, graph=g)
        h = self._external2(h, data.edge_index, data.edge_attr, graph=g)
        means_per_graph = ops.mean_readout(h, g)
        return self._graph_mlp(means_per_graph)


# the name used by BASELINE.json / older DeepRank-GNN releases
NaiveNetwork = VanillaNetwork
NaiveConvolutionalLayer = VanillaConvolutionalLayer
