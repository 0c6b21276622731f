"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the DeepRank2 GNN message-passing path.

A functional, torch-CPU restatement of what the reference computes on the path
named by BASELINE.json.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / the CPU baseline.  The
product package (``deeprank2_b200``) never imports it and has no CPU fallback.

Why torch and not numpy/C: the reference *is* torch-on-CPU (ATen ``index``,
``mm``, ``scatter_add_``, ``_softmax``); restating it with the same ATen calls keeps
both the rounding behaviour and the CPU cost profile of the reference, which is
what the parity tolerance and the ``cpu_baseline`` number are quoted against.

Pinning: the reference's own tests pin no numeric value of
any GNN layer, so this file is pinned against the reference *executed here*:
``oracle/make_golden.py`` runs the unmodified files from ``/root/reference`` (through
``oracle/reference_loader.py``) and stores inputs/outputs/gradients under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those vectors (bit-exact for integer outputs, fp32 allclose otherwise).
The third-party primitives are restated in ``oracle/thirdparty.py`` from the pinned
upstream versions -- that part is "parity unpinned" by construction (the wheels
are absent) and is covered by hand-computed toy cases instead.

Parameters are plain ``dict[str, Tensor]`` keyed exactly like the reference
``state_dict`` so weights move between the reference modules,
this oracle and the CUDA modules without renaming.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from oracle import thirdparty as tp


# =========================================================================== index structures
def csr_by_destination(index: torch.Tensor, num_segments: int):
    """Stable counting sort of ``index`` (int64 [E]) -> (ptr int32 [S+1], perm int32 [E]).

    ``perm`` lists element ids grouped by segment, ascending id inside a segment
    (``torch.sort(stable=True)``), which is also the order in which the reference's
    CPU ``scatter_add_`` visits them (torch_scatter 2.1.2 ``scatter_sum`` ->
    ``Tensor.scatter_add_``; reference call site ``ginet.py:58``).
    """
    index = index.to(torch.int64)
    order = torch.sort(index, stable=True).indices
    counts = torch.bincount(index, minlength=num_segments)[:num_segments]
    ptr = torch.zeros(num_segments + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr.to(torch.int32), order.to(torch.int32)


def graph_csr(edge_index: torch.Tensor, num_nodes: int):
    """Destination-sorted CSR of a DeepRank2 edge list.

    Destination = ``edge_index[0]`` ("row"), gathered source = ``edge_index[1]`` ("col"),
    see ``ginet.py:41,45-46,58`` / ``vanilla_gnn.py:28-30,35``.
    Returns (rowptr int32 [N+1], colidx int32 [E], perm int32 [E]).
    """
    rowptr, perm = csr_by_destination(edge_index[0], num_nodes)
    colidx = edge_index[1][perm.long()].to(torch.int32)
    return rowptr, colidx, perm


def graph_csc(edge_index: torch.Tensor, num_nodes: int):
    """Source-sorted twin (used by the backward pass): (colptr, rowidx, perm)."""
    colptr, perm = csr_by_destination(edge_index[1], num_nodes)
    rowidx = edge_index[0][perm.long()].to(torch.int32)
    return colptr, rowidx, perm


def batch_offsets(batch: torch.Tensor, num_graphs: int | None = None):
    """``ptr`` of PyG collate (``Batch.from_data_list``): node offsets per graph, int64 [B+1].

    ``num_graphs`` defaults to ``batch.max()+1`` as in ``scatter_mean`` without ``dim_size``
    (``ginet_nocluster.py:103``).
    """
    if num_graphs is None:
        num_graphs = int(batch.max()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=num_graphs)
    ptr = torch.zeros(num_graphs + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


# =========================================================================== row A/B: GINetConvLayer
def ginet_conv_init(in_channels: int, out_channels: int, number_edge_features: int = 1, generator=None):
    """Parameter shapes and init of ``GINetConvLayer`` (``ginet.py:23-38``): three bias-free
    Linear weights, ALL drawn from U(+-1/sqrt(in_channels))."""
    bound = 1.0 / math.sqrt(in_channels)

    def u(*shape):
        return (torch.rand(*shape, generator=generator) * 2 - 1) * bound

    return {
        "fc.weight": u(out_channels, in_channels),
        "fc_edge_attr.weight": u(number_edge_features, number_edge_features),
        "fc_attention.weight": u(1, 2 * out_channels + number_edge_features),
    }


def ginet_conv(x, edge_index, edge_attr, p, prefix=""):
    """``GINetConvLayer.forward`` as written (``ginet.py:40-60`` == ``ginet_nocluster.py:37-57``).

    The softmax runs over the singleton axis of ``alpha [E,1]`` so alpha == 1; it is kept
    so autograd produces the exact-zero gradients of the two dead weights.
    """
    row, col = edge_index[0], edge_index[1]
    if edge_attr.dim() == 1:
        edge_attr = edge_attr.unsqueeze(-1)
    w = p[prefix + "fc.weight"]
    b = p.get(prefix + "fc.bias")
    x_col = F.linear(x[col], w, b)
    x_row = F.linear(x[row], w, b)
    ed = F.linear(edge_attr, p[prefix + "fc_edge_attr.weight"], p.get(prefix + "fc_edge_attr.bias"))
    alpha = F.linear(torch.cat([x_row, x_col, ed], dim=1), p[prefix + "fc_attention.weight"], p.get(prefix + "fc_attention.bias"))
    alpha = F.softmax(F.leaky_relu(alpha), dim=1)
    h = alpha * x_col
    out = torch.zeros(x.shape[0], w.shape[0])
    return tp.scatter_sum(h, row, dim=0, out=out)


def ginet_conv_effective(x, edge_index, w):
    """What row B reduces to (alpha == 1): ``z[i] = sum_{e: row[e]=i} W x[col[e]]``."""
    z = torch.zeros(x.shape[0], w.shape[0], dtype=x.dtype)
    return z.index_add_(0, edge_index[0], F.linear(x[edge_index[1]], w))


def ginet_conv_segment_softmax(x, edge_index, edge_attr, p, prefix=""):
    """The *intended* operator of BASELINE.json north_star (softmax of the attention logit
    over the edges of each destination).  The reference never computes this (it uses
    ``dim=1``): PARITY UNPINNED for the normalisation; the logit itself is pinned to the reference's own modules
    through ``tests/golden/attention_segment_softmax.npz`` (``oracle/make_golden_attention.py``)."""
    row, col = edge_index[0], edge_index[1]
    if edge_attr.dim() == 1:
        edge_attr = edge_attr.unsqueeze(-1)
    w = p[prefix + "fc.weight"]
    x_col = F.linear(x[col], w)
    x_row = F.linear(x[row], w)
    ed = F.linear(edge_attr, p[prefix + "fc_edge_attr.weight"])
    logit = F.leaky_relu(F.linear(torch.cat([x_row, x_col, ed], dim=1), p[prefix + "fc_attention.weight"])).squeeze(1)
    n = x.shape[0]
    seg_max = torch.full((n,), -math.inf).scatter_reduce(0, row, logit, reduce="amax")
    ex = torch.exp(logit - seg_max[row])
    denom = torch.zeros(n).index_add_(0, row, ex)
    alpha = (ex / denom[row]).unsqueeze(1)
    return torch.zeros(n, w.shape[0]).index_add_(0, row, alpha * x_col)


# =========================================================================== row E: readout
def mean_readout(x, batch, num_graphs=None):
    """``scatter_mean(x, batch, dim=0)`` (``ginet_nocluster.py:103-104``, ``vanilla_gnn.py:62``)."""
    return tp.scatter_mean(x, batch, dim=0, dim_size=num_graphs)


# =========================================================================== row C: GINet
def _ginet_init(input_shape, output_shape, input_shape_edge, generator):
    p = {}
    for name, (fi, fo) in {"conv1": (input_shape, 16), "conv2": (16, 32), "conv1_ext": (input_shape, 16), "conv2_ext": (16, 32)}.items():
        for k, v in ginet_conv_init(fi, fo, input_shape_edge, generator).items():
            p[f"{name}.{k}"] = v
    p.update(_linear_init("fc1", 64, 128, generator))
    p.update(_linear_init("fc2", 128, output_shape, generator))
    return p


def _linear_init(name, fan_in, fan_out, generator=None):
    """``nn.Linear`` default init: kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(fan_in)) for weight and bias."""
    bound = 1.0 / math.sqrt(fan_in)
    return {
        f"{name}.weight": (torch.rand(fan_out, fan_in, generator=generator) * 2 - 1) * bound,
        f"{name}.bias": (torch.rand(fan_out, generator=generator) * 2 - 1) * bound,
    }


def ginet_nocluster_init(input_shape, output_shape=1, input_shape_edge=1, generator=None):
    """``ginet_nocluster.GINet.__init__`` (``ginet_nocluster.py:72-82``)."""
    return _ginet_init(input_shape, output_shape, input_shape_edge, generator)


def _ginet_head(g, p, training, dropout_p, keep=None):
    """fc1 -> ReLU -> dropout(0.4, training) -> fc2 (``ginet_nocluster.py:106-109``).
    ``keep`` ([B,128] of 0 or 1/(1-p)) replays a given dropout mask instead of drawing one (the CUDA path draws its
    masks with Philox on the device; the parity test feeds them back here)."""
    g = F.relu(F.linear(g, p["fc1.weight"], p["fc1.bias"]))
    g = g * keep if keep is not None else F.dropout(g, dropout_p, training=training)
    return F.linear(g, p["fc2.weight"], p["fc2.bias"])


def ginet_nocluster_forward(p, data, training=False, dropout_p=0.4, keep=None, conv=None):
    """``ginet_nocluster.GINet.forward`` (``ginet_nocluster.py:84-111``): two branches of
    conv(50->16) -> ReLU -> conv(16->32) -> ReLU on the same graph, per-graph mean, MLP head.
    ``conv``: the convolution restatement to use (default ``ginet_conv``, the reference's; ``ginet_conv_segment_softmax`` for the
    opt-in attention mode)."""
    conv = conv or ginet_conv
    x, ei, ea = data.x, data.edge_index, data.edge_attr
    a = F.relu(conv(x, ei, ea, p, "conv1."))
    a = F.relu(conv(a, ei, ea, p, "conv2."))
    b = F.relu(conv(x, ei, ea, p, "conv1_ext."))
    b = F.relu(conv(b, ei, ea, p, "conv2_ext."))
    g = torch.cat([mean_readout(a, data.batch), mean_readout(b, data.batch)], dim=1)
    return _ginet_head(g, p, training, dropout_p, keep)


# =========================================================================== rows I/J: community pooling
def preloaded_cluster(cluster, batch):
    """``get_preloaded_cluster`` (``community_pooling.py:23-27``): make per-graph cluster ids
    globally unique by adding, graph after graph, ``max(previous graph's ids) + 1``.
    Mutates and returns ``cluster`` like the reference."""
    n_graphs = int(batch.max()) + 1
    for g in range(1, n_graphs):
        cluster[batch == g] += cluster[batch == g - 1].max() + 1
    return cluster


def preloaded_cluster_closed_form(cluster, batch):
    """Same result without the sequential loop: offset of graph g = sum_{h<g} (max_h + 1)."""
    n_graphs = int(batch.max()) + 1
    per_graph_max = torch.zeros(n_graphs, dtype=cluster.dtype).scatter_reduce(0, batch, cluster, reduce="amax", include_self=False)
    offs = torch.cumsum(per_graph_max + 1, 0) - (per_graph_max + 1)
    return cluster + offs[batch]


def community_pool(cluster, data):
    """``community_pooling`` (``community_pooling.py:165-242``) for a batched input:
    relabel clusters, segment-max of x, pooled+coalesced edges (summed attributes),
    mean position, pooled batch vector; ``cluster0/1`` are carried over untouched."""
    cluster, perm = tp.consecutive_cluster(cluster)
    x, _ = tp.scatter_max(data.x, cluster, dim=0)
    edge_index, edge_attr = tp.pool_edge(cluster, data.edge_index, data.edge_attr)
    out = SimpleNamespace(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=tp.pool_batch(perm, data.batch))
    if getattr(data, "pos", None) is not None:
        out.pos = tp.scatter_mean(data.pos, cluster, dim=0)
    if hasattr(data, "cluster0"):
        out.cluster0, out.cluster1 = data.cluster0, data.cluster1
    return out


def ginet_init(input_shape, output_shape=1, input_shape_edge=1, generator=None):
    """clustered ``ginet.GINet.__init__`` (``ginet.py:77-88``) -- same parameters as no-cluster."""
    return _ginet_init(input_shape, output_shape, input_shape_edge, generator)


def _ginet_cluster_branch(p, names, data, conv=None):
    conv = conv or ginet_conv
    c1, c2 = names
    d = SimpleNamespace(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in vars(data).items()})
    d.x = F.relu(conv(d.x, d.edge_index, d.edge_attr, p, c1))
    d = community_pool(preloaded_cluster(d.cluster0, d.batch), d)
    d.x = F.relu(conv(d.x, d.edge_index, d.edge_attr, p, c2))
    x, batch = tp.max_pool_x(preloaded_cluster(d.cluster1, d.batch), d.x, d.batch)
    return tp.scatter_mean(x, batch, dim=0)


def ginet_forward(p, data, training=False, dropout_p=0.4, conv=None):
    """clustered ``ginet.GINet.forward`` (``ginet.py:90-125``).  Both branches start from
    their own deep copy (``data.clone()`` at ``:92``), so the in-place cluster-offset edits of
    one branch do not leak into the other.  ``conv``: see ``ginet_nocluster_forward``."""
    ns = data if isinstance(data, SimpleNamespace) else SimpleNamespace(**{k: v for k, v in vars(data).items()})
    g = torch.cat([_ginet_cluster_branch(p, ("conv1.", "conv2."), ns, conv), _ginet_cluster_branch(p, ("conv1_ext.", "conv2_ext."), ns, conv)], dim=1)
    return _ginet_head(g, p, training, dropout_p)


# =========================================================================== row F/G: Fout
def fout_conv_init(in_channels, out_channels, bias=True, generator=None):
    """``FoutLayer`` parameters (``foutnet.py:25-46``): wc, wn [Fi,Fo], bias [Fo], all U(+-1/sqrt(Fi))."""
    bound = 1.0 / math.sqrt(in_channels)
    p = {
        "wc": (torch.rand(in_channels, out_channels, generator=generator) * 2 - 1) * bound,
        "wn": (torch.rand(in_channels, out_channels, generator=generator) * 2 - 1) * bound,
    }
    if bias:
        p["bias"] = (torch.rand(out_channels, generator=generator) * 2 - 1) * bound
    return p


def fout_conv_loop(x, edge_index, p, prefix=""):
    """``FoutLayer.forward`` literally (``foutnet.py:48-66``): per-node Python loop, O(N*E).
    Small inputs only.  An empty neighbourhood gives ``mean`` of a [0,Fo] slice == NaN."""
    n = x.shape[0]
    center = x @ p[prefix + "wc"]
    neigh = x @ p[prefix + "wn"]
    rows = []
    for node in range(n):
        src = edge_index[1, edge_index[0] == node]
        rows.append(neigh[src].mean(dim=0))
    out = center + torch.stack(rows) if n else center
    if p.get(prefix + "bias") is not None:
        out = out + p[prefix + "bias"]
    return out


def fout_conv(x, edge_index, p, prefix=""):
    """Vectorised equivalent of :func:`fout_conv_loop` (segment sum / count, 0/0 = NaN kept)."""
    n = x.shape[0]
    center = x @ p[prefix + "wc"]
    neigh = x @ p[prefix + "wn"]
    total = torch.zeros(n, neigh.shape[1]).index_add_(0, edge_index[0], neigh[edge_index[1]])
    count = torch.zeros(n).index_add_(0, edge_index[0], torch.ones(edge_index.shape[1]))
    out = center + total / count.unsqueeze(1)
    if p.get(prefix + "bias") is not None:
        out = out + p[prefix + "bias"]
    return out


def foutnet_init(input_shape, output_shape=1, generator=None):
    """``FoutNet.__init__`` (``foutnet.py:83-97``)."""
    p = {}
    for name, (fi, fo) in {"conv1": (input_shape, 16), "conv2": (16, 32)}.items():
        for k, v in fout_conv_init(fi, fo, True, generator).items():
            p[f"{name}.{k}"] = v
    p.update(_linear_init("fc1", 32, 64, generator))
    p.update(_linear_init("fc2", 64, output_shape, generator))
    return p


def foutnet_forward(p, data, conv=fout_conv):
    """``FoutNet.forward`` (``foutnet.py:99-118``)."""
    d = SimpleNamespace(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in vars(data).items()})
    d.x = F.relu(conv(d.x, d.edge_index, p, "conv1."))
    d = community_pool(preloaded_cluster(d.cluster0, d.batch), d)
    d.x = F.relu(conv(d.x, d.edge_index, p, "conv2."))
    x, batch = tp.max_pool_x(preloaded_cluster(d.cluster1, d.batch), d.x, d.batch)
    g = tp.scatter_mean(x, batch, dim=0)
    g = F.relu(F.linear(g, p["fc1.weight"], p["fc1.bias"]))
    return F.linear(g, p["fc2.weight"], p["fc2.bias"])


# =========================================================================== row H: Vanilla ("Naive") network
def vanilla_conv_init(count_node_features, count_edge_features, generator=None):
    """``VanillaConvolutionalLayer.__init__`` (``vanilla_gnn.py:18-24``): message size 32."""
    p = {}
    for k, v in _linear_init("_edge_mlp.0", 2 * count_node_features + count_edge_features, 32, generator).items():
        p[k] = v
    for k, v in _linear_init("_node_mlp.0", count_node_features + 32, count_node_features, generator).items():
        p[k] = v
    return p


def vanilla_conv(x, edge_index, edge_attr, p, prefix=""):
    """``VanillaConvolutionalLayer.forward`` (``vanilla_gnn.py:26-38``): per-edge MLP on
    [x_i, x_j, e] + ReLU, sum per destination i, node MLP on [x, sum] + ReLU."""
    dst, src = edge_index[0], edge_index[1]
    message_in = torch.cat([x[dst], x[src], edge_attr], dim=1)
    messages = F.relu(F.linear(message_in, p[prefix + "_edge_mlp.0.weight"], p[prefix + "_edge_mlp.0.bias"]))
    summed = tp.scatter_sum(messages, dst, dim=0, out=torch.zeros(x.shape[0], messages.shape[1]))
    return F.relu(F.linear(torch.cat([x, summed], dim=1), p[prefix + "_node_mlp.0.weight"], p[prefix + "_node_mlp.0.bias"]))


def vanilla_init(input_shape, output_shape, input_shape_edge, generator=None):
    """``VanillaNetwork.__init__`` (``vanilla_gnn.py:52-57``)."""
    p = {}
    for layer in ("_external1", "_external2"):
        for k, v in vanilla_conv_init(input_shape, input_shape_edge, generator).items():
            p[f"{layer}.{k}"] = v
    p.update(_linear_init("_graph_mlp.0", input_shape, 128, generator))
    p.update(_linear_init("_graph_mlp.2", 128, output_shape, generator))
    return p


def vanilla_forward(p, data):
    """``VanillaNetwork.forward`` (``vanilla_gnn.py:59-65``)."""
    h = vanilla_conv(data.x, data.edge_index, data.edge_attr, p, "_external1.")
    h = vanilla_conv(h, data.edge_index, data.edge_attr, p, "_external2.")
    g = mean_readout(h, data.batch)
    g = F.relu(F.linear(g, p["_graph_mlp.0.weight"], p["_graph_mlp.0.bias"]))
    return F.linear(g, p["_graph_mlp.2.weight"], p["_graph_mlp.2.bias"])


# =========================================================================== SGAT ("next" row f4)
def sgat_conv(x, edge_index, edge_attr, p, prefix="", undirected=True):
    """``SGraphAttentionLayer.forward`` (``sgat.py:56-84``): edge_attr * ([x_i || x_j] W),
    ``scatter_mean`` per destination (with ``out=`` zeros), + bias."""
    row, col = edge_index[0], edge_index[1]
    if edge_attr.dim() == 1:
        edge_attr = edge_attr.unsqueeze(-1)
    alpha = edge_attr * (torch.cat([x[row], x[col]], dim=-1) @ p[prefix + "weight"])
    out = tp.scatter_mean(alpha, row, dim=0, out=torch.zeros(x.shape[0], alpha.shape[1]))
    if not undirected:
        out = tp.scatter_mean(alpha, col, dim=0, out=out)
    if p.get(prefix + "bias") is not None:
        out = out + p[prefix + "bias"]
    return out


# =========================================================================== row K: the train step
def as_parameters(state: dict) -> dict:
    """detach + clone + requires_grad, preserving key order (== ``model.parameters()`` order)."""
    return {k: v.detach().clone().requires_grad_(True) for k, v in state.items()}


def regression_loss(pred, y):
    """``_format_output`` for regress (``trainer.py:828-831``: ``pred.reshape(-1)``) + ``MSELoss``
    (default loss for regression, ``trainer.py:428-432,455``)."""
    return F.mse_loss(pred.reshape(-1), y)


def make_adam(params: dict, lr=1e-3, weight_decay=1e-5):
    """``configure_optimizers`` defaults (``trainer.py:401-419``): Adam, lr 1e-3, weight_decay 1e-5."""
    return torch.optim.Adam(list(params.values()), lr=lr, weight_decay=weight_decay)


def train_step(forward, params, optimizer, data, loss_fn=regression_loss, **fwd_kwargs):
    """Body of ``Trainer._epoch`` for one batch (``trainer.py:682-694``):
    zero_grad -> model(batch) -> loss -> backward -> optimizer.step -> loss.item()."""
    optimizer.zero_grad()
    pred = forward(params, data, **fwd_kwargs)
    loss = loss_fn(pred, data.y)
    loss.backward()
    optimizer.step()
    return pred.detach(), float(loss.detach())
