"""Synthetic protein-protein-interface graphs with the statistics of DeepRank2 data.

There is no network for datasets, so the benchmark and the large parity cases use
point clouds whose contact graphs have the node count / degree of the reference's
residue-level and atom-level graphs:

* residue level: n ~ U{240..360} points uniform in a ball of density 0.0105 A^-3,
  contact edges at 8.5 A  ->  mean directed degree ~ 20;
* atom level:    n ~ U{2700..3300}, density 0.059 A^-3, 4.5 A cutoff -> degree ~ 20.

The per-graph tensors follow ``GraphDataset.load_one_graph`` exactly (reference
``deeprank2/dataset.py:937-1004``): ``edge_index = vstack((pairs, flip(pairs, 1))).T``
(all (i,j) first, then all (j,i), same order), ``edge_attr`` duplicated the same way,
``x`` float32 [n,F], ``y`` float32 [1], ``pos`` float32 [n,3].  Graph ``g`` is seeded with
``np.random.default_rng(seed + g)`` so every rank / test regenerates identical data.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import Batch, Data

RESIDUE = dict(n_lo=240, n_hi=360, density=0.0105, cutoff=8.5)
ATOM = dict(n_lo=2700, n_hi=3300, density=0.059, cutoff=4.5)


def _ball(rng: np.random.Generator, n: int, density: float) -> np.ndarray:
    radius = (3.0 * n / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    r = radius * rng.random(n) ** (1.0 / 3.0)
    return direction * r[:, None]


def contact_pairs(pos: np.ndarray, cutoff: float) -> np.ndarray:
    """Undirected contact pairs (i<j) within ``cutoff`` in cKDTree's order, int64 [P,2]."""
    from scipy.spatial import cKDTree

    if pos.shape[0] < 2:
        return np.zeros((0, 2), dtype=np.int64)
    return cKDTree(pos).query_pairs(cutoff, output_type="ndarray").astype(np.int64).reshape(-1, 2)


def make_graph(
    g: int,
    n_node_features: int = 50,
    n_edge_features: int = 1,
    level: dict = RESIDUE,
    seed: int = 1000,
    n: int | None = None,
    with_clusters: bool = False,
) -> Data:
    rng = np.random.default_rng(seed + g)
    if n is None:
        n = int(rng.integers(level["n_lo"], level["n_hi"] + 1))
    pos = _ball(rng, n, level["density"])
    pairs = contact_pairs(pos, level["cutoff"])
    both = np.vstack((pairs, np.flip(pairs, 1))).T  # dataset.py:947
    dist = np.linalg.norm(pos[pairs[:, 0]] - pos[pairs[:, 1]], axis=1) if len(pairs) else np.zeros(0)
    edge_cols = [dist]
    for _ in range(n_edge_features - 1):
        edge_cols.append(rng.random(len(pairs)))
    half = np.stack(edge_cols, axis=1) if n_edge_features > 0 else np.zeros((len(pairs), 0))
    edge_attr = np.vstack((half, half))  # dataset.py:994-995
    x = rng.standard_normal((n, n_node_features))
    y = rng.random(1)
    data = Data(
        x=torch.tensor(x, dtype=torch.float),
        edge_index=torch.tensor(np.ascontiguousarray(both), dtype=torch.long).reshape(2, -1),
        edge_attr=torch.tensor(edge_attr, dtype=torch.float),
        y=torch.tensor(y, dtype=torch.float),
        pos=torch.tensor(pos, dtype=torch.float),
    )
    data.entry_names = f"synthetic-{seed}-{g}"
    if with_clusters:
        # MCL is unavailable offline; deterministic two-level clustering with the same
        # structure as clustering/mcl/depth_{0,1}: local, 0-based ids
        c0 = np.arange(n) // 8
        data.cluster0 = torch.tensor(c0, dtype=torch.long)
        data.cluster1 = torch.tensor(np.arange(int(c0.max()) + 1 if n else 0) // 4, dtype=torch.long)
    return data


def make_batch(n_graphs: int = 256, first: int = 0, **kwargs) -> Batch:
    """Collated batch of graphs ``first .. first+n_graphs-1`` (host tensors)."""
    return Batch.from_data_list([make_graph(first + g, **kwargs) for g in range(n_graphs)])
