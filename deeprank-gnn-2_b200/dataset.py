"""``GraphDataset``: DeepRank2 graph HDF5 files -> ``Data`` objects for the Trainer.

Same constructor, attributes and ``get(idx)`` contract as ``deeprank2.dataset.GraphDataset`` (reference
``dataset.py:776-1122``): feature selection (``"all"`` / list), per-feature transforms and standardisation,
target selection / filtering / sigmoid-log transform, classification classes, inheritance of all of these from a
training dataset (``train_source``), pre-computed cluster vectors.  The tensors of one graph follow
``load_one_graph`` (``dataset.py:883-1052``) exactly -- see SURVEY.md 8a row D:

    x          float32 [n, F]   np.hstack of the selected node features in list order (multi-channel features
                                contribute several columns)
    edge_index int64  [2, 2E']  all (i,j) of ``edge_features/_index`` first, then all (j,i), same order
    edge_attr  float32 [2E', Fe] the edge features stacked twice the same way
    y          float32 [1]      (sigmoid(log(y)) if target_transform), pos float32 [n,3],
    cluster0 / cluster1 int64   from ``clustering/<method>/depth_{0,1}``

What is different (B200-first): the reference re-opens the HDF5 file for every graph of every epoch
(``dataset.py:893``).  Here every entry is parsed ONCE into host tensors (``_cache``); epochs iterate over
memory.  Files are read with h5py when it is installed and with the bundled ``hdf5_lite`` reader otherwise.
"""
from __future__ import annotations

import inspect
import logging
import os
import warnings

import numpy as np
import torch

from .data import Data
from .domain import edgestorage as Efeat
from .domain import nodestorage as Nfeat
from .domain import targetstorage as targets

_log = logging.getLogger(__name__)


def open_hdf5(path: str):
    """h5py if available, else the bundled read-only reader (same ``File``/group/dataset subset of the API)."""
    try:
        import h5py

        return h5py.File(path, "r")
    except ImportError:
        from . import hdf5_lite

        return hdf5_lite.File(path, "r")


class GraphDataset:
    _INHERITED = ("node_features", "edge_features", "features_transform", "target", "target_transform", "task", "classes", "classes_to_index")

    def __init__(
        self,
        hdf5_path,
        subset=None,
        train_source=None,
        node_features="all",
        edge_features="all",
        features_transform=None,
        clustering_method=None,
        target=None,
        target_transform=False,
        target_filter=None,
        task=None,
        classes=None,
        use_tqdm=True,
        root="./",
        check_integrity=True,
    ):
        if isinstance(hdf5_path, str):
            self.hdf5_paths = [hdf5_path]
        elif isinstance(hdf5_path, list):
            self.hdf5_paths = list(hdf5_path)
        else:
            raise TypeError(f"hdf5_path: unexpected type: {type(hdf5_path)}")
        self.root = root
        self.subset = subset
        self.train_source = train_source
        self.target = target
        self.target_transform = target_transform
        self.target_filter = target_filter
        self.use_tqdm = use_tqdm
        if check_integrity:
            self._drop_unreadable_files()
        self._set_task_and_classes(task, classes)
        self._create_index_entries()

        self.df = None
        self.means = self.devs = None
        self.train_means = self.train_devs = None
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

        sig = inspect.signature(type(self).__init__).parameters
        self.default_vars = {k: v.default for k, v in sig.items() if v.default is not inspect.Parameter.empty}
        self.default_vars["classes_to_index"] = None
        self.node_features = node_features
        self.edge_features = edge_features
        self.clustering_method = clustering_method
        self.features_transform = features_transform
        self._cache: dict = {}

        if train_source is not None:
            self.inherited_params = list(self._INHERITED)
            self._inherit_from_train_source()
            self._check_features()
        else:
            self._check_features()
            self.inherited_params = None
            if not self.index_entries:
                raise IndexError("No entries found in the dataset. Please check the dataset parameters.")
            fname, entry = self.index_entries[0]
            with open_hdf5(fname) as f5:
                present = list(f5[entry][targets.VALUES].keys())
            if self.target is None:
                raise ValueError(f"Please set the target during training dataset definition; targets present in the file/s are {present}.")
            if self.target not in present:
                raise ValueError(f"Target {self.target} not present in the file/s; targets present in the file/s are {present}.")

        self.features_dict = {Nfeat.NODE: self.node_features, Efeat.EDGE: self.edge_features}
        if self.target is not None:
            self.features_dict[targets.VALUES] = [self.target] if isinstance(self.target, str) else self.target

        standardize = bool(self.features_transform) and any(spec.get("standardize") for spec in self.features_transform.values())
        if standardize and train_source is None:
            self.hdf5_to_pandas()
            self._compute_mean_std()
        elif standardize:
            self.means, self.devs = self.train_means, self.train_devs

    # ------------------------------------------------------------------ setup helpers
    def _drop_unreadable_files(self) -> None:
        keep = []
        for path in self.hdf5_paths:
            try:
                with open_hdf5(path) as f5:
                    if len(list(f5.keys())) == 0:
                        _log.info(f"    -> {path} is empty ")
                        continue
                keep.append(path)
            except Exception as e:  # noqa: BLE001
                _log.error(e)
                _log.info(f"    -> {path} is corrupted ")
        self.hdf5_paths = keep

    def _set_task_and_classes(self, task, classes) -> None:
        self.task = targets.DEFAULT_TASK.get(self.target) if task is None else task
        if self.task not in (targets.CLASSIF, targets.REGRESS) and self.target is not None:
            raise ValueError(f"User target detected: {self.target} -> The task argument must be 'classif' or 'regress', currently set as {self.task}")
        if task and task != self.task:
            warnings.warn(f"Target {self.target} expects {self.task}, but was set to task {task} by user. User set task is ignored and {self.task} will be used.")
        if self.task == targets.CLASSIF:
            if classes is None:
                classes = [0, 1, 2, 3, 4, 5] if self.target == targets.CAPRI else [0, 1]
            self.classes = classes
            self.classes_to_index = {c: i for i, c in enumerate(self.classes)}
        else:
            self.classes = None
            self.classes_to_index = None

    def _create_index_entries(self) -> None:
        self.index_entries = []
        for path in self.hdf5_paths:
            try:
                with open_hdf5(path) as f5:
                    names = list(f5.keys())
                    if self.subset is not None:
                        present = set(names)
                        names = [n for n in self.subset if n in present]
                    for name in names:
                        if self.target_filter is None or self._filter_targets(f5[name]):
                            self.index_entries.append((path, name))
            except Exception:  # noqa: BLE001
                _log.exception(f"on {path}")

    def _filter_targets(self, grp) -> bool:
        """``target_filter = {target_name: "<op> value"}`` keeps entries whose stored target satisfies the condition."""
        if self.target_filter is None:
            return True
        present = list(grp[targets.VALUES].keys())
        for name, condition in self.target_filter.items():
            if name not in present:
                _log.warning(f"   :Filter {name} not found for entry {grp}\n   :Filter options are: {present}")
                continue
            if isinstance(condition, str):
                value = grp[targets.VALUES][name][()]
                expr = condition
                for op in (">", "<", "==", "<=", ">=", "!="):
                    expr = expr.replace(op, f"{value}" + op)
                if not eval(expr):  # noqa: S307  (same mini-language as the reference, dataset.py:283-290)
                    return False
            elif condition is not None:
                raise ValueError("Conditions not supported", condition)
        return True

    def _inherit_from_train_source(self) -> None:
        src = self.train_source
        if isinstance(src, str):
            try:
                state = torch.load(src, map_location=None if torch.cuda.is_available() else torch.device("cpu"), weights_only=False)
            except Exception as e:  # noqa: BLE001
                raise ValueError("The path provided to `train_source` is not a valid DeepRank2 pre-trained model.") from e
            if state.get("data_type") is not GraphDataset and getattr(state.get("data_type"), "__name__", "") != "GraphDataset":
                raise TypeError(f"The pre-trained model has been trained with data of type {state.get('data_type')}, not GraphDataset.")
            self.train_means, self.train_devs = state["means"], state["devs"]
            if state["features_transform"]:
                for spec in state["features_transform"].values():
                    if spec["transform"] is not None and isinstance(spec["transform"], str):
                        spec["transform"] = eval(spec["transform"])  # noqa: S307  (lambda source stored by Trainer._save_model)
            values = state
        elif isinstance(src, GraphDataset):
            self.train_means, self.train_devs = src.means, src.devs
            values = vars(src)
        else:
            raise TypeError(f"The train data provided is invalid: {type(src)}.\n\tPlease provide a valid training GraphDataset or the path to a valid DeepRank2 pre-trained model.")
        mine = vars(self)
        for param in self.inherited_params:
            if mine[param] != values[param]:
                if mine[param] != self.default_vars.get(param):
                    _log.warning(f"The {param} parameter set here is: {mine[param]}, which is not equivalent to the one in the training phase: {values[param]}. Overwriting it.")
                setattr(self, param, values[param])

    def _check_features(self) -> None:
        with open_hdf5(self.hdf5_paths[0]) as f5:
            first = next(iter(f5.keys()))
            self.available_node_features = [k for k in f5[f"{first}/{Nfeat.NODE}"].keys() if k[0] != "_"]
            self.available_edge_features = [k for k in f5[f"{first}/{Efeat.EDGE}"].keys() if k[0] != "_"]

        def resolve(requested, available, key):
            if requested == "all":
                self.default_vars[key] = available
                return list(available), []
            if not isinstance(requested, list):
                requested = [] if requested is None else [requested]
            return requested, [f for f in requested if f not in available]

        self.node_features, missing_nodes = resolve(self.node_features, self.available_node_features, "node_features")
        self.edge_features, missing_edges = resolve(self.edge_features, self.available_edge_features, "edge_features")
        if missing_nodes or missing_edges:
            parts = []
            if missing_nodes:
                parts.append(f"\nMissing node features: {missing_nodes}\nAvailable node features: {self.available_node_features}")
            if missing_edges:
                parts.append(f"\nMissing edge features: {missing_edges}\nAvailable edge features: {self.available_edge_features}")
            raise ValueError(
                f"Not all features could be found in the file {self.hdf5_paths[0]}.\n\tCheck feature_modules passed to the preprocess function.\n\t"
                "Probably, the feature wasn't generated during the preprocessing step.\n\t" + "".join(parts)
            )

    # ------------------------------------------------------------------ statistics for standardisation
    def _transform_for(self, feat):
        if not self.features_transform:
            return None, None
        everyone = self.features_transform.get("all", {})
        transform, standard = everyone.get("transform"), everyone.get("standardize")
        own = self.features_transform.get(feat, {}) if feat in self.features_transform else {}
        if transform is None:
            transform = own.get("transform")
        if standard is None:
            standard = own.get("standardize")
        return transform, standard

    def hdf5_to_pandas(self):
        """One row per entry, one column per feature channel (``feat`` or ``feat_<i>``) holding that entry's values
        (transformed), plus ``id`` -- the table ``_compute_mean_std`` reduces (reference ``dataset.py:302-349``)."""
        import pandas as pd

        frames = []
        for fname in self.hdf5_paths:
            with open_hdf5(fname) as f5:
                names = [n for n in f5.keys() if self.subset is None or n in self.subset]
                cols = {"id": names}
                if names:
                    probe = f5[names[0]]
                    for group, feats in self.features_dict.items():
                        for feat in feats:
                            transform, _ = self._transform_for(feat)
                            arr0 = np.asarray(probe[group][feat][()])
                            if arr0.ndim == 2:
                                for ch in range(arr0.shape[1]):
                                    vals = [np.asarray(f5[n][group][feat][()])[:, ch] for n in names]
                                    cols[f"{feat}_{ch}"] = [transform(v) for v in vals] if transform else vals
                            else:
                                vals = [np.asarray(f5[n][group][feat][()]) for n in names]
                                vals = [v if v.ndim == 1 else v[()] for v in vals]
                                cols[feat] = [transform(v) for v in vals] if transform else vals
                frames.append(pd.DataFrame(data=cols))
        self.df = pd.concat(frames).reset_index(drop=True) if frames else pd.DataFrame()
        return self.df

    def _compute_mean_std(self) -> None:
        """mean / std per column over all entries, NaN-aware and ROUNDED TO ONE DECIMAL like the reference (``:455-470``)."""
        means, devs = {}, {}
        for col in self.df.columns[1:]:
            values = self.df[col].to_numpy()
            flat = np.concatenate(values) if isinstance(values[0], np.ndarray) and values[0].ndim > 0 else np.asarray(values, dtype=float)
            means[col] = round(np.nanmean(flat), 1)
            devs[col] = round(np.nanstd(flat), 1)
        self.means, self.devs = means, devs

    # ------------------------------------------------------------------ access
    def len(self) -> int:
        return len(self.index_entries)

    __len__ = len

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return self.get(int(idx))
        raise TypeError("GraphDataset supports integer indexing only")

    def get(self, idx: int) -> Data:
        fname, entry = self.index_entries[idx]
        key = (fname, entry)
        item = self._cache.get(key)
        if item is None:
            item = self.load_one_graph(fname, entry)
            self._cache[key] = item
        return item.clone()

    def preload(self) -> None:
        """Parse every entry now (one pass over the files) instead of lazily on first use."""
        by_file: dict = {}
        for fname, entry in self.index_entries:
            by_file.setdefault(fname, []).append(entry)
        for fname, entries in by_file.items():
            with open_hdf5(fname) as f5:
                for entry in entries:
                    if (fname, entry) not in self._cache:
                        self._cache[(fname, entry)] = self._graph_from_group(f5[entry], fname, entry)

    def load_one_graph(self, fname: str, entry_name: str) -> Data:
        with open_hdf5(fname) as f5:
            return self._graph_from_group(f5[entry_name], fname, entry_name)

    def _feature_block(self, grp, group_name: str, feats: list, fname: str, entry_name: str):
        blocks = []
        for feat in feats:
            if feat[0] == "_":  # meta features (_name, _index, ...) are never inputs
                continue
            vals = np.asarray(grp[f"{group_name}/{feat}"][()])
            transform, standard = self._transform_for(feat)
            if transform:
                with warnings.catch_warnings(record=True) as caught:
                    warnings.simplefilter("always")
                    vals = transform(vals)
                    if caught:
                        raise ValueError(f"Invalid value occurs in {entry_name}, file {fname}, when applying {transform} for feature {feat}.\n\tPlease change the transformation function for {feat}.")
            if vals.ndim == 1:
                vals = vals.reshape(-1, 1)
                if standard:
                    vals = (vals - self.means[feat]) / self.devs[feat]
            elif standard:
                # the reference selects the channel statistics by SUBSTRING match of the feature name (dataset.py:925-926)
                mean = [v for k, v in self.means.items() if feat in k]
                dev = [v for k, v in self.devs.items() if feat in k]
                vals = (vals - mean) / dev
            blocks.append(vals)
        return blocks

    def _graph_from_group(self, grp, fname: str, entry_name: str) -> Data:
        # node features
        if len(self.node_features) > 0:
            x = torch.tensor(np.hstack(self._feature_block(grp, Nfeat.NODE, self.node_features, fname, entry_name)), dtype=torch.float)
        else:
            x = None
            _log.warning("No node features set.")
        # edges, stored once per undirected pair on disk, both directions in memory
        if Efeat.INDEX in grp[Efeat.EDGE]:
            ind = np.asarray(grp[f"{Efeat.EDGE}/{Efeat.INDEX}"][()])
            if ind.ndim == 2:
                ind = np.vstack((ind, np.flip(ind, 1))).T
            edge_index = torch.tensor(np.ascontiguousarray(ind), dtype=torch.long).contiguous()
        else:
            edge_index = torch.empty((2, 0), dtype=torch.long)
        if len(self.edge_features) > 0:
            half = np.hstack(self._feature_block(grp, Efeat.EDGE, self.edge_features, fname, entry_name))
            edge_attr = torch.tensor(np.vstack((half, half)), dtype=torch.float).contiguous()
        else:
            edge_attr = torch.empty((edge_index.shape[1], 0), dtype=torch.float)
        # target
        y = None
        if self.target is not None:
            if targets.VALUES in grp and self.target in grp[targets.VALUES]:
                y = torch.tensor([grp[f"{targets.VALUES}/{self.target}"][()]], dtype=torch.float).contiguous()
                if self.target_transform is True:
                    if self.task != targets.REGRESS:
                        raise ValueError(f'Sigmoid transformation not possible for {self.task} tasks. Please change `task` to "regress" or set `target_transform` to `False`.')
                    y = torch.sigmoid(torch.log(y))
            elif self.train_source is None:
                present = list(grp[targets.VALUES].keys()) if targets.VALUES in grp else []
                raise ValueError(f"Target {self.target} missing in entry {entry_name} in file {fname}, possible targets are {present}.\n\tUse the query class to add more target values to input data.")
        pos = torch.tensor(np.asarray(grp[f"{Nfeat.NODE}/{Nfeat.POSITION}"][()]), dtype=torch.float).contiguous()
        # pre-computed clusters
        cluster0 = cluster1 = None
        if self.clustering_method is not None and "clustering" in grp:
            if self.clustering_method in grp["clustering"]:
                cg = grp[f"clustering/{self.clustering_method}"]
                if "depth_0" in cg and "depth_1" in cg:
                    cluster0 = torch.tensor(np.asarray(cg["depth_0"][()]), dtype=torch.long)
                    cluster1 = torch.tensor(np.asarray(cg["depth_1"][()]), dtype=torch.long)
                else:
                    _log.warning("no clusters detected")
            else:
                _log.warning(f"no clustering/{self.clustering_method} detected")
        data = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos)
        data.cluster0 = cluster0
        data.cluster1 = cluster1
        data.entry_names = entry_name
        return data


def file_size_of(path: str) -> int:
    return os.path.getsize(path)


class InMemoryGraphDataset(GraphDataset):
    """A :class:`GraphDataset` over graphs that already live in host memory (synthetic benchmarks, tests, or data produced
    by another pipeline).  It behaves like a training dataset whose HDF5 files have been parsed: same attributes
    (``node_features``, ``edge_features``, ``target``, ``task``, ``classes`` ...), same ``get`` / ``len`` contract."""

    def __init__(self, graphs, target="y", task=targets.REGRESS, classes=None, node_features=None, edge_features=None, clustering_method=None, train_source=None):
        if len(graphs) == 0:
            raise IndexError("No entries found in the dataset. Please check the dataset parameters.")
        self.hdf5_paths = []
        self.root = "./"
        self.subset = None
        self.train_source = train_source
        self.target = target
        self.target_transform = False
        self.target_filter = None
        self.use_tqdm = False
        self._set_task_and_classes(task, classes)
        first = graphs[0]
        self.node_features = node_features if node_features is not None else [f"x{i}" for i in range(first.num_node_features)]
        self.edge_features = edge_features if edge_features is not None else [f"e{i}" for i in range(first.num_edge_features)]
        self.available_node_features, self.available_edge_features = self.node_features, self.edge_features
        self.features_transform = None
        self.clustering_method = clustering_method
        self.df = None
        self.means = self.devs = self.train_means = self.train_devs = None
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.default_vars = {}
        self.inherited_params = None
        self.features_dict = {Nfeat.NODE: self.node_features, Efeat.EDGE: self.edge_features, targets.VALUES: [self.target]}
        self.index_entries = []
        self._cache = {}
        for i, g in enumerate(graphs):
            name = getattr(g, "entry_names", None) or f"graph-{i}"
            if not hasattr(g, "cluster0"):
                g.cluster0 = g.cluster1 = None
            g.entry_names = name
            key = ("<memory>", name)
            self.index_entries.append(key)
            self._cache[key] = g

    def load_one_graph(self, fname: str, entry_name: str) -> Data:
        raise KeyError(f"{entry_name} is not part of this in-memory dataset")
