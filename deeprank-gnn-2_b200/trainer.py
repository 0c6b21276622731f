"""``Trainer``: the caller of the message-passing path.

Public surface of ``deeprank2.trainer.Trainer`` (reference ``trainer.py:27-1004``): constructor arguments and their
validation errors, ``configure_optimizers``, ``set_lossfunction``, ``train``, ``test``, the checkpoint dictionary keys
of ``_save_model`` (``:926-956``) and the exporter hook.  What changes is how a step is executed:

* batches are collated on the host from the dataset's in-memory cache, pinned, and copied to the device
  asynchronously; the graph index (CSR/CSC/offsets) is built on the device as part of the step;
* nothing is read back per step: losses, predictions and targets stay on the device and are fetched ONCE per
  pass (the reference synchronises three times per step, ``trainer.py:694-703``); the exporters receive exactly the
  same lists;
* evaluation runs under ``torch.no_grad()`` (the reference builds and drops autograd graphs in ``_eval``);
* under ``torchrun`` (``torch.distributed`` initialised) every rank trains on its shard of each mini-batch; for the benchmark
  model the gradient exchange happens inside the step's finalize kernel over NVLink peer memory (``fused.GINetFusedStep``), every
  other network sums its gradients with one NCCL all-reduce per step (``parallel.GradAllReduce``) -- this replaces
  ``nn.DataParallel`` (``:387-389``).  The epoch loss, the train / validation split and the per-batch choice between the two paths
  are made identical on all ranks; the exporters receive the whole pass on rank 0.
"""
from __future__ import annotations

import copy
import inspect
import logging
import re
import warnings
from time import time
from typing import Any

import numpy as np
import torch
import torch.distributed as dist
from torch import nn
from torch.nn.functional import softmax

from .data import Batch
from .dataset import GraphDataset
from .domain import losstypes as losses
from .domain import targetstorage as targets
from .utils.earlystopping import EarlyStopping
from .utils.exporters import HDF5OutputExporter, OutputExporter, OutputExporterCollection

_log = logging.getLogger(__name__)


class BatchLoader:
    """Mini-batches of a :class:`GraphDataset` as device-resident ``Batch`` objects.

    Stands in for ``torch_geometric.loader.DataLoader`` (``trainer.py:541-557``): same ``batch_size`` / ``shuffle``
    semantics (a fresh permutation per epoch), collate rules of ``Batch.from_data_list``.  With ``world_size > 1`` each
    rank receives the ``rank``-th contiguous slice of every global mini-batch (``parallel.shard_indices``).
    """

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, device=None, pin_memory: bool = False, rank: int = 0, world_size: int = 1, seed: int | None = None,
                 num_workers: int | None = None):
        """``num_workers``: threads that collate (and page-lock) batches ahead of the consumer, in order (the big concatenations
        release the GIL); ``None`` picks min(4, cores / 2), 0 collates in the calling thread like the reference's default."""
        self.dataset, self.batch_size, self.shuffle = dataset, int(batch_size), shuffle
        self.device, self.pin_memory = device, pin_memory
        self.rank, self.world_size = rank, world_size
        self.only = None
        import os

        self.num_workers = min(4, max(1, (os.cpu_count() or 2) // 2)) if num_workers is None else max(0, int(num_workers))
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(seed)
        elif world_size > 1:
            self._gen.manual_seed(0)  # all ranks must draw the same permutation

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def _host_batches(self):
        from .parallel import shard_indices

        n = len(self.dataset)
        order = torch.randperm(n, generator=self._gen).tolist() if self.shuffle else list(range(n))

        def collate(graphs):
            batch = Batch.from_data_list(graphs)
            if self.pin_memory:
                batch.pin_memory(only=self.only)  # with `only` set, the other tensors travel lazily (if ever): no need to page-lock them
            return batch

        def jobs():
            for start in range(0, n, self.batch_size):
                ids = order[start : start + self.batch_size]
                global_size = len(ids)
                if self.world_size > 1:
                    ids = [ids[i] for i in shard_indices(len(ids), self.rank, self.world_size)]
                # the dataset is read in the calling thread (its parse-once cache and h5py are not meant for concurrent use)
                yield ([self.dataset.get(i) for i in ids] if ids else None), global_size

        if self.num_workers == 0:
            for graphs, global_size in jobs():
                yield (collate(graphs) if graphs is not None else None), global_size
            return
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=self.num_workers, thread_name_prefix="drk-collate") as pool:
            pending: deque = deque()
            for graphs, global_size in jobs():
                pending.append((pool.submit(collate, graphs) if graphs is not None else None, global_size))
                if len(pending) > self.num_workers:  # results leave in submission order
                    fut, gs = pending.popleft()
                    yield (fut.result() if fut is not None else None), gs
            while pending:
                fut, gs = pending.popleft()
                yield (fut.result() if fut is not None else None), gs

    def __iter__(self):
        """Device batches, copied ONE BATCH AHEAD on a side stream (``pipeline.DevicePrefetcher``): the host collate and the PCIe
        copy of batch i+1 overlap the step on batch i.  ``only`` (set by the Trainer when the per-graph step kernels are in use)
        names the tensors to copy eagerly; everything else travels on first access."""
        if self.device is None or self.device.type != "cuda":
            yield from self._host_batches()
            return
        from .pipeline import DevicePrefetcher

        feed = DevicePrefetcher((), self.device, only=self.only)
        pending = None
        for host_batch, global_size in self._host_batches():
            issued = (feed._issue(host_batch) if host_batch is not None else None, global_size)
            if pending is not None:
                yield (feed._hand_out(pending[0]) if pending[0] is not None else None), pending[1]
            pending = issued
        if pending is not None:
            yield (feed._hand_out(pending[0]) if pending[0] is not None else None), pending[1]


class SelectionBatch:
    """A mini-batch of a device-resident graph set: a prepared selection of graph ids (``fused.ResidentGraphSet.select``) plus what the
    Trainer's bookkeeping needs (targets and entry names in the selection's slot order)."""

    __slots__ = ("graph_set", "prepared", "y", "entry_names")

    def __init__(self, graph_set, prepared, y=None):
        self.graph_set, self.prepared = graph_set, prepared
        selection, slot_ids = prepared
        y_all = graph_set.batch.__dict__.get("y")
        self.y = y if y is not None else (y_all.index_select(0, selection.order.long()) if y_all is not None else None)
        names = graph_set.entry_names
        self.entry_names = [names[i] for i in slot_ids] if names is not None else [str(i) for i in slot_ids]


class ResidentBatches:
    """Loader over a :class:`fused.ResidentGraphSet`: the same ``batch_size`` / ``shuffle`` / rank-slice semantics as
    :class:`BatchLoader`, but a batch is a list of graph ids -- no per-epoch collate, no per-batch PCIe copy of the graphs."""

    def __init__(self, graph_set, batch_size: int, shuffle: bool, rank: int = 0, world_size: int = 1, seed: int | None = None, collate: bool = False):
        """``collate=False``: batches are :class:`SelectionBatch` (ids for the per-graph step kernel, graphs read in place);
        ``collate=True``: batches are device ``Batch`` objects gathered out of the resident set (any network)."""
        self.graph_set, self.batch_size, self.shuffle = graph_set, int(batch_size), shuffle
        self.rank, self.world_size = rank, world_size
        self.collate = collate
        self.only = ()
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(seed)
        elif world_size > 1:
            self._gen.manual_seed(0)

    def __len__(self):
        return (self.graph_set.num_graphs + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        from .parallel import shard_indices

        n = self.graph_set.num_graphs
        order = torch.randperm(n, generator=self._gen).tolist() if self.shuffle else list(range(n))
        plan = []
        for start in range(0, n, self.batch_size):
            ids = order[start : start + self.batch_size]
            global_size = len(ids)
            if self.world_size > 1:
                ids = [ids[i] for i in shard_indices(len(ids), self.rank, self.world_size)]
            plan.append((ids, global_size))
        if self.collate:
            for ids, global_size in plan:
                yield (self.graph_set.collate(ids) if ids else None), global_size
            return
        # ids in place: the whole epoch's selections in ONE upload, the targets of the epoch in ONE gather; a batch is two views
        all_ids, prepared = self.graph_set.select_epoch([ids for ids, _ in plan if ids])
        y_all = self.graph_set.batch.__dict__.get("y")
        y_epoch = y_all.index_select(0, all_ids.long()) if y_all is not None and all_ids.numel() else None
        k = off = 0
        for ids, global_size in plan:
            if not ids:
                yield None, global_size
                continue
            sel = prepared[k]
            y = y_epoch[off : off + len(ids)] if y_epoch is not None else None
            k, off = k + 1, off + len(ids)
            yield SelectionBatch(self.graph_set, sel, y=y), global_size


class Trainer:
    def __init__(
        self,
        neuralnet=None,
        dataset_train: GraphDataset | None = None,
        dataset_val: GraphDataset | None = None,
        dataset_test: GraphDataset | None = None,
        val_size: float | int | None = None,
        test_size: float | int | None = None,
        class_weights: bool = False,
        pretrained_model: str | None = None,
        cuda: bool = False,
        ngpu: int = 0,
        output_exporters: list[OutputExporter] | None = None,
    ):
        self.neuralnet = neuralnet
        self.pretrained_model = pretrained_model
        self._init_datasets(dataset_train, dataset_val, dataset_test, val_size, test_size)
        self.cuda, self.ngpu = cuda, ngpu
        self._select_device()
        self._output_exporters = OutputExporterCollection(*(output_exporters if output_exporters is not None else [HDF5OutputExporter("./output")]))

        self.data_type = None
        self.batch_size_train = self.batch_size_test = None
        self.shuffle = None
        self.model_load_state_dict = None
        self._grad_sync = None
        self._fused = None

        if self.pretrained_model is None:
            if self.dataset_train is None:
                raise ValueError("No training data specified. Training data is required if there is no pretrained model.")
            if self.neuralnet is None:
                raise ValueError("No neural network specified. Specifying a model framework is required if there is no pretrained model.")
            self._init_from_dataset(self.dataset_train)
            self.optimizer = None
            self.class_weights = class_weights
            self.subset = self.dataset_train.subset
            self.epoch_saved_model = None
            if self.target is None:
                raise ValueError("No target set. You need to choose a target (set in the dataset) for training.")
            self._load_model()
            if self.clustering_method is not None:
                if self.clustering_method not in ("mcl", "louvain"):
                    raise ValueError(f"Invalid node clustering method: {self.clustering_method}. Please set clustering_method to 'mcl', 'louvain' or None.")
                self._precluster(self.dataset_train)
                if self.dataset_val is not None:
                    self._precluster(self.dataset_val)
                else:
                    _log.warning("No validation dataset given. Randomly splitting training set in training set and validation set.")
                    self.dataset_train, self.dataset_val = _divide_dataset(self.dataset_train, splitsize=self.val_size)
                if self.dataset_test is not None:
                    self._precluster(self.dataset_test)
        else:
            if self.neuralnet is None:
                raise ValueError("No neural network class found. Please add it to complete loading the pretrained model.")
            if self.dataset_test is None:
                raise ValueError("No dataset_test found. Please add it to evaluate the pretrained model.")
            if self.dataset_train is not None:
                self.dataset_train = None
                _log.warning("Pretrained model loaded: dataset_train will be ignored.")
            if self.dataset_val is not None:
                self.dataset_val = None
                _log.warning("Pretrained model loaded: dataset_val will be ignored.")
            self._init_from_dataset(self.dataset_test)
            self._load_params()
            self._load_pretrained_model()

    # ------------------------------------------------------------------ construction helpers
    def _select_device(self) -> None:
        if self.cuda and torch.cuda.is_available():
            local = int(__import__("os").environ.get("LOCAL_RANK", "0")) if self._distributed() else torch.cuda.current_device()
            self.device = torch.device("cuda", local)
            if self.ngpu == 0:
                self.ngpu = 1
                _log.info("CUDA detected. Setting number of GPUs to 1.")
        elif self.cuda:
            msg = "\n--> CUDA not detected: Make sure that CUDA is installed and that you are running on GPUs.\n--> To turn CUDA off set cuda=False in Trainer.\n--> Aborting the experiment \n\n"
            _log.error(msg)
            raise ValueError(msg)
        else:
            self.device = torch.device("cpu")
            if self.ngpu > 0:
                msg = "\n--> CUDA not detected.\n    Set cuda=True in Trainer to turn CUDA on.\n--> Aborting the experiment \n\n"
                _log.error(msg)
                raise ValueError(msg)
        _log.info(f"Device set to {self.device}.")

    @staticmethod
    def _distributed() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _init_datasets(self, dataset_train, dataset_val, dataset_test, val_size, test_size) -> None:
        self._check_dataset_equivalence(dataset_train, dataset_val, dataset_test)
        self.dataset_train, self.dataset_val, self.dataset_test = dataset_train, dataset_val, dataset_test
        self.val_size, self.test_size = val_size, test_size
        if test_size is not None:
            if dataset_test is None:
                self.dataset_train, self.dataset_test = _divide_dataset(dataset_train, test_size)
            else:
                _log.warning("Test dataset was provided to Trainer; test_size parameter is ignored.")
        if val_size is not None:
            if dataset_val is None:
                self.dataset_train, self.dataset_val = _divide_dataset(self.dataset_train, val_size)
            else:
                _log.warning("Validation dataset was provided to Trainer; val_size parameter is ignored.")

    @staticmethod
    def _check_dataset_equivalence(dataset_train, dataset_val, dataset_test) -> None:
        if dataset_train is None:
            if dataset_test is None:
                raise ValueError("Please provide at least a train or test dataset")
            return
        if not isinstance(dataset_train, GraphDataset):
            raise TypeError(f"train dataset is not the right type {type(dataset_train)}. Make sure it's a GraphDataset")
        for other, kind in ((dataset_val, "valid"), (dataset_test, "test")):
            if other is None:
                continue
            if other.train_source is None:
                raise ValueError(f"{kind} dataset has train_source parameter set to None. Make sure to set it as a valid training data source.")
            if other.train_source != dataset_train:
                raise ValueError(f"{kind} dataset has different train_source parameter from Trainer. Make sure to assign equivalent train_source in Trainer.")

    def _init_from_dataset(self, dataset) -> None:
        if not isinstance(dataset, GraphDataset):
            raise TypeError(f"Incorrect `dataset` type provided: {type(dataset)}. Please provide a `GraphDataset` object instead.")
        self.clustering_method = dataset.clustering_method
        self.node_features, self.edge_features = dataset.node_features, dataset.edge_features
        self.features = None
        self.features_transform = dataset.features_transform
        self.means, self.devs = dataset.means, dataset.devs
        self.target, self.target_transform = dataset.target, dataset.target_transform
        self.task, self.classes, self.classes_to_index = dataset.task, dataset.classes, dataset.classes_to_index

    def _load_model(self) -> None:
        self._put_model_to_device(self.dataset_train)
        self.configure_optimizers()
        self.set_lossfunction()

    def _precluster(self, dataset: GraphDataset) -> None:
        """The reference runs MCL / Louvain here and WRITES ``clustering/<method>/depth_{0,1}`` into the HDF5 files
        (``trainer.py:319-348``).  Community detection is CPU preprocessing outside the path and its
        dependencies are not in this image, so files must already carry the clusters (the reference fixtures do)."""
        for i in range(len(dataset)):
            d = dataset.get(i)
            if d.cluster0 is None or d.cluster1 is None:
                fname, entry = dataset.index_entries[i]
                raise NotImplementedError(
                    f"{fname}:{entry} has no clustering/{self.clustering_method}/depth_0,1. Pre-cluster the files with DeepRank2 "
                    "(community detection is offline preprocessing and is not part of deeprank2_b200)."
                )

    def _put_model_to_device(self, dataset: GraphDataset) -> None:
        if self.task == targets.REGRESS:
            self.output_shape = 1
        elif self.task == targets.CLASSIF:
            self.output_shape = len(self.classes)
        first = dataset.get(0)
        target_shape = first.y.shape[0] if first.y is not None else None
        self.model = self.neuralnet(first.num_node_features, self.output_shape, len(dataset.edge_features)).to(self.device)
        if self._distributed():
            from .parallel import GradAllReduce, broadcast_parameters

            broadcast_parameters(self.model)
            self._grad_sync = GradAllReduce(self.model)
        elif self.ngpu > 1:
            raise ValueError(
                "ngpu > 1 needs one process per GPU: launch with `torchrun --nproc-per-node <ngpu>` and call "
                "torch.distributed.init_process_group('nccl') before building the Trainer (replaces nn.DataParallel)."
            )
        for exporter in self._output_exporters:
            if not exporter.is_compatible_with(self.output_shape, target_shape):
                raise ValueError(f"Output exporter of type {type(exporter)}\n\tis not compatible with output shape {self.output_shape}\n\tand target shape {target_shape}.")

    # ------------------------------------------------------------------ public configuration
    def configure_optimizers(self, optimizer=None, lr: float = 0.001, weight_decay: float = 1e-05) -> None:
        self.lr, self.weight_decay = lr, weight_decay
        self._fused = None
        if optimizer is None:
            # same defaults as the reference (trainer.py:416); `fused` keeps the step counters on the device, which lets the
            # per-graph step kernel's finalize apply the update itself (fused.GINetFusedStep)
            on_gpu = next(self.model.parameters()).is_cuda
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay, **({"fused": True} if on_gpu else {}))
            return
        try:
            self.optimizer = optimizer(self.model.parameters(), lr=lr, weight_decay=weight_decay)
        except Exception as e:
            _log.error(e)
            _log.info("Invalid optimizer. Please use only optimizers classes from torch.optim package.")
            raise

    def set_lossfunction(self, lossfunction=None, override_invalid: bool = False) -> None:
        def invalid() -> None:
            text = f"The provided loss function ({lossfunction}) is not appropriate for {self.task} tasks.\n\t"
            if override_invalid:
                _log.warning(text + "You have set override_invalid to True, so the training will run with this loss function nonetheless.\n\tThis will likely cause other errors or exceptions down the line.")
                return
            text += "If you want to use this loss function anyway, set override_invalid to True."
            _log.error(text)
            raise ValueError(text)

        custom = False
        if lossfunction in losses.other_losses:
            invalid()
        elif lossfunction not in (losses.regression_losses + losses.classification_losses):
            custom = True
        fitting = losses.regression_losses if self.task == targets.REGRESS else losses.classification_losses
        if lossfunction is None:
            lossfunction = nn.MSELoss if self.task == targets.REGRESS else nn.CrossEntropyLoss
            _log.info(f"No loss function provided, the default loss function for {self.task} tasks is used: {lossfunction}")
        elif custom:
            _log.warning(f"The provided loss function ({lossfunction}) is not part of the default list.\n\tPlease ensure that this loss function is appropriate for {self.task} tasks.\n\t")
        elif lossfunction not in fitting:
            invalid()
        if self.task == targets.CLASSIF and self.class_weights:
            self.lossfunction = lossfunction  # instantiated with the class weights in train()
        else:
            self.lossfunction = lossfunction()

    # ------------------------------------------------------------------ training / testing
    def _resident_loader(self, dataset, batch_size, shuffle):
        """A :class:`ResidentBatches` loader if the whole dataset fits the device (collated once, kept in HBM), else None.  When the
        per-graph step kernels apply to this model / loss a mini-batch is a list of graph ids the kernels read in place; for every
        other network it is a ``Batch`` cut out of the resident set on the device (``ResidentGraphSet.collate``) -- either way
        there is no per-epoch host collate and no per-batch PCIe copy of the graphs.  ``DRK_NO_RESIDENT=1`` disables it."""
        import os

        from . import _lib
        from .fused import GINetFusedStep, ResidentGraphSet

        if os.environ.get("DRK_NO_RESIDENT") or self.device.type != "cuda" or len(dataset) == 0:
            return None
        cache = self.__dict__.setdefault("_resident_sets", {})
        gset = cache.get(id(dataset))
        if gset is None:
            graphs = [dataset.get(i) for i in range(len(dataset))]
            nbytes = sum(v.numel() * v.element_size() for g in graphs for v in g.__dict__.values() if isinstance(v, torch.Tensor))
            free, _total = torch.cuda.mem_get_info(self.device)
            if 3 * nbytes > free:  # packed copy + pairs + head room
                return None
            gset = ResidentGraphSet(graphs, self.device)
            cache[id(dataset)] = gset
        loss_fn = getattr(self, "lossfunction", None)
        model = getattr(self, "model", None)
        in_place = model is not None and loss_fn is not None and not isinstance(loss_fn, type) and GINetFusedStep.supports(model, loss_fn)
        if in_place:
            fi, out = int(gset.batch.x.shape[1]), int(self.model.fc2.weight.shape[0])
            if gset.batch.__dict__.get("_pairs") is None and gset.batch.__dict__.get("_edge_ptr32") is None:
                in_place = False
            elif not _lib.load().drk_ginet_step_supported(fi, out, gset.info.max_nodes, gset.info.max_edges) or fi != self.model.conv1.fc.weight.shape[1]:
                in_place = False
        rank, world = (dist.get_rank(), dist.get_world_size()) if self._distributed() else (0, 1)
        return ResidentBatches(gset, batch_size, shuffle, rank=rank, world_size=world, collate=not in_place)

    def _loader(self, dataset, batch_size, shuffle):
        resident = self._resident_loader(dataset, batch_size, shuffle)
        if resident is not None:
            return resident
        rank, world = (dist.get_rank(), dist.get_world_size()) if self._distributed() else (0, 1)
        workers = getattr(self, "_num_workers", 0) or None  # DataLoader's num_workers=0 means "no helpers asked for": pick a default
        return BatchLoader(dataset, batch_size=batch_size, shuffle=shuffle, device=self.device, pin_memory=self.device.type == "cuda", rank=rank, world_size=world,
                           num_workers=workers)

    def train(
        self,
        nepoch: int = 1,
        batch_size: int = 32,
        shuffle: bool = True,
        earlystop_patience: int | None = None,
        earlystop_maxgap: float | None = None,
        min_epoch: int = 10,
        validate: bool = False,
        num_workers: int = 0,  # collate threads of the streamed loader (0 = default: min(4, cores / 2)); unused when the dataset is resident in HBM
        best_model: bool = True,
        filename: str | None = "model.pth.tar",
    ) -> None:
        if self.dataset_train is None:
            raise ValueError("No training dataset provided.")
        self.data_type = type(self.dataset_train)
        self.batch_size_train, self.shuffle = batch_size, shuffle
        self._num_workers = int(num_workers)
        self._fused = None  # rebuilt lazily: the optimizer (and its state tensors) may have changed since the last call
        self.train_loader = self._loader(self.dataset_train, batch_size, shuffle)
        if self.dataset_val is not None:
            self.valid_loader = self._loader(self.dataset_val, batch_size, shuffle)
        else:
            self.valid_loader = None
            _log.warning("Training data will be used both for learning and model selection, which may lead to overfitting.\nIt is usually preferable to use a validation set during the training phase.")

        if self.task == targets.CLASSIF and self.class_weights:
            all_targets = torch.cat([self.dataset_train.get(i).y for i in range(len(self.dataset_train))]).reshape(-1).tolist()
            self.weights = torch.tensor([all_targets.count(c) for c in self.classes], dtype=torch.float32)
            self.weights = 1.0 / self.weights
            self.weights = self.weights / self.weights.sum()
            try:
                self.lossfunction = self.lossfunction(weight=self.weights.to(self.device))
            except TypeError as e:
                text = f"Loss function {self.lossfunction} does not allow for weighted classes.\n\tPlease use a different loss function or set class_weights to False.\n"
                _log.error(text)
                raise ValueError(text) from e
        else:
            self.weights = None

        train_losses, valid_losses = [], []
        saved_model = False
        stopper = EarlyStopping(patience=earlystop_patience, maxgap=earlystop_maxgap, min_epoch=min_epoch, trace_func=_log.info) if (earlystop_patience or earlystop_maxgap) else None

        with self._output_exporters:
            self.nepoch = nepoch
            self._eval(self.train_loader, 0, "training")
            if validate:
                if self.valid_loader is None:
                    raise ValueError("No validation dataset provided.")
                self._eval(self.valid_loader, 0, "validation")
            epoch = 0
            for epoch in range(1, nepoch + 1):
                self.model.train()
                loss_ = self._epoch(epoch, "training")
                train_losses.append(loss_)
                if validate:
                    loss_ = self._eval(self.valid_loader, epoch, "validation")
                    valid_losses.append(loss_)
                    if best_model and _nanmin(valid_losses) == loss_:
                        checkpoint_model = self._save_model()
                        saved_model, self.epoch_saved_model = True, epoch
                    if stopper:
                        stopper(epoch, valid_losses[-1], train_losses[-1])
                        if stopper.early_stop:
                            break
                elif best_model and _nanmin(train_losses) == loss_:
                    checkpoint_model = self._save_model()
                    saved_model, self.epoch_saved_model = True, epoch
            if best_model is False or not saved_model:
                checkpoint_model = self._save_model()
                self.epoch_saved_model = epoch
                if not saved_model and best_model:
                    warnings.warn(
                        "A model has been saved but the validation and/or the training losses were NaN;\n\t"
                        "try to increase the cutoff distance during the data processing or the number of data points during the training.",
                    )
        if filename and (not self._distributed() or dist.get_rank() == 0):
            torch.save(checkpoint_model, filename)
        self.opt_loaded_state_dict = checkpoint_model["optimizer_state"]
        self.model_load_state_dict = checkpoint_model["model_state"]
        self.optimizer.load_state_dict(self.opt_loaded_state_dict)
        self.model.load_state_dict(self.model_load_state_dict)

    def _run_pass(self, loader, epoch_number: int, pass_name: str, train: bool):
        """One pass over ``loader``.  Everything stays on the device until the single readback at the end."""
        loss_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        count = 0
        preds, ys, names = [], [], []
        t0 = time()
        in_place_loader = isinstance(loader, ResidentBatches) and not loader.collate
        for batch, global_size in loader:
            if batch is None:
                # ragged tail: this rank has no graphs but must join the step's collectives -- on the path the other ranks take
                if train and (in_place_loader or self._fused_step(None) is not None):
                    self._ensure_fused().empty_step()
                elif train and self._grad_sync is not None:
                    self.optimizer.zero_grad()
                    self._grad_sync(local_weight=0.0)
                    self.optimizer.step()
                continue
            resident = isinstance(batch, SelectionBatch)
            fused = self._fused_step(batch) if (train and not resident) else None
            if resident:
                # graphs resident in HBM: the batch is a list of ids; step / inference run in place
                from .fused import ginet_infer

                if train:
                    loss_, pred, _ = self._ensure_fused().step_selection(batch.graph_set, prepared=batch.prepared, global_size=global_size)
                    if global_size != pred.shape[0]:  # the kernel scales by the global batch: back to this rank's mean
                        loss_ = loss_ * (global_size / pred.shape[0])  # (else: the step's own loss buffer, consumed below before the next step overwrites it)
                    pred, y = self._format_output(pred.clone(), batch.y)
                else:
                    with torch.no_grad():
                        pred = ginet_infer(self.model, batch.graph_set.batch, selection=batch.prepared[0])
                        pred, y = self._format_output(pred, batch.y)
                        loss_ = self.lossfunction(pred, y) if y is not None else None
            elif fused is not None:
                if getattr(loader, "only", 0) is None:
                    loader.only = type(fused).FIELDS  # later batches: copy only what the step kernels read (the rest stays lazy)
                # whole step (index, forward, loss, backward, gradient all-reduce, optimizer) in the per-graph kernels
                loss_, pred = fused(batch, global_size=global_size)
                loss_ = loss_ * (global_size / pred.shape[0])  # the kernel scales by the global batch: back to this rank's mean
                pred, y = self._format_output(pred.clone(), batch.y)
            elif train:
                self.optimizer.zero_grad()
                pred = self.model(batch)
                pred, y = self._format_output(pred, batch.y)
                loss_ = self.lossfunction(pred, y)
                loss_.backward()
                if self._grad_sync is not None:
                    self._grad_sync(local_weight=pred.shape[0] / global_size)
                self.optimizer.step()
            else:
                with torch.no_grad():
                    pred = self.model(batch)
                    pred, y = self._format_output(pred, batch.y)
                    loss_ = self.lossfunction(pred, y) if y is not None else None
            n_here = pred.shape[0]
            if y is not None:
                loss_sum.add_(loss_.detach(), alpha=float(n_here))  # "convert mean back to sum" (trainer.py:694), accumulated in float64
                count += n_here
                ys.append(y.detach())
            else:
                ys.append(None)
            preds.append(softmax(pred.detach(), dim=1) if self.task == targets.CLASSIF else pred.detach().reshape(-1))
            names += batch.entry_names if isinstance(batch.entry_names, list) else [batch.entry_names]
        # ---- the one host round trip of the pass
        outputs = torch.cat(preds).cpu().numpy().tolist() if preds else []
        target_vals = []
        for y, p in zip(ys, preds):
            target_vals += y.cpu().numpy().tolist() if y is not None else [None] * p.shape[0]
        bad_targets = getattr(self, "_bad_targets", None)
        self._bad_targets = None
        if self._distributed():
            # every rank must see the SAME epoch loss: it drives best-model selection and early stopping, and ranks that disagree
            # would leave the epoch loop at different times (the others then hang in the gradient exchange)
            tot = torch.stack([loss_sum, torch.tensor(float(count), dtype=torch.float64, device=self.device)])
            dist.all_reduce(tot)
            total_loss, total_count = (float(v) for v in tot.tolist())
            epoch_loss = total_loss / total_count if total_count > 0 else None
            # the exporters get the whole pass on rank 0 (every rank holds only its slices of the mini-batches)
            gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
            dist.gather_object((names, outputs, target_vals), gathered, dst=0)
            if gathered is not None:
                names, outputs, target_vals = ([v for part in gathered for v in part[i]] for i in range(3))
        else:
            epoch_loss = float(loss_sum.item()) / count if count > 0 else None
        if self.device.type == "cuda":
            from .utils.community_pooling import check_status as check_pool_status

            check_pool_status(self.device)  # the pooling kernels' status words, accumulated on the device during the pass
        if bad_targets is not None and bool(bad_targets.item()):
            raise ValueError(f"a target value of the {pass_name} pass is not one of the dataset's classes {self.classes} (trainer.py:812 raises KeyError there)")
        if not self._distributed() or dist.get_rank() == 0:
            self._output_exporters.process(pass_name, epoch_number, names, outputs, target_vals, epoch_loss)
        _log.info(f"{pass_name} loss {epoch_loss} | time {time() - t0}")
        return epoch_loss

    def _ensure_fused(self):
        from .fused import GINetFusedStep

        if self._fused is None or self._fused is False:
            world = dist.get_world_size() if self._distributed() else 1
            self._fused = GINetFusedStep(self.model, self.optimizer, self.lossfunction, target_fn=lambda b: self._format_output(None, b.y)[1], world_size=world)
        return self._fused

    def _fused_step(self, batch):
        """The per-graph step kernels (``fused.GINetFusedStep``) when the model is the reference ``ginet_nocluster.GINet``, the
        loss is MSELoss / unweighted CrossEntropyLoss and every graph of ``batch`` fits the kernel's plan; else None (the
        layer kernels through autograd)."""
        from .fused import GINetFusedStep, step_supported

        if self._fused is None:
            loss_fn = self.lossfunction
            if isinstance(loss_fn, type) or not GINetFusedStep.supports(self.model, loss_fn):
                self._fused = False
            else:
                world = dist.get_world_size() if self._distributed() else 1
                self._fused = GINetFusedStep(self.model, self.optimizer, loss_fn, target_fn=lambda b: self._format_output(None, b.y)[1], world_size=world)
        if self._fused is False:  # decided by model and loss alone: the same on every rank
            return None
        ok = batch is None or step_supported(self.model, batch)
        if self._distributed():
            # the choice must be collective: the fused path's in-kernel gradient exchange and the autograd path's NCCL all-reduce are
            # different protocols, and this rank's shard may hold a graph that does not fit the kernel's plan while the others' do not
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        return self._fused if ok else None

    def _epoch(self, epoch_number: int, pass_name: str):
        return self._run_pass(self.train_loader, epoch_number, pass_name, train=True)

    def _eval(self, loader, epoch_number: int, pass_name: str):
        self.model.eval()
        return self._run_pass(loader, epoch_number, pass_name, train=False)

    def _format_output(self, pred, target=None):
        """regress: ``pred.reshape(-1)``; classif: targets -> class indices (``trainer.py:807-835``), done on the device
        with one comparison against the class table instead of a Python loop + host tensor."""
        if self.task == targets.CLASSIF and target is not None:
            table = getattr(self, "_class_table", None)
            if table is None or table.device != target.device:
                table = torch.tensor([float(c) for c in self.classes], device=target.device)
                self._class_table = table
            match = target.reshape(-1, 1).to(table.dtype) == table.reshape(1, -1)
            missing = (~match.any(dim=1)).any()  # a value outside `classes`: the reference's dict lookup raises; checked at the pass's read-back
            prev = getattr(self, "_bad_targets", None)
            self._bad_targets = missing if prev is None else (prev | missing)
            target = match.to(torch.int64).argmax(dim=1)
            if isinstance(self.lossfunction, (nn.BCELoss, nn.BCEWithLogitsLoss)):
                raise ValueError("BCELoss and BCEWithLogitsLoss are currently not supported.\n\tFor further details see: https://github.com/DeepRank/deeprank2/issues/318")
            if isinstance(self.lossfunction, losses.classification_losses) and not isinstance(self.lossfunction, losses.classification_tested):
                raise ValueError(f"{self.lossfunction} is currently not supported.\n\tSupported loss functions for classification: {losses.classification_tested}.")
        elif self.task == targets.REGRESS and pred is not None:
            pred = pred.reshape(-1)
        if target is not None:
            target = target.to(self.device)
        return pred, target

    def test(self, batch_size: int = 32, num_workers: int = 0) -> None:  # noqa: ARG002
        if (not self.pretrained_model) and (not self.model_load_state_dict):
            raise ValueError("No pretrained model provided and no training performed. Please provide a pretrained model or train the model before testing.")
        self.batch_size_test = batch_size
        if self.dataset_test is None:
            _log.error("No test dataset provided.")
            raise ValueError("No test dataset provided.")
        self.test_loader = self._loader(self.dataset_test, batch_size, False)
        with self._output_exporters:
            self._eval(self.test_loader, self.epoch_saved_model, "testing")

    # ------------------------------------------------------------------ checkpoints (same keys as trainer.py:926-956)
    _STATE_KEYS = (
        "target", "target_transform", "task", "classes", "classes_to_index", "class_weights", "batch_size_train", "batch_size_test",
        "val_size", "test_size", "lr", "weight_decay", "epoch_saved_model", "subset", "shuffle", "clustering_method", "node_features",
        "edge_features", "features", "means", "devs", "cuda", "ngpu",
    )

    def _save_model(self) -> dict[str, Any]:
        transforms = copy.deepcopy(self.features_transform)
        if transforms:
            for spec in transforms.values():
                if spec.get("transform") is None or isinstance(spec["transform"], str):
                    continue
                source = inspect.getsource(spec["transform"])
                found = re.search(r"[\"|\']transform[\"|\']:.*(lambda.*).*,.*[\"|\']standardize[\"|\'].*", source)
                spec["transform"] = found.group(1) if found else source.strip()
        state = {
            "data_type": self.data_type,
            "model_state": copy.deepcopy(self.model.state_dict()),
            "optimizer": self.optimizer,
            "optimizer_state": copy.deepcopy(self.optimizer.state_dict()),
            "lossfunction": self.lossfunction,
            "features_transform": transforms,
        }
        for key in self._STATE_KEYS:
            state[key] = getattr(self, key)
        return state

    def _load_params(self) -> None:
        state = torch.load(self.pretrained_model, map_location=None if torch.cuda.is_available() else torch.device("cpu"), weights_only=False)
        self.data_type = state["data_type"]
        self.model_load_state_dict = state["model_state"]
        self.optimizer = type(state["optimizer"])
        self.opt_loaded_state_dict = state["optimizer_state"]
        self.lossfunction = state["lossfunction"]
        self.features_transform = state["features_transform"]
        for key in self._STATE_KEYS:
            setattr(self, key, state[key])
        self._select_device()

    def _load_pretrained_model(self) -> None:
        self._put_model_to_device(self.dataset_test)  # the model first: the loader asks it which step kernels apply
        self.test_loader = self._loader(self.dataset_test, 1, False)
        self.optimizer = self.optimizer(self.model.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        self.optimizer.load_state_dict(self.opt_loaded_state_dict)
        self.model.load_state_dict(self.model_load_state_dict)


def _nanmin(values):
    clean = [v for v in values if v is not None and v == v]
    return min(clean) if clean else None


def _divide_dataset(dataset: GraphDataset, splitsize: float | int | None = None):
    """Random split into (main, split) datasets sharing the parent's settings (``trainer.py:961-1004``)."""
    if splitsize is None:
        splitsize = 0.25
    full_size = len(dataset)
    if isinstance(splitsize, float):
        n_split = int(splitsize * full_size)
    elif isinstance(splitsize, int):
        n_split = splitsize
    else:
        raise TypeError(f"type(splitsize) must be float, int or None ({type(splitsize)} detected.)")
    if n_split >= full_size or n_split < 0:
        raise ValueError(f"Invalid Split size: {n_split}.\nSplit size must be a float between 0 and 1 OR an int smaller than the size of the dataset ({full_size} datapoints)")
    if splitsize == 0:
        return dataset, None
    indices = np.arange(full_size)
    seed = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [int(np.random.SeedSequence().entropy % (2**63)) if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)  # every rank must draw the same train / validation split
        seed = box[0]
    np.random.default_rng(seed).shuffle(indices)
    main, split = copy.copy(dataset), copy.copy(dataset)
    main.index_entries = [dataset.index_entries[i] for i in indices[n_split:]]
    split.index_entries = [dataset.index_entries[i] for i in indices[:n_split]]
    return main, split
