"""Per-epoch output sinks of the Trainer (host side; out of scope for kernels).

Same hook as the reference (``deeprank2/utils/exporters.py:16-87``): the Trainer enters the collection once per
``train()`` / ``test()`` and calls ``process(pass_name, epoch, entry_names, outputs, targets, loss)`` once per pass.
``HDF5OutputExporter`` keeps the reference's table (columns phase/epoch/entry/output/target/loss); it is written with
``pandas.to_hdf`` when PyTables is installed and as CSV otherwise (this image has no PyTables / h5py).
"""
from __future__ import annotations

import os

import pandas as pd


class OutputExporter:
    def __init__(self, directory_path: str | None = None):
        self._directory_path = directory_path or "./output"
        os.makedirs(self._directory_path, exist_ok=True)

    def __enter__(self):
        return self

    def __exit__(self, exception_type, exception, traceback):
        return None

    def process(self, pass_name: str, epoch_number: int, entry_names: list, output_values: list, target_values: list, loss: float) -> None:
        """entry_names, output_values and target_values have the same length."""

    def is_compatible_with(self, output_data_shape: int, target_data_shape: int | None = None) -> bool:  # noqa: ARG002
        return True


class OutputExporterCollection:
    def __init__(self, *exporters: OutputExporter):
        self._output_exporters = exporters

    def __enter__(self):
        for e in self._output_exporters:
            e.__enter__()
        return self

    def __exit__(self, exception_type, exception, traceback):
        for e in self._output_exporters:
            e.__exit__(exception_type, exception, traceback)

    def process(self, pass_name, epoch_number, entry_names, output_values, target_values, loss) -> None:
        for e in self._output_exporters:
            e.process(pass_name, epoch_number, entry_names, output_values, target_values, loss)

    def __iter__(self):
        return iter(self._output_exporters)


class HDF5OutputExporter(OutputExporter):
    """Every data point of every pass: phase, epoch, entry, output, target, loss."""

    COLUMNS = ("phase", "epoch", "entry", "output", "target", "loss")

    def __init__(self, directory_path: str):
        self.phase = None
        super().__init__(directory_path)

    def __enter__(self):
        self.df = pd.DataFrame({c: [] for c in self.COLUMNS})
        return self

    def __exit__(self, exception_type, exception, traceback):
        if self.phase is None:
            return
        key = "training" if self.phase == "validation" else self.phase
        try:
            import tables  # noqa: F401

            self.df.to_hdf(os.path.join(self._directory_path, "output_exporter.hdf5"), key=key, mode="a")
        except ImportError:
            self.df.to_csv(os.path.join(self._directory_path, f"output_exporter_{key}.csv"), index=False)

    def process(self, pass_name, epoch_number, entry_names, output_values, target_values, loss) -> None:
        self.phase = pass_name
        n = len(output_values)
        block = pd.DataFrame({"phase": [pass_name] * n, "epoch": [epoch_number] * n, "entry": entry_names, "output": output_values, "target": target_values, "loss": [loss] * n})
        self.df = pd.concat([self.df, block]).reset_index(drop=True)
