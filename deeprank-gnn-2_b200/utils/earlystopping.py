"""Early stopping on the validation loss (same constructor / call contract as ``deeprank2/utils/earlystopping.py``)."""
from __future__ import annotations

from collections.abc import Callable


class EarlyStopping:
    """Stop when the validation loss has not improved by more than ``delta`` for ``patience`` epochs, or when, after
    ``min_epoch`` epochs, it exceeds the training loss by more than ``maxgap``."""

    def __init__(self, patience: int = 10, delta: float = 0, maxgap: float | None = None, min_epoch: int = 10, verbose: bool = True, trace_func: Callable = print):
        self.patience, self.delta, self.maxgap, self.min_epoch = patience, delta, maxgap, min_epoch
        self.verbose, self.trace_func = verbose, trace_func
        self.early_stop = False
        self.counter = 0
        self.best_score = None
        self.val_loss_min = None

    def _say(self, text: str) -> None:
        if self.verbose:
            self.trace_func(text)

    def __call__(self, epoch: int, val_loss: float, train_loss: float | None = None):
        score = -val_loss
        if self.best_score is None:
            self.best_score, self.val_loss_min = score, val_loss
        elif score < self.best_score + self.delta:
            self.counter += 1
            margin = f"more than {self.delta} " if self.delta else ""
            self._say(f"Validation loss did not decrease {margin}({self.val_loss_min:.6f} --> {val_loss:.6f}). EarlyStopping counter: {self.counter} out of {self.patience}")
            if self.patience is not None and self.counter >= self.patience:
                self.trace_func(f"EarlyStopping activated at epoch # {epoch} because patience of {self.patience} has been reached.")
                self.early_stop = True
        else:
            self._say(f"Validation loss decreased ({self.val_loss_min:.6f} --> {val_loss:.6f}).")
            self.best_score, self.counter = score, 0
        if score >= self.best_score:
            self.best_score, self.val_loss_min = score, val_loss
        if self.maxgap and epoch > self.min_epoch:
            if train_loss is None:
                raise ValueError("Cannot compute gap because no train_loss is provided to EarlyStopping.")
            gap = val_loss - train_loss
            if gap > self.maxgap:
                self.trace_func(f"EarlyStopping activated at epoch # {epoch} due to overfitting. The difference between validation and training loss of {gap} exceeds the maximum allowed ({self.maxgap})")
                self.early_stop = True
