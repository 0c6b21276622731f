"""pytest configuration: the ``gpu`` marker and shared golden-vector helpers."""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["toy_edgecases", "toy_edgecases_fe3", "synthetic_small", "fixture_1ATN", "fixture_variants_fe5"]
CLUSTERED_CASES = ["synthetic_small", "fixture_1ATN", "fixture_variants_fe5"]

# fp32 tolerance of the path: allclose(rtol=1e-5, atol=1e-5 * max|ref|)
RTOL = 1e-5
ATOL_SCALE = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """View of one ``tests/golden/<case>.npz`` (key scheme in oracle/make_golden.py)."""

    def __init__(self, case: str):
        self.case = case
        self._z = np.load(os.path.join(GOLDEN_DIR, f"{case}.npz"))

    def has(self, key: str) -> bool:
        return key in self._z.files

    def t(self, key: str) -> torch.Tensor:
        return torch.from_numpy(self._z[key].copy())

    def group(self, prefix: str) -> dict:
        prefix = prefix.rstrip("/") + "/"
        return {k[len(prefix):]: self.t(k) for k in self._z.files if k.startswith(prefix)}

    def inputs(self) -> SimpleNamespace:
        return SimpleNamespace(**self.group("in"))


def load_golden(case: str) -> Golden:
    return Golden(case)


def assert_close(actual: torch.Tensor, expected: torch.Tensor, what: str = "", rtol: float = RTOL, atol_scale: float = ATOL_SCALE):
    """fp32 parity bar: rtol 1e-5 with atol = 1e-5 * max|expected| (NaN positions must coincide)."""
    actual = actual.detach().cpu()
    expected = expected.detach().cpu()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} != {tuple(expected.shape)}"
    nan_e = torch.isnan(expected)
    assert torch.equal(torch.isnan(actual), nan_e), f"{what}: NaN pattern differs"
    finite_e = expected[~nan_e]
    scale = float(finite_e.abs().max()) if finite_e.numel() else 0.0
    atol = atol_scale * scale
    a = actual[~nan_e].double()
    e = finite_e.double()
    err = (a - e).abs()
    bad = err > (atol + rtol * e.abs())
    if bool(bad.any()):
        worst = float(err.max())
        raise AssertionError(f"{what}: {int(bad.sum())}/{e.numel()} outside rtol={rtol} atol={atol:.3e}; max|d|={worst:.3e} max|ref|={scale:.3e}")


def assert_equal_int(actual: torch.Tensor, expected: torch.Tensor, what: str = ""):
    actual = actual.detach().cpu()
    expected = expected.detach().cpu()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} != {tuple(expected.shape)}"
    assert torch.equal(actual.to(torch.int64), expected.to(torch.int64)), f"{what}: integer arrays differ"


ADAM_LR, ADAM_EPS, ADAM_WD = 1e-3, 1e-8, 1e-5


def assert_adam_close(actual: torch.Tensor, expected: torch.Tensor, what: str = "", grad_ref: torch.Tensor | None = None, w_before: torch.Tensor | None = None):
    """Weights after ONE Adam(lr 1e-3, wd 1e-5) step.

    The first Adam step is dw = -lr * g' / (|g'| + eps), g' = g + wd * w: a sign-like function whose slope at
    g' ~ 0 is lr / eps = 1e5.  A gradient that is within the fp32 parity bar (|dg| <= 1e-5 max|g| + 1e-5 |g|) can
    therefore move the updated weight by up to  lr * eps * |dg| / (|g'| + eps)^2  (capped at 2 lr).  With the
    reference gradient at hand the bar is exactly that propagated bound; without it, 5% of lr."""
    actual = actual.detach().cpu().double()
    expected = expected.detach().cpu().double()
    assert actual.shape == expected.shape, what
    err = (actual - expected).abs()
    if grad_ref is not None and w_before is not None:
        g = grad_ref.detach().cpu().double() + ADAM_WD * w_before.detach().cpu().double()
        dg = ATOL_SCALE * float(grad_ref.abs().max()) + RTOL * g.abs()
        slack = ADAM_LR * torch.clamp(ADAM_EPS * dg / (g.abs() + ADAM_EPS) ** 2, max=2.0)
    else:
        slack = torch.full_like(err, 0.05 * ADAM_LR)
    bad = err > (slack + RTOL * expected.abs() + 1e-9)
    assert not bool(bad.any()), f"{what}: {int(bad.sum())}/{err.numel()} weights outside the propagated gradient tolerance; max|d|={float(err.max()):.3e}"


@pytest.fixture(autouse=True)
def _fresh_status_words():
    """The per-device status word of the per-graph kernels is shared by all batches (fused._status_word): a test that provokes a flag
    and reads the word directly must not leak it into the next test."""
    yield
    try:
        import sys

        fused = sys.modules.get("deeprank2_b200.fused")
        if fused is not None:
            for word in fused._STATUS_WORDS.values():
                word.zero_()
        cp = sys.modules.get("deeprank2_b200.utils.community_pooling")
        if cp is not None:
            for word in cp._STATUS.values():
                word.zero_()
    except Exception:  # noqa: BLE001 - never let the cleanup fail a test
        pass
