"""TEST INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference package into ``oracle/_ref/``.

    python -m oracle.install_reference

``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>`` (the
source tree is read-only, so the build runs from a copy under /tmp; ``--no-deps`` because torch_geometric,
torch_scatter, h5py ... are not in the offline wheelhouse -- ``oracle/thirdparty.py`` stands in for the symbols the
model files import).  ``oracle/_ref/`` is git-ignored and NOT gpurun-ignored: it travels to the GPU box, where
``bench.py --impl reference`` and the ``gpu_torch_baseline`` leg execute the reference's own
``deeprank2/neuralnets/gnn/*.py`` through ``oracle/reference_loader.py``.  Nothing under it is product source and
nothing in the product package imports it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "deeprank2", "neuralnets", "gnn", "ginet_nocluster.py"))


def install(force: bool = False) -> bool:
    """True if oracle/_ref holds the reference afterwards.  No-op when it is already there or when /root/reference is absent
    (the GPU box: only the prebuilt copy is used)."""
    if installed() and not force:
        return True
    if not os.path.isdir(SOURCE):
        return False
    tmp = tempfile.mkdtemp(prefix="drk_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "tests", "docs", "tutorials", "paper"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"installing the reference into oracle/_ref failed:\n{res.stdout}\n{res.stderr}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


if __name__ == "__main__":
    print("oracle/_ref installed" if install(force="--force" in sys.argv) else "reference tree not available: nothing installed")
