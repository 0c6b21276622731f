"""CPU: the driver-facing contract of ``bench.py`` that can be checked without a GPU -- the reference arm prints exactly ONE JSON
line on stdout (library chatter goes to stderr) with the keys the driver reads, and our arm refuses to run without a CUDA device
instead of falling back to anything."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-graphs", "4")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ginet_train_step_graphs_per_s" and d["unit"] == "graphs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    # "reference" = the unmodified deeprank2 module executed from oracle/_ref (or /root/reference); "port" only when neither exists
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "4-graph" in d["cpu_baseline"]["sample"] and d["config"]["graphs_per_step"] == 4, "the line must describe what actually ran"
    assert d["e2e"] == {"value": d["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_our_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box this arm is exercised by the driver itself
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0, "bench.py must fail loudly without a CUDA device (no CPU fallback)"
    assert not r.stdout.strip(), r.stdout
