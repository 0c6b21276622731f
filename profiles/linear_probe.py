"""drk_node_linear: tensor-core + bulk-copy kernel (default for row-contiguous operands) vs the SIMT kernel (DRK_LINEAR_TC=0), C2 / C3 shapes.
    gpurun -- python profiles/linear_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deeprank2_b200 import ops

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t_us(fn, reps=20):
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(c)
    return 1e3 * tot / reps


for n, k, m in ((77469, 50, 64), (77469, 50, 32), (77469, 64, 50), (193434, 38, 32), (193434, 16, 64)):
    a = torch.randn(n, k, device=dev)
    w = torch.randn(m, k, device=dev)
    out = torch.empty(n, m, device=dev)
    by = 4 * n * (k + m)
    res = {}
    for mode in ("1", "0"):
        os.environ["DRK_LINEAR_TC"] = mode
        f = lambda: ops.node_linear(a, w, True, out=out)  # noqa: E731
        f()
        res[mode] = t_us(f)
    print(f"N {n} K {k} M {m}: tensor cores + bulk copy {res['1']:7.1f} us ({by / res['1'] / 1e3 / 6544.3:5.3f} of HBM peak)   SIMT + cp.async {res['0']:7.1f} us ({by / res['0'] / 1e3 / 6544.3:5.3f})")
