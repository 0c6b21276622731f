"""CPU emulation of the 3xTF32 product (hi/lo split, three MMAs per product, fp32 accumulate) for the dense shapes of the layer
kernels, against the path's parity bar (rtol 1e-5, atol 1e-5 * max|ref|, reference = fp32 sgemm).  Two accumulator models:
round-to-nearest and round-toward-zero (the worst case for a tensor-core accumulator).  Result (numpy, seed 0): max error
<= 1.3e-6 * max|ref| and no element outside the bar for K = 50..128 -- a tensor-core `drk_node_linear` is numerically viable; the
attempt dropped earlier in the round failed a Vanilla parity test for another reason.
usage: python profiles/tf32_emulation.py"""
import numpy as np

rng = np.random.default_rng(0)


def tf32(x):  # cvt.rna.tf32.f32: round to nearest, ties away, 10 explicit mantissa bits
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def add_rz(acc, p):  # fp32 add rounded toward zero
    s = acc.astype(np.float64) + p.astype(np.float64)
    f = s.astype(np.float32)
    away = np.abs(f.astype(np.float64)) > np.abs(s)
    return np.where(away, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def gemm_3xtf32(a, b, rz):
    ah, bh = tf32(a), tf32(b)
    al, bl = tf32(a - ah), tf32(b - bh)
    acc = np.zeros((a.shape[0], b.shape[0]), np.float32)
    for k0 in range(0, a.shape[1], 8):  # one m16n8k8 step: the two small products first, then the main term
        for x, y in ((al, bh), (ah, bl), (ah, bh)):
            p = (x[:, k0 : k0 + 8].astype(np.float64) @ y[:, k0 : k0 + 8].astype(np.float64).T).astype(np.float32)
            acc = add_rz(acc, p) if rz else (acc + p).astype(np.float32)
    return acc


for n, k, m, note in ((4096, 50, 64, "Vanilla uv"), (4096, 82, 50, "Vanilla node mlp"), (4096, 64, 50, "backward dX"), (4096, 128, 64, "wide")):
    a = (rng.standard_normal((n, k)) * 3.0).astype(np.float32)
    b = (rng.uniform(-1, 1, (m, k)) / np.sqrt(k)).astype(np.float32)
    ref = (a @ b.T).astype(np.float32)
    line = []
    for name, c in (("RN", gemm_3xtf32(a, b, False)), ("RZ", gemm_3xtf32(a, b, True))):
        err = np.abs(c.astype(np.float64) - ref.astype(np.float64))
        tol = 1e-5 * np.abs(ref).max() + 1e-5 * np.abs(ref)
        line.append(f"{name}: max err {err.max() / np.abs(ref).max():.2e} of max|ref|, {int((err > tol).sum())} outside the bar")
    print(f"{note:18s} K={k:3d} M={m:3d}  " + "   ".join(line))
