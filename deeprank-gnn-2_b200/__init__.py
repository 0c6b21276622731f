"""B200-native message-passing hot path for DeepRank2-style GNNs.

Host side: Python/PyTorch mirroring the reference's ``torch.nn.Module`` /
``Trainer`` / ``GraphDataset`` API for this path.  Device side: hand-written sm_100a
CUDA kernels behind the C ABI declared in ``include/drk_b200.h`` (``csrc/``), loaded
with ``ctypes``.  There is no CPU fallback: every op raises if the extension is
missing or a tensor is not on a CUDA device.
"""
__version__ = "0.1.0"
