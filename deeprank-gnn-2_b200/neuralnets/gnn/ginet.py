"""Clustered ``GINet`` (mirror of ``deeprank2/neuralnets/gnn/ginet.py``).

``GINetConvLayer`` (``ginet.py:13-63``) is shared with the no-cluster variant.
"""
from __future__ import annotations

from ._common import GINetConvLayer  # noqa: F401
