"""iifb200 — B200-native clique belief-convolution hot path of IncrementalInference.jl.

Host-side mirror of the reference interface (Python) above the C-ABI of libiifb200.so.
"""
from . import _abi, compile, graph  # noqa: F401
from .graph import *  # noqa: F401,F403
