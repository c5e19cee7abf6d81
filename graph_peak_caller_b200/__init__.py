"""B200-native callpeaks core behind graph_peak_caller's Python API
(reference graph_peak_caller/__init__.py:1 exports the same three names).

Importing the package is cheap (numpy only); the CUDA library
(libgpc_b200.so, built by ``__graft_entry__.build()``) is loaded on first use
and there is no CPU fallback."""
from .callpeaks import CallPeaks, CallPeaksFromQvalues, Configuration  # noqa: F401
