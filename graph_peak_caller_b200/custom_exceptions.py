"""Error types of the reference API (graph_peak_caller/custom_exceptions.py)."""


class GraphNotFoundException(Exception):
    pass


class InvalidPileupInterval(Exception):
    """A read names a node that is not part of the graph
    (reference sample/sparsegraphpileup.py:83-85)."""


class IntervalNotInGraphException(Exception):
    """A vg alignment names a node that is not in the graph it is read against
    (pyvg.vgobjects.IntervalNotInGraphException, caught by the reference at
    command_line_interface.py:701)."""


class DeviceLibraryMissing(RuntimeError):
    """libgpc_b200.so (the CUDA kernels) is not built or no CUDA device is
    visible.  There is no CPU fallback on this path."""
