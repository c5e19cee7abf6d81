"""TEST INFRASTRUCTURE ONLY -- make /root/reference importable on numpy 2 / scipy 1.18.

Usage (build container: /root/reference; GPU box: oracle/_ref, see oracle/build_ref.py):

    import refcompat; refcompat.install()
    import graph_peak_caller            # the unmodified reference
"""
import os
import sys

_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref")


def _default_root():
    """/root/reference in the build container; on the GPU box the copy that
    oracle/build_ref.py pip-installed into oracle/_ref (git-ignored, travels)."""
    if os.path.isdir("/root/reference/graph_peak_caller"):
        return "/root/reference"
    return _INSTALLED


REFERENCE_ROOT = os.environ.get("GPC_REFERENCE_ROOT") or _default_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "graph_peak_caller"))


def install():
    import numpy as np
    import scipy
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # API removals the reference still uses (sparsepvalues.py:20,124; logging_config.py:55)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if not hasattr(np, "float"):
            np.float = float
        # legacy/sparsepileup.py:48 (only used by the reference's tests to build inputs)
        if not hasattr(np, "int"):
            np.int = int
    if not hasattr(scipy, "errstate"):
        scipy.errstate = np.errstate
    if not hasattr(scipy, "seterr"):
        scipy.seterr = np.seterr
