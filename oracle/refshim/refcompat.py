"""TEST INFRASTRUCTURE ONLY -- make /root/reference importable on numpy 2 / scipy 1.18.

Usage (build container only; /root/reference does not exist on the GPU box):

    import refcompat; refcompat.install()
    import graph_peak_caller            # the unmodified reference
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("GPC_REFERENCE_ROOT", "/root/reference")


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
