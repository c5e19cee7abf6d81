"""Comparison rules of the parity tests.

Integer tracks and peaks: exact.  Float tracks (lambda, p, q): compared as
step functions on the union of both sides' breakpoints, within
    |a - b| <= RTOL * |b| + ATOL,   RTOL = 1e-6 (north_star), ATOL = 1e-12
(the absolute floor only matters for p-values below 1e-6, where the
reference's own  log(1 - cdf)  is quantised to multiples of 2^-53).
Run boundaries of float tracks may differ where neighbouring values differ by
ulps (the reference accumulates +-weight steps sequentially in float64, the
kernels multiply an integer count once), so boundaries are not compared.
"""
import numpy as np

RTOL = 1e-6
ATOL = 1e-12


def step_at(idx, val, query):
    j = np.searchsorted(idx, query, side="right") - 1
    return np.where(j >= 0, np.asarray(val)[np.maximum(j, 0)], 0.0)


def assert_step_close(a, b, what="", rtol=RTOL, atol=ATOL, upto=None):
    """a, b: (indices, values); b is the oracle."""
    ai, av = np.asarray(a[0]), np.asarray(a[1], dtype=np.float64)
    bi, bv = np.asarray(b[0]), np.asarray(b[1], dtype=np.float64)
    q = np.union1d(ai, bi)
    if upto is not None:
        q = q[q < upto]
    va, vb = step_at(ai, av, q), step_at(bi, bv, q)
    err = np.abs(va - vb)
    tol = rtol * np.abs(vb) + atol
    bad = np.flatnonzero(err > tol)
    assert bad.size == 0, "%s: %d/%d positions differ, first at %s: got %r want %r" % (
        what, bad.size, q.size, q[bad[0]], va[bad[0]], vb[bad[0]])


def assert_exact(a, b, what=""):
    assert np.array_equal(np.asarray(a[0]), np.asarray(b[0])), what + " indices"
    assert np.array_equal(np.asarray(a[1]), np.asarray(b[1])), what + " values"
