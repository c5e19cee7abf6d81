"""ORACLE (test infrastructure, CPU, numpy) -- run-length track arithmetic.

Restates graph_peak_caller/sparsediffs.py of the reference (v1.2.3) on plain
numpy arrays.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
leg may import this package; the product never does.

Parity status: PINNED -- checked bit-for-bit against the reference itself
(imported through oracle/refshim) by tests/test_oracle_vs_reference.py and
against the committed golden vectors in tests/golden/.
"""
import numpy as np


def canonical_runs(indices, values):
    """SparseValues._sanitize (sparsediffs.py:13-20): of entries sharing an
    index keep the last, then drop entries repeating the previous value."""
    indices = np.asarray(indices)
    values = np.asarray(values)
    if indices.size == 0:
        return indices, values
    last_of_index = np.r_[indices[1:] != indices[:-1], True]
    indices, values = indices[last_of_index], values[last_of_index]
    new_value = np.r_[True, values[1:] != values[:-1]]
    return indices[new_value], values[new_value]


def diffs_to_runs(indices, diffs):
    """SparseDiffs.get_sparse_values (sparsediffs.py:137-141): stable sort by
    index, running sum, canonicalise."""
    order = np.argsort(indices, kind="mergesort")
    return canonical_runs(np.asarray(indices)[order],
                          np.cumsum(np.asarray(diffs)[order]))


def runs_to_diffs(indices, values):
    """SparseDiffs.clean second half (sparsediffs.py:124-128)."""
    return indices, np.ediff1d(values, to_begin=values[0])


def drop_zero_diffs(indices, diffs):
    """SparseDiffs._sanitize (sparsediffs.py:161-165)."""
    keep = diffs != 0
    return indices[keep], diffs[keep]


def starts_ends_to_diffs(starts_ends):
    """SparseDiffs.from_starts_and_ends (sparsediffs.py:171-178): row 0 are
    +1 positions, row 1 are -1 positions."""
    flat = starts_ends.ravel(order="C")
    # default (unstable) sort, exactly as the reference: the order of equal
    # indices only changes float rounding of intermediate sums
    order = np.argsort(flat)
    diffs = np.ones(flat.size)
    diffs[order >= starts_ends.shape[1]] = -1.0
    return flat[order], diffs


def clip_min_diffs(indices, diffs, min_value):
    """SparseDiffs.clip_min (sparsediffs.py:130-135)."""
    values = np.cumsum(diffs)
    values = np.maximum(values, min_value)
    out = np.empty_like(values)
    out[0] = values[0]
    out[1:] = np.diff(values)
    return drop_zero_diffs(indices, out)


def _merge_two(idx_a, diffs_a, idx_b, diffs_b):
    """Shared front half of SparseDiffs.maximum / apply_binary_func
    (sparsediffs.py:143-154, 208-219): stable merge of both index lists, each
    track's running sum taken along the merged axis."""
    all_idx = np.r_[idx_a, idx_b]
    order = np.argsort(all_idx, kind="mergesort")
    merged = all_idx[order]
    last_of_index = np.ediff1d(merged, to_end=1) != 0
    mine = order < idx_a.size
    d = np.zeros((2, all_idx.size))
    d[0][mine] = diffs_a
    d[1][~mine] = diffs_b
    vals = np.cumsum(d, axis=1)
    return merged[last_of_index], vals, last_of_index


def maximum_diffs(idx_a, diffs_a, idx_b, diffs_b):
    """SparseDiffs.maximum (sparsediffs.py:143-159)."""
    idx, vals, keep = _merge_two(idx_a, diffs_a, idx_b, diffs_b)
    top = np.max(vals, axis=0)[keep]
    return drop_zero_diffs(idx, np.ediff1d(top, to_begin=top[0]))


def binary_func_runs(idx_a, diffs_a, idx_b, diffs_b, func):
    """SparseDiffs.apply_binary_func(..., return_values=True)
    (sparsediffs.py:208-222)."""
    idx, vals, keep = _merge_two(idx_a, diffs_a, idx_b, diffs_b)
    return canonical_runs(idx, func(vals[0], vals[1])[keep])


def threshold_runs(indices, values, cutoff):
    """SparseValues.threshold_copy (sparsediffs.py:58-62)."""
    return canonical_runs(indices, values >= cutoff)


def runs_to_dense(indices, values, size):
    """SparseValues.to_dense_pileup (sparsediffs.py:64-76)."""
    values = np.asarray(values)
    is_bool = values.dtype == bool
    if is_bool:
        values = values.astype(np.int64)
    diffs = np.ediff1d(values, to_begin=values[0])
    dense = np.zeros(size + 1, dtype=values.dtype)
    dense[np.asarray(indices)[:diffs.size]] = diffs
    dense = np.cumsum(dense[:-1])
    return dense.astype(bool) if is_bool else dense


def dense_to_runs(dense):
    """SparseValues.from_dense_pileup (sparsediffs.py:78-85)."""
    change = np.flatnonzero(dense[1:] != dense[:-1]) + 1
    idx = np.r_[0, change]
    return idx, dense[idx]


def step_values_at(indices, values, query):
    """Value of the step function (indices, values) at integer positions
    ``query`` (helper for tolerance comparisons of float tracks)."""
    j = np.searchsorted(indices, query, side="right") - 1
    return np.where(j >= 0, np.asarray(values)[np.maximum(j, 0)], 0)
