"""GPU: tracks from arbitrary event lists (gpc_diffs_to_runs / gpc_diffs_merge /
gpc_diffs_combine, csrc/tracks.cu) through the C ABI and through the reference's
container methods (SparseDiffs.get_sparse_values / maximum / apply_binary_func / clip_min /
from_pileup, sparsediffs.py:130-224) against the oracle's numpy restatement, which is pinned
on the live reference.

Tolerance: integer-valued and dyadic diffs -> bit-identical (their partial sums are exact
whatever the order of the additions); arbitrary float diffs -> 1e-9 relative to the largest
partial sum (np.cumsum adds left to right, the scan does not)."""
import numpy as np
import pytest

from oracle import tracks as otracks

pytestmark = pytest.mark.gpu


def _events(n, span, rng, kind="int"):
    idx = rng.integers(0, span, size=n).astype(np.int64)
    if kind == "int":
        diffs = rng.integers(-3, 4, size=n).astype(np.float64)
    elif kind == "dyadic":
        diffs = rng.integers(-40, 41, size=n).astype(np.float64) / 8.0
    else:
        diffs = rng.normal(size=n) * 10.0 ** rng.integers(-3, 4, size=n)
    return idx, diffs


def _host(t):
    a = t.cpu().numpy()
    return a.view(np.uint32).astype(np.int64) if a.dtype == np.int32 else a


@pytest.mark.parametrize("n,span", [(1, 10), (2, 1), (5, 3), (257, 40), (4096, 900), (4097, 10 ** 6),
                                    (100_003, 5_000), (100_003, 2 ** 31), (1_300_000, 400_000),
                                    (3_000_017, 2 ** 32 - 1)])
@pytest.mark.parametrize("kind", ["int", "dyadic"])
def test_diffs_to_runs_exact(n, span, kind):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n + span % 1000)
    idx, diffs = _events(n, span, rng, kind)
    pos, val = kernels.diffs_to_runs(idx, diffs)
    ri, rv = otracks.diffs_to_runs(idx, diffs)
    assert np.array_equal(_host(pos), ri)
    assert np.array_equal(_host(val), rv)


def test_diffs_to_runs_empty_and_errors():
    from graph_peak_caller_b200 import kernels
    pos, val = kernels.diffs_to_runs(np.zeros(0, dtype=np.int64), np.zeros(0))
    assert pos.numel() == 0 and val.numel() == 0
    with pytest.raises(Exception):
        kernels.diffs_to_runs(np.array([3, -1, 5]), np.ones(3))
    with pytest.raises(Exception):
        kernels.diffs_to_runs(np.array([3, 2 ** 32, 5]), np.ones(3))
    # the library is usable after the error
    pos, val = kernels.diffs_to_runs(np.array([4, 2, 4]), np.array([1.0, 2.0, -1.0]))
    assert _host(pos).tolist() == [2] and _host(val).tolist() == [2.0]


@pytest.mark.parametrize("n", [3000, 1_200_000])
def test_float_diffs_within_rounding(n):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n)
    ia, da = _events(n, n // 3, rng, "float")
    ib, db = _events(n // 2, n // 3, rng, "float")
    pos, a, b = kernels.diffs_merge(ia, da, ib, db)
    idx, vals, keep = otracks._merge_two(ia, da, ib, db)
    assert np.array_equal(_host(pos), idx)
    for got, col, d in ((a, vals[0], da), (b, vals[1], db)):
        tol = 1e-9 * np.abs(np.cumsum(np.abs(d))).max()
        assert np.allclose(_host(got), col[keep], rtol=0, atol=tol)


@pytest.mark.parametrize("na,nb,span", [(1, 1, 5), (7, 0, 30), (0, 9, 30), (5000, 3000, 700),
                                        (900_000, 700_000, 2_000_000)])
def test_merge_and_combine_exact(na, nb, span):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(na * 7 + nb)
    ia, da = _events(na, span, rng)
    ib, db = _events(nb, span, rng, "dyadic")
    idx, vals, keep = otracks._merge_two(ia, da, ib, db)
    pos, a, b = kernels.diffs_merge(ia, da, ib, db)
    assert np.array_equal(_host(pos), idx)
    assert np.array_equal(_host(a), vals[0][keep])
    assert np.array_equal(_host(b), vals[1][keep])
    for name, func in (("maximum", np.maximum), ("minimum", np.minimum), ("add", np.add),
                       ("subtract", np.subtract), ("multiply", np.multiply)):
        pos, val = kernels.diffs_combine(ia, da, ib, db, name)
        ri, rv = otracks.binary_func_runs(ia, da, ib, db, func)
        assert np.array_equal(_host(pos), ri), name
        assert np.array_equal(_host(val), rv), name
    # SparseDiffs.maximum: zero steps dropped, the leading one included
    pos, val = kernels.diffs_combine(ia, da, ib, db, "maximum", drop_leading_zero=True)
    mi, md = otracks.maximum_diffs(ia, da, ib, db)
    v = _host(val)
    assert np.array_equal(_host(pos), mi)
    assert np.array_equal(np.ediff1d(v, to_begin=v[:1]) if v.size else v, md)


def test_container_methods_on_hand_made_tracks():
    """The reference's unit tests build SparseDiffs by hand; with a device those methods
    are kernels, and give what the reference's numpy statements give."""
    from graph_peak_caller_b200.sparsediffs import SparseDiffs, SparseValues
    rng = np.random.default_rng(11)
    ia, da = _events(4000, 900, rng)
    ib, db = _events(2500, 900, rng)
    a, b = SparseDiffs(ia, da), SparseDiffs(ib, db)
    sv = a.get_sparse_values()
    assert sv.on_device()
    ri, rv = otracks.diffs_to_runs(ia, da)
    assert np.array_equal(sv.indices, ri) and np.array_equal(sv.values, rv)
    m = a.maximum(b)
    mi, md = otracks.maximum_diffs(ia, da, ib, db)
    assert np.array_equal(m._indices, mi) and np.array_equal(m._diffs, md)
    # known ufunc (kernel) and arbitrary callable (merge + sums on the device)
    for func in (np.minimum, lambda x, y: np.abs(x - 2 * y)):
        got = a.apply_binary_func(func, b, return_values=True)
        ri, rv = otracks.binary_func_runs(ia, da, ib, db, func)
        assert isinstance(got, SparseValues)
        assert np.array_equal(got.indices, ri) and np.array_equal(got.values, rv)
    got = a.apply_binary_func(np.add, b)
    idx, vals, keep = otracks._merge_two(ia, da, ib, db)
    ret = (vals[0] + vals[1])[keep]
    assert np.array_equal(got._indices, idx)
    assert np.array_equal(got._diffs, np.ediff1d(ret, to_begin=ret[0]))
    # clip_min on a sorted track
    si = np.unique(ia)
    sd = rng.integers(-3, 4, size=si.size).astype(np.float64)
    c = SparseDiffs(si.copy(), sd.copy())
    c.clip_min(0.5)
    ci, cd = otracks.clip_min_diffs(si, sd, 0.5)
    assert np.array_equal(c._indices, ci) and np.array_equal(c._diffs, cd)


def test_from_pileup_inputs():
    """SURVEY step 2's recipe: the reference's from_pileup inputs (read starts, ends and the
    per-node carried values, sparsediffs.py:180-190) through the generic kernel."""
    from graph_peak_caller_b200.sparsediffs import SparseDiffs

    class Pileup:
        pass
    rng = np.random.default_rng(3)
    node_indexes = np.r_[0, np.cumsum(rng.integers(1, 40, size=3000))]
    size = int(node_indexes[-1])
    p = Pileup()
    p.starts = rng.integers(0, size, size=20_000)
    p.ends = np.minimum(p.starts + rng.integers(1, 60, size=p.starts.size), size)
    p.node_starts = rng.integers(-2, 3, size=node_indexes.size).astype(np.float64)
    sd = SparseDiffs.from_pileup(p, node_indexes)
    sv = sd.get_sparse_values()
    idx = np.r_[p.starts, p.ends, node_indexes]
    diffs = np.r_[np.ones(p.starts.size), -np.ones(p.ends.size), p.node_starts]
    ri, rv = otracks.diffs_to_runs(idx, diffs)
    assert np.array_equal(sv.indices, ri) and np.array_equal(sv.values, rv)


def test_in_place_scaling_after_the_track_was_looked_at():
    """a kernel-made track whose host copy exists (somebody read ``_indices``)
    must carry ``*=`` / ``/=`` on BOTH forms -- the p-value stage reads the device runs with
    their scale first (reference sparsediffs.py:200-206 scales the one array it has)."""
    from scipy.stats import poisson
    from graph_peak_caller_b200.device import to_device
    from graph_peak_caller_b200.sparsediffs import SparseDiffs
    from graph_peak_caller_b200.sparsepvalues import PValuesFinder
    t = SparseDiffs.from_device_runs(to_device(np.array([0, 5, 9], dtype=np.int32)),
                                     to_device(np.array([2, 6, 0], dtype=np.int32)))
    assert t._indices.tolist() == [0, 5, 9]            # materialises the host copy
    t *= 0.5
    assert t.device_runs()[2] == 0.5 and t._diffs.tolist() == [1.0, 2.0, -3.0]
    t /= 0.25
    assert t.device_runs()[2] == 2.0 and t._diffs.tolist() == [4.0, 8.0, -12.0]
    control = SparseDiffs(np.array([0, 7]), np.array([1.5, 1.0]))      # lambda 1.5, then 2.5
    p = PValuesFinder(t, control).get_p_values_pileup()
    want = [-poisson.logsf(4, 1.5) / np.log(10), -poisson.logsf(12, 1.5) / np.log(10),
            -poisson.logsf(12, 2.5) / np.log(10), 0.0]
    assert p.indices.tolist() == [0, 5, 7, 9]
    assert np.allclose(p.values, want, rtol=1e-9, atol=0)
