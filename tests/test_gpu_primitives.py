"""GPU: the sort / scan building blocks through the C ABI against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a):
    from graph_peak_caller_b200.device import to_device
    return to_device(a)


@pytest.mark.parametrize("n,bits", [(1, 32), (2, 32), (257, 32), (4096, 32), (4097, 32),
                                    (100_003, 32), (1_000_000, 29), (3_000_017, 24),
                                    (50_000, 8), (70_000, 13),
                                    (2_000_003, 26), (300_000, 27), (123_457, 17), (5000, 25)])
def test_sort_u32(n, bits):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** bits, size=n, dtype=np.uint64).astype(np.uint32)
    out = kernels.sort_u32(_dev(keys.view(np.int32)), 0, bits)
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, np.sort(keys))


def test_sort_u32_skewed():
    """Few distinct keys (the shape of pileup events in a peak)."""
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(5)
    keys = (rng.integers(0, 37, size=500_000) * 1001).astype(np.uint32)
    got = kernels.sort_u32(_dev(keys.view(np.int32)), 0, 32).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, np.sort(keys))


@pytest.mark.parametrize("n", [1, 5, 2049, 300_001])
def test_sort_u64_pairs_stable(n):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n)
    # doubles mapped to order-preserving keys, many ties
    x = np.round(rng.random(n) * 1000.0, 1) * rng.choice([1.0, 1e3, 1e-3], size=n)
    u = x.view(np.uint64) | np.uint64(1 << 63)
    vals = np.arange(n, dtype=np.int32)
    k, v = kernels.sort_u64_pairs(_dev(u.view(np.int64)), _dev(vals))
    k = k.cpu().numpy().view(np.uint64)
    v = v.cpu().numpy()
    order = np.argsort(u, kind="stable")
    assert np.array_equal(k, u[order])
    assert np.array_equal(v, vals[order])


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 1_234_567])
def test_exclusive_scan(n):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n)
    x = rng.integers(0, 9, size=n).astype(np.int32)
    out, total = kernels.exclusive_scan_u32(_dev(x))
    ref = np.cumsum(x) - x
    assert np.array_equal(out.cpu().numpy(), ref)
    assert total == int(x.sum())


@pytest.mark.parametrize("n_table,n_runs", [(1, 70_000), (37, 70_000), (35_000, 300_000),
                                            (800_000, 2_000_000), (1_200_000, 100_000),
                                            (500, 3_000)])
def test_qvalues_apply_exact_key_lookup(n_table, n_runs):
    """gpc_qvalues_apply: every run's p is looked up by exact key (sparsepvalues.py:116-125:
    a dict from p to q; p == 0 -> 0).  Tracks of >= 2^16 runs with tables of <= 2^20 entries go
    through the per-call hash table, everything else through the binary search: both routes
    against numpy, on tables of distinct doubles that differ in their last bits."""
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n_table + n_runs)
    base = np.sort(rng.random(n_table) * 50.0 + 1e-9)[::-1]
    table_p = np.unique(np.r_[base, np.nextafter(base[: n_table // 3], np.inf)])[::-1].copy()
    table_q = np.minimum.accumulate(rng.random(table_p.size) * 10.0)[::-1][::-1].copy()
    pick = rng.integers(0, table_p.size, size=n_runs)
    p = table_p[pick]
    zero = rng.random(n_runs) < 0.1
    p[zero] = 0.0
    got = kernels.qvalues_apply(_dev(p), _dev(table_p), _dev(table_q)).cpu().numpy()
    want = np.where(zero, 0.0, table_q[pick])
    assert np.array_equal(got, want)
    # a p that is not in the table is an error, not a silent zero
    p[0] = 123.456
    with pytest.raises(Exception):
        kernels.qvalues_apply(_dev(p), _dev(table_p), _dev(table_q))
