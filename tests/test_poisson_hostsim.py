"""CPU: the exact p-value code the kernels run (csrc/poisson.cuh, compiled for
the host) against scipy.stats.poisson.logsf as the reference uses it,
including the underflow -> 1000 rule."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.stats import clean_p_values

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def poisson_eval(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostsim") / "poisson_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "hostsim", "poisson_host.cpp")])
    lib = ctypes.CDLL(so)

    def ev(k, lam):
        k = np.ascontiguousarray(k, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        out = np.empty_like(k)
        lib.poisson_eval(k.ctypes.data_as(ctypes.c_void_p), lam.ctypes.data_as(ctypes.c_void_p),
                         out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(k.size))
        return out
    return ev


def check(ev, K, L):
    ref = clean_p_values(K.copy(), L.copy())
    got = ev(K, L)
    assert np.array_equal(ref == 1000, got == 1000), "underflow rule differs"
    m = ref != 1000
    assert np.all(np.abs(got[m] - ref[m]) <= 1e-6 * np.abs(ref[m]) + 1e-12)


def test_known_values_of_the_reference_tests(poisson_eval):
    # reference tests/test_pvalues.py:58-66
    got = poisson_eval([2.0, 3.0], [1.0, 1.0])
    assert np.allclose(got, [-np.log10(0.08030), -np.log10(0.018988)], rtol=2e-4)
    assert np.allclose(poisson_eval([1.0], [1.0]), [-np.log10(0.26424)], rtol=2e-5)


def test_grid(poisson_eval):
    ks = np.r_[np.arange(0, 60), np.arange(60, 400, 7), np.arange(400, 5000, 97),
               [0.5, 1.5, 2.999, 17.3]]
    lams = np.r_[0.0, np.logspace(-4, 3.7, 120), [0.02, 0.2, 1.0, 5.0, 50.0]]
    K, L = np.meshgrid(ks, lams)
    check(poisson_eval, K.ravel(), L.ravel())


def test_underflow_boundary(poisson_eval):
    rng = np.random.default_rng(1)
    K = rng.integers(1, 700, size=100000).astype(float)
    L = 10 ** rng.uniform(-3, 1.6, size=K.size)
    check(poisson_eval, K, L)
