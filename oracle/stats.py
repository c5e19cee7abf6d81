"""ORACLE (test infrastructure, CPU) -- Poisson p-values and q-values.

Restates graph_peak_caller/sparsepvalues.py of the reference.  The special
function is the very one the reference calls, scipy.stats.poisson.logsf
(third-party, unpinned in setup.py; scipy 1.18.1 in this image evaluates it
as log(pdtrc(floor(k), mu)) -- scipy/stats/_discrete_distns.py), so the oracle
inherits the reference's behaviour at the underflow boundary by construction.
"""
import numpy as np
from scipy.stats import poisson

from . import tracks

LN10 = np.log(10)


def clean_p_values(counts, lambdas):
    """PValuesFinder.get_p_values_pileup.clean_p_values (sparsepvalues.py:19-26)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p = poisson.logsf(counts, lambdas)
        p /= -LN10
        p[counts == 0] = 0
        p[np.isinf(p)] = 1000
    return p


def p_value_runs(sample_idx, sample_diffs, control_idx, control_diffs):
    """PValuesFinder.get_p_values_pileup (sparsepvalues.py:16-31)."""
    return tracks.binary_func_runs(sample_idx, sample_diffs,
                                   control_idx, control_diffs, clean_p_values)


def run_lengths(indices, track_size):
    """PToQValuesMapper.__get_sub_counts (sparsepvalues.py:70-74)."""
    return np.ediff1d(indices, to_end=track_size - indices[-1])


def p_table(p_values, lengths):
    """PToQValuesMapper._from_subcounts (sparsepvalues.py:50-59): distinct p
    in descending order with cumulative base-pair counts."""
    p_values = np.asarray(p_values).ravel()
    lengths = np.asarray(lengths).ravel()
    order = np.argsort(p_values)[::-1]
    sp = p_values[order]
    cum = np.cumsum(lengths[order])
    last = np.ediff1d(sp, to_end=1) != 0
    return sp[last], cum[last]


def q_table(sorted_p, cum_counts):
    """PToQValuesMapper.get_p_to_q_values (sparsepvalues.py:95-103).  Returns
    the dict p -> q."""
    log_n = np.log10(cum_counts[-1])
    q = sorted_p + np.log10(1 + np.r_[0, cum_counts[:-1]]) - log_n
    q[0] = max(0, q[0])
    q = np.minimum.accumulate(q)
    table = dict(zip(sorted_p, q))
    table[0] = 0
    return table


def q_value_runs(p_idx, p_val, table):
    """QValuesFinder.get_q_values (sparsepvalues.py:116-125): exact-key lookup,
    result not re-canonicalised."""
    return p_idx, np.array([table[p] for p in p_val.tolist()], dtype=np.float64)
