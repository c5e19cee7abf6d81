"""CPU, build container only: the oracle against the live, unmodified
reference (imported through oracle/refshim) on randomised graphs.  Skipped
where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest

from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import pipeline, refbridge, control as ocontrol, tracks
from oracle import postprocess as opost

pytestmark = pytest.mark.skipif(not refbridge.available(),
                                reason="/root/reference not present")


def _case(trial):
    rng = np.random.default_rng(1000 + trial)
    bp = int(rng.integers(2000, 12000))
    g = make_graph(bp, site_rate=float(rng.choice([0.01, 0.03, 0.1])), seed=trial,
                   mnp_frac=float(rng.choice([0, 0.2])),
                   indel_mean=float(rng.choice([2, 5, 20])))
    R = int(rng.integers(6, 20))
    F = int(rng.integers(R + 5, 80))
    ns = int(rng.integers(200, 900))
    nc = int(rng.integers(100, 900))
    s = make_reads(g, ns, R, F, seed=trial * 2 + 1, peak_frac=0.7,
                   peak_spacing=int(rng.choice([800, 2500])), dup_frac=0.02)
    c = make_reads(g, nc, R, F, seed=trial * 2 + 2, peak_frac=0.0, dup_frac=0.0)
    return g, s, c, R, F, bool(rng.integers(0, 2))


@pytest.mark.parametrize("trial", range(8))
def test_whole_path_bit_exact(trial):
    g, s, c, R, F, has_control = _case(trial)
    ref = refbridge.ref_callpeaks(g, s, c, R, F, has_control=has_control,
                                  unique=True)
    su, cu = pipeline.unique_reads(s), pipeline.unique_reads(c)
    assert (ref["n_sample"], ref["n_control"]) == (len(su.start_node),
                                                   len(cu.start_node))
    lm = ocontrol.linear_map_from_graph(g)
    rl = refbridge.ref_linear_map(g)
    assert np.array_equal(lm[0], rl[0]) and np.array_equal(lm[1], rl[1])
    out = pipeline.callpeaks(g, lm, su, cu, R, F, has_control=has_control)
    out["sample"] = tracks.diffs_to_runs(*out["sample_scaled_diffs"])
    for key in ("sample", "direct", "control_diffs", "p_values", "q_values",
                "hole_cleaned"):
        assert np.array_equal(np.asarray(out[key][0]), np.asarray(ref[key][0])), key
        assert np.array_equal(np.asarray(out[key][1]), np.asarray(ref[key][1])), key
    assert np.array_equal(out["touched"], ref["touched"])
    assert [(p[0], p[1], p[2]) for p in out["all_max_paths"]] == ref["all_max_paths"]
    assert [tuple(p) for p in out["peaks"]] == [tuple(p) for p in ref["peaks"]]


@pytest.mark.parametrize("trial", range(40))
def test_hole_cleaner_and_max_paths_random_tracks(trial):
    """Random alternating boolean tracks straight into HolesCleaner /
    SparseMaxPaths (much denser in holes and peaks than real q-value tracks)."""
    rng = np.random.default_rng(500 + trial)
    g = make_graph(int(rng.integers(300, 1500)),
                   site_rate=float(rng.choice([0.03, 0.1, 0.2])), seed=trial,
                   mnp_frac=0.2, indel_mean=float(rng.choice([2, 6])))
    G = g.size_bp
    k = int(rng.integers(2, max(3, G // 8)))
    idx = np.unique(np.r_[0, rng.integers(1, G, size=k)])
    first = int(rng.integers(0, 2))
    val = ((np.arange(idx.size) + first) % 2).astype(bool)
    max_size = int(rng.integers(1, 24))
    touched = None
    if rng.random() < 0.7:
        touched = set((np.flatnonzero(rng.random(g.n_nodes) < 0.8)
                       + g.min_node).tolist())
    if not touched:
        touched = None
    try:
        ri, rv = refbridge.ref_clean_holes(g, idx, val, max_size, touched)
    except Exception as e:   # the reference itself fails on some inputs
        with pytest.raises(type(e)):
            opost.clean_holes(g, idx, val, max_size, touched)
        return
    oi, ov = opost.clean_holes(g, idx, val, max_size, touched)
    assert np.array_equal(oi, ri) and np.array_equal(ov, rv)
    # score track: random integer-valued pileup
    sk = int(rng.integers(1, max(2, G // 5)))
    sidx = np.unique(np.r_[0, rng.integers(1, G, size=sk)])
    sval = rng.integers(0, 9, size=sidx.size).astype(np.float64)
    rp, rs = refbridge.ref_max_paths(g, oi, ov, sidx, sval, G, with_subgraphs=True)
    op, osub = opost.max_paths(g, oi, ov, sidx, sval, G, with_subgraphs=True)
    assert [(p[0], p[1], p[2], tuple(p[3])) for p in op] == \
        [(p[0], p[1], p[2], tuple(p[3])) for p in rp]
    # the sub-graph of every peak (what the reference's Reporter.sub_graphs saves)
    assert len(osub) == len(rs) == len(op)
    for (on, om), (rn, rm) in zip(osub, rs):
        assert np.array_equal(np.asarray(on), np.asarray(rn))
        assert om.shape == rm.shape and np.array_equal(om.toarray(), rm.toarray())
