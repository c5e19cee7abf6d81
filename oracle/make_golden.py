"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the unmodified
reference (build container only: needs /root/reference, imported through
oracle/refshim).

    python oracle/make_golden.py

Each file holds the flat inputs of one case (graph CSR, sample/control reads,
configuration) and every intermediate track plus the final peaks produced by
the reference's CallPeaks.run on them.  The cases are (a) hand-built graphs
taken from the shapes of the reference's unit tests (tests/test_whole_callpeaks.py,
tests/test_create_sample_pileup.py: linear, split, diamond, hierarchical
bubbles) with planted fragments, and (b) seeded synthetic variation graphs
from graph_peak_caller_b200.synthetic.
"""
import json
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from graph_peak_caller_b200.graph import Graph  # noqa: E402
from graph_peak_caller_b200.reads import Interval, ReadBatch  # noqa: E402
from graph_peak_caller_b200.synthetic import make_graph, make_reads  # noqa: E402
from oracle import refbridge  # noqa: E402

OUT = os.path.join(_ROOT, "tests", "golden")


def planted_reads(graph, fragments, read_length, fragment_length, copies=10):
    """reference tests/test_whole_callpeaks.py:12-22: per planted fragment,
    `copies` forward reads at its left end and `copies` reversed reads at its
    right end."""
    reads = []
    for frag in fragments:
        frag.graph = graph
        for _ in range(copies):
            reads.append(frag.get_subinterval(0, read_length))
            right = frag.get_subinterval(fragment_length - read_length,
                                         fragment_length)
            reads.append(right.get_reverse())
    return ReadBatch.from_intervals(reads)


def hand_built_cases():
    cases = []
    # split graph, tests/test_whole_callpeaks.py:57-97
    g = Graph({i: 15 for i in range(1, 5)}, {1: [2, 3], 2: [4], 3: [4]})
    for name, frags in [
        ("split_single", [Interval(3, 8, [2])]),
        ("split_crossing", [Interval(13, 3, [1, 2])]),
        ("split_two", [Interval(13, 3, [1, 2]), Interval(13, 3, [2, 4])]),
    ]:
        r = planted_reads(g, frags, 1, 5)
        cases.append((name, g, r, r, dict(read_length=1, fragment_length=5,
                                          has_control=False)))
    # long linear graph with two fragments
    g = Graph({i: 40 for i in range(1, 9)}, {i: [i + 1] for i in range(1, 8)})
    r = planted_reads(g, [Interval(30, 20, [2, 3]), Interval(5, 35, [6])], 6, 30)
    cases.append(("linear_two", g, r, r, dict(read_length=6, fragment_length=30,
                                              has_control=False)))
    # hierarchical bubbles, shape of tests/test_create_control_whole.py:20-36
    sizes = {1: 10, 2: 5, 3: 12, 4: 4, 5: 6, 6: 3, 7: 9, 8: 7, 9: 2, 10: 8, 11: 10}
    edges = {1: [2, 3], 2: [4, 5], 3: [7], 4: [6], 5: [6], 6: [7], 7: [8, 9],
             8: [10], 9: [10], 10: [11]}
    g = Graph(sizes, edges)
    r = planted_reads(g, [Interval(4, 3, [1, 3, 7]), Interval(2, 5, [7, 8, 10])],
                      4, 21, copies=6)
    cases.append(("hier_bubbles", g, r, r, dict(read_length=4, fragment_length=21,
                                                has_control=True)))
    return cases


def synthetic_cases():
    cases = []
    specs = [
        # name, bp, site_rate, mnp, indel_mean, R, F, n_s, n_c, has_control, min_node
        ("syn_snp_small", 4000, 0.03, 0.0, 5, 10, 40, 500, 700, True, 1),
        ("syn_mnp_scaled", 6000, 0.05, 0.3, 8, 12, 50, 900, 1000, True, 1),
        ("syn_input_only", 8000, 0.02, 0.0, 5, 15, 60, 800, 900, False, 1),
        ("syn_dense_var", 5000, 0.12, 0.2, 20, 9, 45, 1000, 1200, True, 1),
        ("syn_more_sample", 5000, 0.03, 0.1, 5, 10, 36, 1200, 500, True, 1),
        ("syn_offset_ids", 3000, 0.03, 0.0, 5, 8, 30, 60, 400, True, 101),
        # more sample than control reads at a size where peaks and holes are plentiful
        # (the reference's float residue at zero coverage)
        ("syn_more_sample_big", 30000, 0.03, 0.1, 5, 12, 60, 6000, 2500, True, 1),
    ]
    for i, (name, bp, rate, mnp, im, R, F, ns, nc, hc, mn) in enumerate(specs):
        g = make_graph(bp, site_rate=rate, seed=7000 + i, min_node=mn,
                       mnp_frac=mnp, indel_mean=im)
        s = make_reads(g, ns, R, F, seed=8000 + i, peak_frac=0.7,
                       peak_spacing=1200, dup_frac=0.03)
        c = make_reads(g, nc, R, F, seed=9000 + i, peak_frac=0.0, dup_frac=0.0)
        cases.append((name, g, s, c, dict(read_length=R, fragment_length=F,
                                          has_control=hc)))
    return cases


def _track(d, key, pair):
    d[key + "_idx"] = np.asarray(pair[0])
    d[key + "_val"] = np.asarray(pair[1])


def main():
    if not refbridge.available():
        raise SystemExit("needs /root/reference")
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])          # python oracle/make_golden.py [case ...]
    for name, g, s, c, cfg in hand_built_cases() + synthetic_cases():
        if only and name not in only:
            continue
        unique = name.startswith("syn_")
        try:
            ref = refbridge.ref_callpeaks(g, s, c, cfg["read_length"],
                                          cfg["fragment_length"],
                                          has_control=cfg["has_control"],
                                          unique=unique)
        except AssertionError as e:
            print("SKIP", name, "reference raised", str(e)[:70])
            continue
        d = dict(node_len=g.node_len, min_node=g.min_node, succ_ptr=g.succ_ptr,
                 succ=g.succ)
        for tag, r in (("s", s), ("c", c)):
            for f in ("start_node", "start_off", "end_node", "end_off",
                      "path_ptr", "path"):
                d["%s_%s" % (tag, f)] = getattr(r, f)
        lm = refbridge.ref_linear_map(g)
        d["lin_starts"], d["lin_ends"] = lm
        for key in ("sample", "direct", "control_diffs", "control", "p_values",
                    "q_values", "hole_cleaned"):
            _track(d, key, ref[key])
        d["touched"] = ref["touched"]
        p2q = sorted(ref["p_to_q"].items(), reverse=True)
        d["p2q_p"] = np.array([float(k) for k, _ in p2q])
        d["p2q_q"] = np.array([float(v) for _, v in p2q])
        meta = dict(cfg, unique=unique, n_sample=int(ref["n_sample"]),
                    n_control=int(ref["n_control"]),
                    track_size=ref["track_size"], peaks=ref["peaks"],
                    all_max_paths=ref["all_max_paths"])
        d["meta"] = np.array(json.dumps(meta))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-18s nodes %5d reads %5d/%5d peaks %3d  %6.1f kB" % (
            name, g.n_nodes, len(s), len(c), len(ref["peaks"]),
            os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
