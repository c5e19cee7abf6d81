"""TEST INFRASTRUCTURE -- full-size parity digests of the BASELINE workloads.

    python oracle/make_digests.py WORKLOAD [--scale F] [--procs P] [--out FILE]

Runs the oracle restatement (oracle/pipeline.py, pinned against the live reference by
tests/test_oracle_vs_reference.py) over every chromosome of a named synthetic
workload (graph_peak_caller_b200/synth_core.py CONFIGS = BASELINE.json `configs`) at
its FULL size -- minutes to an hour of CPU, which is why it runs here, once, and not
in the test suite -- and writes SHA-256 digests of everything the path must reproduce
bit for bit (unique reads, direct pileup, fragment pileup, touched nodes, thresholded
and hole-cleaned tracks, peak intervals with their node lists), plus a thinned sample of
the float tracks (control lambda, p, q: every k-th run) for the 1e-6 comparison.  The
whole-genome join is the reference's: ONE p -> q mapping from the p-values of all
chromosomes (sparsepvalues.py:76-93).  tests/test_gpu_digests.py recomputes the same
digests from the CUDA path's output on the GPU box.

Phases (multiprocessing over chromosomes): 1. reads -> p-values, per chromosome, state
pickled to a scratch directory; 2. joined p table -> q table; 3. p-values -> peaks.
"""
import argparse
import hashlib
import importlib.util
import json
import os
import pickle
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

FLOAT_SAMPLES = 512           # runs kept per float track


def synth_core():
    spec = importlib.util.spec_from_file_location(
        "gpc_synth_core", os.path.join(ROOT, "graph_peak_caller_b200", "synth_core.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sha(*arrays):
    """Digest of integer arrays, each taken as little-endian int64."""
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(np.asarray(a).astype("<i8")).tobytes())
    return h.hexdigest()


def peaks_digest(peaks):
    """Order-free digest of peak intervals: sorted by (start, end, nodes)."""
    rows = sorted((int(p[0]), int(p[1]), tuple(int(x) for x in p[2])) for p in peaks)
    h = hashlib.sha256()
    for s, e, nodes in rows:
        h.update(np.array([s, e, len(nodes)] + list(nodes), dtype="<i8").tobytes())
    return h.hexdigest(), len(rows)


def thin(idx, val):
    """Every k-th run of a float track: positions (exact) and values (to 1e-6)."""
    idx, val = np.asarray(idx), np.asarray(val, dtype=np.float64)
    k = max(1, idx.size // FLOAT_SAMPLES)
    return dict(every=int(k), n=int(idx.size), positions=[int(x) for x in idx[::k]],
                values=[float(x) for x in val[::k]])


def _phase1(job):
    cfg, ci, scratch = job
    from oracle import control as ocontrol, pipeline as opipeline
    sc = synth_core()
    t0 = time.time()
    g, s, c = sc.chromosome_inputs(cfg, ci)
    lm = ocontrol.linear_map_from_graph(g)
    su = opipeline.unique_reads(s)
    cu = opipeline.unique_reads(c) if c is not None else su
    st = opipeline.run_to_p_values(g, lm, su, cu, cfg["read_length"], cfg["fragment_length"],
                                   cfg["has_control"], sc.global_min_of(cfg))
    keep = dict(p_values=st["p_values"], touched=st["touched"], direct=st["direct"],
                track_size=st["track_size"])
    with open(os.path.join(scratch, "state_%d.pkl" % ci), "wb") as fh:
        pickle.dump(keep, fh, protocol=4)
    from oracle import tracks
    sample = st["sample"] if "sample" in st else tracks.diffs_to_runs(*st["sample_diffs"])
    rec = dict(name=cfg["chromosomes"][ci][0], graph_bp=int(g.node_off[-1]), nodes=int(g.n_nodes),
               edges=int(g.succ.size), reads=[len(s), len(c) if c is not None else len(s)],
               unique_reads=[len(su.start_node), len(cu.start_node)],
               direct=sha(*st["direct"]), direct_runs=int(len(st["direct"][0])),
               fragment=sha(*sample), fragment_runs=int(len(sample[0])),
               touched=sha(st["touched"]), touched_nodes=int(len(st["touched"])),
               control=thin(*st["control"]), p_values=thin(*st["p_values"]),
               p_runs=int(len(st["p_values"][0])), seconds_phase1=time.time() - t0)
    # the join needs (distinct p, base pairs) only: reduced here, so that the parent never
    # holds the p tracks of 24 chromosomes (integer counts: the reduced tables join exactly)
    from oracle import stats as ostats
    sp, cum = ostats.p_table(st["p_values"][1], ostats.run_lengths(st["p_values"][0], st["track_size"]))
    return ci, rec, (sp, np.ediff1d(cum, to_begin=cum[0]))


def _phase3(job):
    cfg, ci, scratch, table = job
    from oracle import pipeline as opipeline
    sc = synth_core()
    t0 = time.time()
    g, _, _ = sc.chromosome_inputs(dict(cfg, n_sample=100, n_control=100 if cfg["n_control"] else None), ci)
    with open(os.path.join(scratch, "state_%d.pkl" % ci), "rb") as fh:
        st = pickle.load(fh)
    out = opipeline.run_from_p_values(g, st, table, cfg["read_length"], cfg["fragment_length"])
    pd, n = peaks_digest(out["peaks"])
    ad, na = peaks_digest(out["all_max_paths"])
    rec = dict(thresholded=sha(out["thresholded"][0], np.asarray(out["thresholded"][1]).astype(np.int64)),
               thresholded_runs=int(len(out["thresholded"][0])),
               hole_cleaned=sha(out["hole_cleaned"][0], np.asarray(out["hole_cleaned"][1]).astype(np.int64)),
               hole_cleaned_runs=int(len(out["hole_cleaned"][0])),
               q_values=thin(*out["q_values"]), peaks=pd, n_peaks=n, all_max_paths=ad,
               n_all_max_paths=na, seconds_phase3=time.time() - t0)
    return ci, rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--procs", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import multiprocessing as mp
    from oracle import pipeline as opipeline, stats as ostats
    sc = synth_core()
    cfg = sc.CONFIGS[a.workload]
    if a.scale > 1:
        cfg = sc.scaled(cfg, a.scale)
    n = len(cfg["chromosomes"])
    # the graph of phase 3 must be the graph of phase 1: it depends on (bp, seed) only
    out_path = a.out or os.path.join(ROOT, "tests", "golden", "digest_%s%s.json" % (
        a.workload, "" if a.scale <= 1 else "_div%g" % a.scale))
    t0 = time.time()
    with tempfile.TemporaryDirectory(prefix="gpc_digest_") as scratch:
        procs = max(1, min(a.procs, n))
        # largest chromosomes first: the pool finishes sooner
        order = sorted(range(n), key=lambda i: -cfg["chromosomes"][i][1])
        with mp.get_context("fork").Pool(procs) as pool:
            res1 = pool.map(_phase1, [(cfg, i, scratch) for i in order], chunksize=1)
        recs = {ci: rec for ci, rec, _ in res1}
        # PToQValuesMapper.from_files (sparsepvalues.py:76-93) over the per-chromosome tables
        parts = [t for _, _, t in sorted(res1, key=lambda r: r[0])]
        sp, cum = ostats.p_table(np.concatenate([t[0] for t in parts]),
                                 np.concatenate([t[1] for t in parts]))
        table = ostats.q_table(sp, cum)
        with mp.get_context("fork").Pool(procs) as pool:
            res3 = pool.map(_phase3, [(cfg, i, scratch, table) for i in order], chunksize=1)
        for ci, rec in res3:
            recs[ci].update(rec)
    keys = np.array(sorted(table.keys()), dtype=np.float64)
    doc = dict(workload=a.workload, scaled_by=a.scale, made_by="oracle/make_digests.py",
               oracle="oracle/pipeline.py (port, pinned against the live reference)",
               read_length=cfg["read_length"], fragment_length=cfg["fragment_length"],
               has_control=cfg["has_control"], global_min=sc.global_min_of(cfg),
               q_table=dict(entries=int(keys.size), total_bp=int(cum[-1]) if len(cum) else 0,
                            p=[float(x) for x in keys[::max(1, keys.size // FLOAT_SAMPLES)]],
                            q=[float(table[x]) for x in keys[::max(1, keys.size // FLOAT_SAMPLES)]]),
               chromosomes=[recs[i] for i in range(n)], seconds=time.time() - t0)
    with open(out_path, "w") as fh:
        json.dump(doc, fh, separators=(",", ":"))
    print("wrote", out_path, "in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
