"""TEST INFRASTRUCTURE ONLY (build container): golden vectors for fragment-length
estimation from the UNMODIFIED reference (PeakModel, HaploTyper, LinearFilter through
oracle/refshim) -> tests/golden/shift/shift_model.npz.

    python -m oracle.make_golden_shift
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbridge  # noqa: E402


def synthetic_positions(seed, genome=1_500_000, sites=260, chroms=("0", "1")):
    """Binding sites with a plus pile d/2 left and a minus pile d/2 right of the centre,
    plus uniform background; returns (positions dict, d)."""
    rng = np.random.default_rng(seed)
    d = int(rng.integers(80, 280))
    pos = {"+": {}, "-": {}}
    for chrom in chroms:
        centers = rng.choice(genome - 4000, sites, replace=False) + 2000
        plus, minus = [rng.integers(0, genome, 4000)], [rng.integers(0, genome, 4000)]
        for c in centers:
            plus.append((c - d // 2 + rng.normal(0, 12, int(rng.integers(8, 50)))).astype(np.int64))
            minus.append((c + d // 2 + rng.normal(0, 12, int(rng.integers(8, 50)))).astype(np.int64))
        pos["+"][chrom] = np.concatenate(plus)
        pos["-"][chrom] = np.concatenate(minus)
    return pos, d, genome * len(chroms)


def reference_model(pos, genome_size, lmfold=5, umfold=50):
    refbridge._ref()
    from graph_peak_caller.shiftestimation.shiftestimation import Opt, PeakModel, Treatment
    opt = Opt(lmfold, umfold)
    opt.gsize = genome_size
    return PeakModel(opt, Treatment({s: {c: list(v) for c, v in pos[s].items()} for s in pos}))


def reference_linear_filter(graph, reads):
    """(path node ids, {"+": [...], "-": [...]}, path length) from the live HaploTyper +
    LinearFilter."""
    obg, _ = refbridge._ref()
    from graph_peak_caller.haplotyping import HaploTyper
    from graph_peak_caller.linear_filter import LinearFilter
    rg = refbridge.to_ref_graph(graph)
    intervals = refbridge.to_ref_intervals(reads, rg)
    typer = HaploTyper(rg, obg.IntervalCollection(intervals))
    typer.build()
    path = typer.get_maximum_interval_through_graph()
    found = LinearFilter((i.start_position for i in intervals), path).find_start_positions()
    return list(path.region_paths), found, path.length()


def main():
    from graph_peak_caller_b200.synthetic import make_graph, make_reads
    out = {}
    for k, seed in enumerate((3, 4)):
        pos, d, genome = synthetic_positions(seed)
        model = reference_model(pos, genome)
        for s, tag in (("+", "plus"), ("-", "minus")):
            for c in pos[s]:
                out["m%d_%s_%s" % (k, tag, c)] = pos[s][c]
        out["m%d_meta" % k] = np.array([genome, d, model.min_tags, model.max_tags], dtype=np.int64)
        out["m%d_d" % k] = np.array([model.d])
        out["m%d_plus_line" % k] = model.plus_line
        out["m%d_minus_line" % k] = model.minus_line
        out["m%d_ycorr" % k] = model.ycorr
        out["m%d_alternative_d" % k] = np.array(model.alternative_d, dtype=np.int64)
    g = make_graph(30000, 0.05, seed=61, mnp_frac=0.2)
    reads = make_reads(g, 3000, 20, 80, seed=62, peak_frac=0.5, peak_spacing=2500)
    path, found, length = reference_linear_filter(g, reads)
    out["lf_graph_args"] = np.array([30000, 61, 3000, 62], dtype=np.int64)
    out["lf_path"] = np.array(path, dtype=np.int64)
    out["lf_plus"] = np.array(found["+"], dtype=np.int64)
    out["lf_minus"] = np.array(found["-"], dtype=np.int64)
    out["lf_length"] = np.array([length], dtype=np.int64)
    name = os.path.join(ROOT, "tests", "golden", "shift", "shift_model.npz")
    np.savez_compressed(name, **out)
    print("wrote", name, os.path.getsize(name), "bytes")


if __name__ == "__main__":
    main()
