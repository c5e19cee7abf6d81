"""Host-side throughput of the ingest row: vg alignment JSON ->
reads SoA through gpc_parse_vg_alignments_host, next to what the json module
alone needs for the same lines (the floor of any Python-object parser such as
pyvg's).  No GPU involved.

    python tools/ingest_rate.py [n_reads]

bench.py adds the result to its JSON line as "ingest" (rank 0, one GPU)."""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def alignment_lines(n_reads, read_length=36, seed=7):
    """`vg view -aj`-shaped lines (name, sequence, quality, path with 1-4 mappings,
    refpos, score) over made-up node ids; ~1 KB per line like real vg output."""
    rng = np.random.default_rng(seed)
    first = rng.integers(1, 3_000_000, size=n_reads)
    n_map = rng.integers(1, 5, size=n_reads)
    reverse = rng.random(n_reads) < 0.5
    offset = rng.integers(0, 30, size=n_reads)
    seq = "ACGT" * (read_length // 4)
    qual = "I" * 48
    out = []
    for i in range(n_reads):
        k = int(n_map[i])
        step = -1 if reverse[i] else 1
        rev = ', "is_reverse": true' if reverse[i] else ""
        share = read_length // k
        maps = []
        for j in range(k):
            length = share if j < k - 1 else read_length - share * (k - 1)
            off = ', "offset": %d' % offset[i] if j == 0 and offset[i] else ""
            maps.append('{"position": {"node_id": %d%s%s}, "rank": %d, "edit": '
                        '[{"from_length": %d, "to_length": %d}]}'
                        % (first[i] + step * j if first[i] + step * j > 0 else 1, off, rev,
                           j + 1, length, length))
        out.append('{"mapping_quality": 60, "sequence": "%s", "identity": 1.0, "path": '
                   '{"mapping": [%s]}, "name": "SRR931836.%d HWI-ST132:571:C0HAPACXX:4:1101:%d '
                   'length=50", "quality": "%s", "score": 46, "refpos": [{"name": "22", '
                   '"offset": %d, "is_reverse": true}]}'
                   % (seq, ", ".join(maps), i, i, qual, int(first[i]) * 17))
    return out


def _split_time(name, repeats):
    """The chromosome splitter (split_vg_json_reads_into_chromosomes) on the same file,
    three made-up node ranges: one native pass, lines written to four files."""
    from graph_peak_caller_b200 import conversion
    d = tempfile.mkdtemp()
    try:
        for chrom, (lo, hi) in {"1": (1, 1_000_000), "2": (1_000_001, 2_000_000),
                                "3": (2_000_001, 4_000_000)}.items():
            with open(os.path.join(d, "node_range_%s.txt" % chrom), "w") as fh:
                fh.write("%d:%d" % (lo, hi))
        copy = os.path.join(d, "reads.json")
        os.link(name, copy) if os.stat(name).st_dev == os.stat(d).st_dev else shutil.copy(name, copy)
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            conversion.split_vg_json_reads_into_chromosomes("1,2,3", copy, d + "/")
            times.append(time.perf_counter() - t0)
        return min(times)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def measure(n_reads=200_000, repeats=3, json_reads=20_000):
    from graph_peak_caller_b200.reads import ReadBatch
    lines = alignment_lines(n_reads)
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
        fh.write("\n".join(lines) + "\n")
        name = fh.name
    try:
        size = os.path.getsize(name)
        best = {}
        for label, threads in (("one_thread", 1), ("all_threads", 0)):
            times = []
            for _ in range(repeats):
                t0 = time.perf_counter()
                batch = ReadBatch.from_vg_json_file(name, n_threads=threads)
                times.append(time.perf_counter() - t0)
            assert len(batch) == n_reads
            best[label] = min(times)
        best["split"] = _split_time(name, repeats)
    finally:
        os.unlink(name)
    t0 = time.perf_counter()
    for line in lines[:json_reads]:
        json.loads(line)
    t_json = time.perf_counter() - t0
    return dict(
        format="vg alignment JSON (vg view -aj), file -> reads SoA on the host",
        reads=n_reads, file_mb=size / 1e6, host_threads=os.cpu_count(),
        reads_per_sec=n_reads / best["all_threads"],
        mb_per_sec=size / 1e6 / best["all_threads"],
        reads_per_sec_one_thread=n_reads / best["one_thread"],
        json_module_reads_per_sec=min(json_reads, n_reads) / t_json,
        split_lines_per_sec=n_reads / best["split"], split_mb_per_sec=size / 1e6 / best["split"],
        note="native parser: mapped file, one parse, copy out; best of %d; json module: "
             "json.loads per line only (no interval objects), one thread" % repeats)


if __name__ == "__main__":
    print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 200_000)))
