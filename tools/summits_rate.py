"""GPU: time of the summit kernel on the peaks and q-values of one
chr22-scale pass, next to what the reference's way costs on the host: a dense float64
array of the whole graph (8 B per base pair) plus a Python loop per peak (oracle port,
bounded sample of the peaks).  NOT YET RUN on a B200 (written after round 1's GPU budget
was spent).

    python tools/summits_rate.py [graph_bp] [reads]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from graph_peak_caller_b200 import CallPeaks, Configuration, _lib
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.intervals import UniqueIntervals
from graph_peak_caller_b200.peakcollection import PeakCollection
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.synthetic import make_graph, make_reads

bp = int(sys.argv[1]) if len(sys.argv) > 1 else 51_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
R, F, WINDOW = 36, 200, 60

g = make_graph(bp, 0.02, seed=20261001)
s = make_reads(g, n_reads, R, F, seed=1)
c = make_reads(g, n_reads, R, F, seed=2, peak_frac=0.0, dup_frac=0.0)
config = Configuration()
config.read_length, config.fragment_length, config.has_control = R, F, True
config.linear_map_name = LinearMap.from_graph(g)
caller = CallPeaks(g, config, Reporter(None))
caller.run(UniqueIntervals(s), UniqueIntervals(c))
peaks, q = caller.max_path_peaks, caller.q_values_pileup
collection = PeakCollection(list(peaks))
collection.summits(q, WINDOW)                                   # warm-up (uploads, allocator)
torch.cuda.synchronize()
t0 = time.perf_counter()
rows = collection.summits(q, WINDOW)
torch.cuda.synchronize()
call_ms = (time.perf_counter() - t0) * 1e3
_lib.profile_enable(True)
collection.summits(q, WINDOW)
kernel_ms = _lib.profile_report()["peak_summits"][1]

from oracle import summits as osummits                          # the checker, timed as the baseline
t0 = time.perf_counter()
dense = osummits.dense_track(q.indices, q.values, g.size_bp)
dense_s = time.perf_counter() - t0
sample = list(range(0, len(peaks), max(1, len(peaks) // 200)))
t0 = time.perf_counter()
for i in sample:
    p = peaks[i]
    summit, length, _ = osummits.cut_around_summit(
        g.node_off, g.node_len, g.min_node, dense, p.start_position.offset, p.end_position.offset,
        list(p.region_paths), WINDOW)
    assert (int(rows[i][0]), int(rows[i][1])) == (summit, length)
loop_s = (time.perf_counter() - t0) * len(peaks) / max(1, len(sample))
print({"peaks": len(peaks), "q_runs": len(q), "kernel_ms": round(kernel_ms, 4),
       "call_ms_with_uploads_and_readback": round(call_ms, 3),
       "host_dense_array_s": round(dense_s, 3), "host_loop_s_extrapolated": round(loop_s, 3),
       "checked_against_oracle": len(sample)})
