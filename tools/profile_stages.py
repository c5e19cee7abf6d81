"""Per-kernel time breakdown of one callpeaks pass (CUDA events recorded by the
library around every launch).  python tools/profile_stages.py [backbone_bp] [reads]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from graph_peak_caller_b200 import _lib, pipeline
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads

bp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 51_000_000
nr = int(float(sys.argv[2])) if len(sys.argv) > 2 else 5_000_000
t = time.time()
g = make_graph(bp, 0.02, seed=20261001)
s = make_reads(g, nr, 36, 200, seed=1)
c = make_reads(g, nr, 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
lm = LinearMap.from_graph(g)
print("generated %r in %.1fs" % (g, time.time() - t), flush=True)
dg = DeviceGraph(g, lm.as_arrays())
ds, dc = DeviceReads(s), DeviceReads(c)
torch.cuda.synchronize()
for it in range(7):
    _lib.profile_enable(it == 6)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, qt, q, thr, cleaned, peaks = pipeline.callpeaks(dg, ds, dc, 36, 200, True, None, True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("iter %d: %.2f ms  (%d+%d unique reads, %d events, %d/%d sample/direct runs, "
          "%d control entries, %d p runs, %d distinct p, %d thr runs, %d cleaned, %d peaks)" % (
              it, dt * 1e3, st.n_sample, st.n_control, st.n_events, st.sample[0].numel(),
              st.direct[0].numel(), st.control[0].numel(), st.p_values[0].numel(),
              qt[0].numel(), thr[0].numel(), cleaned[0].numel(), len(peaks)), flush=True)
rep = _lib.profile_report()
tot = sum(v[1] for v in rep.values())
print("kernel time total %.3f ms over %d launches" % (tot, sum(v[0] for v in rep.values())))
for name, (calls, ms, nbytes) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s %4d calls %9.3f ms  %5.1f%%" % (name, calls, ms, 100 * ms / tot))

# GPU idle time between consecutive launches (GPC_PROF_TIMELINE=<file> set before the run)
import os
path = os.environ.get("GPC_PROF_TIMELINE")
if path and os.path.exists(path):
    recs = [l.split() for l in open(path)]
    recs = [(n, float(a), float(d)) for n, a, d in recs]
    gaps = []
    for (n0, a0, d0), (n1, a1, d1) in zip(recs, recs[1:]):
        gaps.append((a1 - (a0 + d0), n0, n1))
    span = recs[-1][1] + recs[-1][2]
    print("timeline: span %.3f ms, kernels %.3f ms, idle %.3f ms" % (
        span, sum(r[2] for r in recs), sum(g[0] for g in gaps)))
    for gap, n0, n1 in sorted(gaps, reverse=True)[:40]:
        print("  idle %7.1f us between %-22s and %s" % (gap * 1e3, n0, n1))
if path and os.path.exists(path):
    print("rs_onesweep launches (ms):", " ".join("%.3f" % d for n_, a_, d in recs if n_ == "rs_onesweep"))
