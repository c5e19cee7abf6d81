"""Where the stream idles: one profiled pass (event pair around every launch), launches in
stream order with the gap after each.  python tools/timeline.py [workload] [scale] [out]"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, ".")
from graph_peak_caller_b200 import _lib, pipeline, synthetic
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads

cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "chr22_scale"]
if len(sys.argv) > 2 and float(sys.argv[2]) > 1:
    cfg = synthetic.scaled(cfg, float(sys.argv[2]))
g, s, c = synthetic.chromosome_inputs(cfg, 0)
dg = DeviceGraph(g, LinearMap.from_graph(g).as_arrays())
ds = DeviceReads(s)
dc = DeviceReads(c) if c is not None else ds
R, F = cfg["read_length"], cfg["fragment_length"]


def step():
    return pipeline.callpeaks(dg, ds, dc, R, F, cfg["has_control"], synthetic.global_min_of(cfg), True)


for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print("plain pass: %.3f ms" % (e0.elapsed_time(e1) / 5))
path = sys.argv[3] if len(sys.argv) > 3 else tempfile.mktemp(suffix=".timeline")
os.environ["GPC_PROF_TIMELINE"] = path
_lib.profile_enable(True)
_lib.profile_report()
step()
_lib.profile_enable(False)
_lib.profile_report()
rows = [l.split() for l in open(path)]
rows = [(n, float(a), float(d)) for n, a, d in rows]
total = rows[-1][1] + rows[-1][2] - rows[0][1]
busy = sum(d for _, _, d in rows)
print("profiled pass: %.3f ms from first launch to last end, %.3f ms inside kernels, %d launches"
      % (total, busy, len(rows)))
print("%-4s %-22s %9s %9s %9s" % ("#", "kernel", "start", "dur", "gap after"))
for i, (n, a, d) in enumerate(rows):
    gap = rows[i + 1][1] - (a + d) if i + 1 < len(rows) else 0.0
    print("%-4d %-22s %9.3f %9.3f %9.3f%s" % (i, n, a, d, gap, "  <--" if gap > 0.02 else ""))
