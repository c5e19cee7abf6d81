"""Lists the reference cycles one CallPeaks.run leaves behind (they delay the
release of device tensors until the cyclic collector runs)."""
import gc, sys, collections
import numpy as np, torch
sys.path.insert(0, ".")
from graph_peak_caller_b200 import CallPeaks, Configuration
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.intervals import UniqueIntervals
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.synthetic import make_graph, make_reads
g = make_graph(1_000_000, 0.02, seed=3)
s = make_reads(g, 100_000, 36, 200, seed=1)
c = make_reads(g, 100_000, 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
config = Configuration(); config.read_length, config.fragment_length, config.has_control = 36, 200, True
config.linear_map_name = LinearMap.from_graph(g)
def step():
    caller = CallPeaks(g, config, Reporter(None))
    caller.run(UniqueIntervals(s), UniqueIntervals(c))
    return len(caller.max_path_peaks)
step(); gc.collect(); gc.disable()
a0 = torch.cuda.memory_allocated()
step()
a1 = torch.cuda.memory_allocated()
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
print("allocated before/after step (no gc):", a0, a1, "unreachable objects:", n)
cnt = collections.Counter(type(o).__name__ for o in gc.garbage)
print(cnt.most_common(25))
for o in gc.garbage:
    if type(o).__module__.startswith("graph_peak_caller_b200"):
        refs = [type(r).__name__ for r in gc.get_referrers(o) if r is not gc.garbage][:6]
        print(type(o).__module__, type(o).__name__, "<-", refs)
