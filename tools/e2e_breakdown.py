"""Where the end-to-end API path spends its wall time."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import cProfile, pstats
from graph_peak_caller_b200 import CallPeaks, Configuration
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.intervals import UniqueIntervals
from graph_peak_caller_b200.reads import ReadBatch
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200 import synthetic
# python tools/e2e_breakdown.py [workload] [scale] [chromosome]
cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "chr22_scale"]
if len(sys.argv) > 2 and float(sys.argv[2]) > 1:
    cfg = synthetic.scaled(cfg, float(sys.argv[2]))
g, s, c = synthetic.chromosome_inputs(cfg, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
if c is None:
    c = s
lm = LinearMap.from_graph(g)
config = Configuration(); config.read_length, config.fragment_length, config.has_control = cfg["read_length"], cfg["fragment_length"], cfg["has_control"]
config.global_min = synthetic.global_min_of(cfg)
config.linear_map_name = lm
pinned = {"s": s.pin_memory(), "c": c.pin_memory()}
print("upload bytes per step", pinned["s"].upload_bytes() + pinned["c"].upload_bytes())
def fresh(tag):
    return pinned[tag].fresh_view()
def step():
    caller = CallPeaks(g, config, Reporter(None))
    caller.run(UniqueIntervals(fresh("s")), UniqueIntervals(fresh("c")))
    return caller.max_path_peaks
for _ in range(3): step()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): p = step()
torch.cuda.synchronize(); print("e2e ms/step", 1e3*(time.perf_counter()-t)/5, len(p))
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

# allocator behaviour: does a step still reach cudaMalloc in steady state?
import graph_peak_caller_b200.device as _dev
_orig_empty = torch.empty
slow = []
def _timed_empty(*a, **k):
    t0 = time.perf_counter(); r = _orig_empty(*a, **k); dt = time.perf_counter() - t0
    if dt > 3e-4: slow.append((round(dt * 1e3, 2), r.numel() * r.element_size()))
    return r
torch.empty = _timed_empty
for i in range(4):
    ms0 = torch.cuda.memory_stats()
    step()
    ms1 = torch.cuda.memory_stats()
    print("step", i, "cudaMalloc calls", ms1["num_device_alloc"] - ms0["num_device_alloc"],
          "cudaFree", ms1["num_device_free"] - ms0["num_device_free"],
          "reserved MB", ms1["reserved_bytes.all.current"] >> 20,
          "allocated MB", ms1["allocated_bytes.all.current"] >> 20, "slow empties", slow)
    slow.clear()
