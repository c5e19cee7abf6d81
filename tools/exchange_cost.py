"""Under torchrun: wall time of the p-table exchange inside one pass.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_cost.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from graph_peak_caller_b200 import multigraph, pipeline
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads
g = make_graph(51_000_000, 0.02, seed=20261001 + rank)
s = make_reads(g, 5_000_000, 36, 200, seed=1 + 10 * rank)
c = make_reads(g, 5_000_000, 36, 200, seed=2 + 10 * rank, peak_frac=0.0, dup_frac=0.0)
dg = DeviceGraph(g, LinearMap.from_graph(g).as_arrays())
ds, dc = DeviceReads(s), DeviceReads(c)
acc = {}
def lap(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + t1 - t0; return t1
for it in range(8):
    if it == 3: acc.clear()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    st = pipeline.run_to_p_values(dg, ds, dc, 36, 200, True, None, True); t = lap("to_p_values", t)
    tp, tc = pipeline.local_p_table(st.p_values, st.track_size); t = lap("local_table", t)
    tp, tc, total = multigraph.gather_p_tables(tp, tc, st.track_size); t = lap("exchange", t)
    qp, qq = pipeline.q_table(tp, tc, total, joined=True); t = lap("q_table", t)
    pipeline.run_from_p_values(dg, st.p_values, st.touched, st.direct, qp, qq, 36, 200); t = lap("from_p_values", t)
print("rank", rank, {k: round(1e3 * v / 5, 3) for k, v in acc.items()}, "joined table", tp.numel(), flush=True)
dist.destroy_process_group()
