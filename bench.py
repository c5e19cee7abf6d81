"""Benchmark of the callpeaks hot path (BASELINE.json: callpeaks reads/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload NAME] [--scale F]

One step = one full callpeaks pass (duplicate filter -> sample pileup ->
control track -> p-values -> joined q-values -> threshold -> hole cleaning -> max
paths -> scored peaks) over the chromosome graphs of a named synthetic workload
(graph_peak_caller_b200/synth_core.py: CONFIGS = BASELINE.json `configs`):

  chr22_scale (default, configs[1])  one chromosome; N > 1: one copy per rank (weak scaling)
  mhc_shaped (configs[0]), dense_variation (configs[4])   one chromosome, likewise
  arabidopsis_shaped (configs[2]), human_wg (configs[3])  5 / 24 chromosomes packed onto
      the N ranks by longest-processing-time (strong scaling), joined by the single
      all-gather of the p-value tables

Prints ONE JSON line on rank 0.  Regions, in order (GPU arm):
  1. W warm-up steps;
  2. the timed region: K steps bracketed by barrier + synchronize, CUDA events on
     the launching stream, max over ranks -> `value`, `ms_per_step`;
  3. the same K steps again with an event pair around every kernel launch (the
     library's profiler) -> per-kernel times for `roofline`; kept out of (2)
     because the event records cost host time;
  4. W warm-up + K timed end-to-end steps through CallPeaks (the reference's API) with
     page-locked HOST arrays (H2D of both read sets and D2H of the peaks inside the
     wall-clock region) -> `e2e`;
  5. rank 0, N == 1: `bench.py --impl reference --steps 1` in a child process ->
     `cpu_baseline`; and the host-side ingest rate -> `ingest`.
`--impl reference` times the UNMODIFIED reference (oracle/_ref through oracle/refshim;
the oracle port only if that cannot be imported) on a bounded slice of the same
workload, one process per chromosome as the reference's README prescribes, inputs
generated and converted outside the timed region, without importing the product.
"""
import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "chr22_scale"
# reads (sample + control, before the duplicate filter) per timed step of the CPU arm:
# ~7 s of the live reference (7-10 k reads/s), ~3 s of the port, per chromosome process
CPU_READS_PER_PROCESS = 60_000


def load_synth_core():
    """The generator, loaded by path: the reference arm must not import the product."""
    spec = importlib.util.spec_from_file_location(
        "gpc_synth_core", os.path.join(ROOT, "graph_peak_caller_b200", "synth_core.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def control_extensions(cfg):
    F = cfg["fragment_length"]
    return [F, 1000, 10000] if cfg["has_control"] else [10000]


def pileup_bp(cfg, n_sample, n_control):
    """nominal deposited coverage."""
    return n_sample * cfg["fragment_length"] + sum(n_control * 2 * (e // 2)
                                                   for e in control_extensions(cfg))


def lpt(costs, n_bins):
    """Longest-processing-time packing (same rule as multigraph.assign_chromosomes)."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(n_bins)
    bins = [[] for _ in range(n_bins)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += costs[i]
    return [sorted(b) for b in bins]


def units_of(cfg, rank, world):
    """[(chromosome index, seed index)] this rank works on, and the scaling kind."""
    n = len(cfg["chromosomes"])
    if n == 1:
        return [(0, rank)], "weak"
    return [(i, i) for i in lpt([b for _, b in cfg["chromosomes"]], world)[rank]], "strong"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        rows = [r.split(", ") for r in open(self.file.name).read().splitlines() if r.strip()]
        os.unlink(self.file.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum,
    averaged over its launches in one pass) from the committed ncu capture of this
    same command; None when the capture has no entry for it."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            table = json.load(fh)
        e = table["kernels"][kernel]
        return e["dram_bytes_per_launch"], table.get("source")
    except (OSError, KeyError, ValueError):
        return None, None


# ---------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref + oracle/refshim), one process per
# chromosome as its README prescribes; inputs are made once, before any timing
# ---------------------------------------------------------------------------

_PREPARED = []          # filled before the pool forks: the children inherit it


def _cpu_run(i):
    from oracle import refrun
    return refrun.run(_PREPARED[i])


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import refrun
    sc = load_synth_core()
    full = sc.CONFIGS[args.workload]
    kind = refrun.reference_kind()
    # the chromosomes the GPU arm works on at this N (all ranks together), scaled down
    units = [u for r in range(args.gpus) for u in units_of(full, r, args.gpus)[0]]
    cores = min(len(units), os.cpu_count() or 1)
    reads_full = full["n_sample"] + (full["n_control"] or full["n_sample"])
    share = sum(full["chromosomes"][i][1] for i, _ in units) / sum(
        b for _, b in full["chromosomes"]) if len(full["chromosomes"]) > 1 else len(units)
    target = args.cpu_reads * (1 if kind == "reference" else 4) * cores
    factor = max(1.0, reads_full * share / target)
    cfg = sc.scaled(full, factor)
    gmin = sc.global_min_of(cfg)
    t_prep = time.perf_counter()
    for ci, si in units:
        g, s, c = sc.chromosome_inputs(cfg, ci, seed_index=si)
        _PREPARED.append(refrun.Prepared(kind, g, s, c, cfg["read_length"], cfg["fragment_length"],
                                         cfg["has_control"], gmin, name="c%d" % len(_PREPARED)))
    t_prep = time.perf_counter() - t_prep
    total_reads, total_time, split, peaks = 0, 0.0, {}, 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if len(_PREPARED) == 1:
            res = [_cpu_run(0)]
        else:
            with mp.get_context("fork").Pool(cores) as pool:
                res = pool.map(_cpu_run, range(len(_PREPARED)), chunksize=1)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            total_reads += sum(r[0] for r in res)
            total_time += dt
            split = {k: sum(r[2].get(k, 0.0) for r in res) for k in res[0][2]}
            peaks = sum(r[3] for r in res)
    value = total_reads / total_time
    ms = 1e3 * total_time / max(args.steps, 1)
    sample = ("%s on %d chromosome(s) of %s scaled 1/%.1f (same generator, same read density): "
              "%d bp of graph, %d nodes, %d reads before the duplicate filter, whole callpeaks pass "
              "(CallPeaks.run split at its method boundaries), inputs prepared outside the timed region"
              % ("unmodified reference (oracle/_ref through oracle/refshim)" if kind == "reference"
                 else "oracle port of the reference algorithm", len(_PREPARED), args.workload, factor,
                 sum(p.graph_bp for p in _PREPARED), sum(p.nodes for p in _PREPARED),
                 sum(p.n_reads_in for p in _PREPARED)))
    line = dict(metric="callpeaks_reads_per_sec", value=value, unit="reads/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling=units_of(full, 0, args.gpus)[1], vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=args.workload, chromosomes=len(_PREPARED),
                            scaled_by=factor, graph_bp=sum(p.graph_bp for p in _PREPARED),
                            nodes=sum(p.nodes for p in _PREPARED),
                            reads=sum(p.n_reads_in for p in _PREPARED),
                            unique_reads_per_step=total_reads // max(args.steps, 1),
                            read_length=cfg["read_length"], fragment_length=cfg["fragment_length"],
                            with_control=cfg["has_control"], global_min=gmin, peaks=peaks),
                cpu_baseline=dict(value=value, unit="reads/s", cores=cores, kind=kind,
                                  sample=sample, stage_seconds=split,
                                  stage_seconds_sum=sum(split.values()),
                                  prepare_seconds_outside_timing=t_prep),
                e2e=dict(value=value, unit="reads/s", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

class Unit:
    """One chromosome of this rank: host inputs and their HBM-resident copies."""


def main_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from graph_peak_caller_b200 import CallPeaks, Configuration, _lib, multigraph, pipeline
    from graph_peak_caller_b200 import synthetic
    from graph_peak_caller_b200.control.linearmap import LinearMap
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    from graph_peak_caller_b200.intervals import UniqueIntervals
    from graph_peak_caller_b200.reporter import Reporter
    from graph_peak_caller_b200.sparsepvalues import PToQValuesMapper

    cfg = synthetic.CONFIGS[args.workload]
    if args.scale > 1:
        cfg = synthetic.scaled(cfg, args.scale)
    R, F, has_control = cfg["read_length"], cfg["fragment_length"], cfg["has_control"]
    gmin = synthetic.global_min_of(cfg)
    mine, scaling = units_of(cfg, rank, world)
    units = []
    for ci, si in mine:
        u = Unit()
        u.name = cfg["chromosomes"][ci][0]
        u.g, u.s, u.c = synthetic.chromosome_inputs(cfg, ci, seed_index=si)
        u.lm = LinearMap.from_graph(u.g)
        u.dg = DeviceGraph(u.g, u.lm.as_arrays())
        u.ds = DeviceReads(u.s)
        u.dc = DeviceReads(u.c) if u.c is not None else u.ds     # no control: the sample is its own
        units.append(u)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    wait_events = []        # (before, after) the exchange: what a rank spends waiting for the others

    def step(collect_wait=False):
        states = [pipeline.run_to_p_values(u.dg, u.ds, u.dc, R, F, has_control, gmin, True)
                  for u in units]
        tabs = [pipeline.local_p_table(st.p_values, st.track_size) for st in states]
        tp = torch.cat([t[0] for t in tabs]) if len(tabs) != 1 else tabs[0][0]
        tc = torch.cat([t[1] for t in tabs]) if len(tabs) != 1 else tabs[0][1]
        total_bp = sum(st.track_size for st in states)
        if collect_wait and world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        tp, tc, total_bp = multigraph.gather_p_tables(tp, tc, total_bp)
        if collect_wait and world > 1:
            b.record()
            wait_events.append((a, b))
        qp, qq = pipeline.q_table(tp, tc, total_bp)
        peaks = []
        for u, st in zip(units, states):
            _, _, _, raw = pipeline.run_from_p_values(u.dg, st.p_values, st.touched, st.direct,
                                                      qp, qq, R, F)
            peaks.append(pipeline.PeakSet(raw, F))
        return states, peaks, (qp, qq)

    # nvidia-smi needs ~0.1 s before its first sample, longer than a short timed
    # region: it is started before the warm-up and runs through the timed and the
    # instrumented steps (the same load throughout)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        states, peaks, qtab = step()
    n_sample = sum(st.n_sample for st in states)
    n_control = sum(st.n_control for st in states)
    n_reads = n_sample + n_control
    n_events = sum(st.n_events for st in states)
    n_peaks = sum(len(p) for p in peaks)
    # ---- timed region: K steps, device time, max over ranks --------------------
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        states, peaks, qtab = step(collect_wait=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    wait_ms = sum(a.elapsed_time(b) for a, b in wait_events)
    launches = (_lib.launch_count() - launches0) // max(args.steps, 1)
    # every rank must hold the same joined q table (sparsepvalues.py:76-103)
    qhash = hashlib.sha256(qtab[0].cpu().numpy().tobytes() + qtab[1].cpu().numpy().tobytes()).digest()
    # ---- the same K steps again with a CUDA-event pair around every launch (on the
    # launching stream): per-kernel durations for the roofline.  Kept out of the
    # region above because 2 x ~120 event records per step cost ~0.5 ms of host time.
    _lib.profile_enable(True)
    _lib.profile_report()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step()
    p1.record()
    barrier()
    instrumented_ms = p0.elapsed_time(p1) / args.steps
    _lib.profile_enable(False)
    prof = _lib.profile_report()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed + instrumented steps"
    busy_ms = (ms - wait_ms) / args.steps
    t = torch.tensor([ms, float(n_reads), float(n_sample), float(n_control), float(n_peaks),
                      float(n_events), busy_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        total_reads, n_sample_all, n_control_all, n_peaks_all, n_events_all = (
            float(tsum[1]), float(tsum[2]), float(tsum[3]), int(tsum[4]), int(tsum[5]))
        busy = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(busy, t[6:7].clone())
        busy = [float(b[0]) for b in busy]
        hashes = [torch.zeros(32, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(hashes, torch.tensor(list(qhash), dtype=torch.uint8, device="cuda"))
        if any(bytes(h.cpu().tolist()) != qhash for h in hashes):
            raise RuntimeError("the ranks hold different joined q tables")
    else:
        total_reads, n_sample_all, n_control_all = float(n_reads), float(n_sample), float(n_control)
        n_peaks_all, n_events_all, busy = n_peaks, n_events, [busy_ms]
    ms_per_step = ms / args.steps
    value = total_reads / (ms_per_step / 1e3)

    # ---- end to end through the public API, host buffers ----------------------
    # page-locked host copies of the reads in the form the API uploads (ReadBatch.pin_memory:
    # the compact wire form when the reads allow it), prepared once like any ingest step
    for u in units:
        u.pinned_s = u.s.pin_memory()
        u.pinned_c = u.c.pin_memory() if u.c is not None else None
        u.config = Configuration()
        u.config.read_length, u.config.fragment_length = R, F
        u.config.has_control, u.config.global_min = has_control, gmin
        u.config.linear_map_name = u.lm
    h2d = sum(v.upload_bytes() for u in units for v in (u.pinned_s, u.pinned_c) if v is not None)

    def fresh(pinned):
        return pinned.fresh_view()

    def e2e_reads(u):
        s = fresh(u.pinned_s)
        c = fresh(u.pinned_c) if u.pinned_c is not None else s    # one upload serves both
        return UniqueIntervals(s), UniqueIntervals(c)

    def e2e_step():
        callers = []
        ahead = e2e_reads(units[0])
        for k, u in enumerate(units):
            caller = CallPeaks(u.g, u.config, Reporter(None))
            s, c = ahead
            if world == 1 and len(units) == 1:
                caller.run(s, c)
                return [caller]
            if k + 1 < len(units):
                # as MultipleGraphsCallpeaks.run_to_p_values does: the next chromosome's
                # reads are uploaded (copy stream) while this one is computed
                ahead = e2e_reads(units[k + 1])
                for reads in set(ahead):
                    reads.prefetch()
            caller.run_to_p_values(s, c)
            callers.append(caller)
        tabs = []
        for u, caller in zip(units, callers):
            pos, p = caller.p_values_pileup.device_arrays()
            tabs.append(pipeline.local_p_table((pos, p), u.g.size_bp))
        tp, tc = torch.cat([t_[0] for t_ in tabs]), torch.cat([t_[1] for t_ in tabs])
        tp, tc, total_bp = multigraph.gather_p_tables(tp, tc, sum(u.g.size_bp for u in units))
        mapping = PToQValuesMapper._from_device_table(tp, tc, total_bp).get_p_to_q_values()
        for caller in callers:
            caller.p_to_q_values_mapping = mapping
            caller.get_q_values()
            caller.call_peaks_from_q_values()
        return callers

    # W untimed steps here too: the API path allocates other sizes than the loop above (fresh
    # uploads every step, buffers shared with the copy stream), and the caching allocator needs
    # a few steps to stop calling cudaMalloc -- with one warm-up step a ~20 ms allocation step
    # fell into the timed region now and then (7.0 vs 11 ms per step at K = 5)
    for _ in range(max(1, args.warmup)):
        callers = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        callers = e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    d2h = int(sum(p.n_all * (4 * 4 + 8 + 1) + p.nodes.size * 4 + 64 * 8 for p in peaks))
    if sum(len(c.max_path_peaks) for c in callers) != n_peaks:
        raise RuntimeError("API path and pipeline path disagree on the peaks")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (live CUDA-event timings) ------------
    peak_gbs, peak_kind = measured_hbm_peak()
    kt = sorted(prof.items(), key=lambda kv: -kv[1][1])
    top, (calls, tms, nbytes) = kt[0]
    achieved = nbytes / (tms / 1e3) / 1e9 if tms > 0 else 0.0
    kernel_ms = sum(v[1] for v in prof.values()) / args.steps
    traffic, traffic_src = ncu_traffic(top)
    roofline = dict(bound="hbm", kernel=top, achieved=achieved, peak=peak_gbs, unit="GB/s",
                    frac=achieved / peak_gbs, traffic=traffic, traffic_source=traffic_src,
                    algorithmic_bytes_per_launch=nbytes / max(calls, 1),
                    peak_source=peak_kind,
                    timing="CUDA-event pair around every launch, second pass over the same "
                           "%d steps (%.3f ms/step with the events)" % (args.steps, instrumented_ms),
                    launches_per_step=calls // args.steps,
                    avg_launch_ms=tms / max(calls, 1),
                    share_of_kernel_time=tms / max(sum(v[1] for v in prof.values()), 1e-9),
                    kernel_ms_per_step=kernel_ms,
                    top8={k: dict(ms_per_step=v[1] / args.steps, launches=v[0] // args.steps,
                                  GBps=(v[2] / (v[1] / 1e3) / 1e9 if v[1] > 0 and v[2] > 0 else None),
                                  frac=(v[2] / (v[1] / 1e3) / 1e9 / peak_gbs
                                        if v[1] > 0 and v[2] > 0 else None))
                          for k, v in kt[:8]})
    if os.environ.get("GPC_BENCH_ALL_KERNELS"):       # every kernel label: ms per step, launches per step
        roofline["all_kernels"] = {k: [round(v[1] / args.steps, 4), v[0] // args.steps] for k, v in kt}
    cpu = None
    ingest = None
    if world == 1 and not args.no_cpu:
        # the reference arm in a process of its own (it must not share this one's imports)
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                                  "--workload", args.workload, "--steps", "2", "--warmup", "0"],
                                 capture_output=True, text=True, timeout=900)
            cpu = json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:   # noqa: BLE001
            cpu = {"error": "%s: %s" % (type(e).__name__, e)}
        # the row before the hot path, host only: vg alignment JSON -> reads SoA
        try:                     # a side figure: it must never cost the line above it
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ingest_rate
            ingest = ingest_rate.measure()
        except Exception as e:   # noqa: BLE001
            ingest = {"error": "%s: %s" % (type(e).__name__, e)}
    graph_bp = sum(u.g.size_bp for u in units)
    line = dict(metric="callpeaks_reads_per_sec", value=value, unit="reads/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="int32+f64",
                data="synthetic",
                config=dict(workload=args.workload, scaled_by=args.scale,
                            chromosomes=len(cfg["chromosomes"]) if scaling == "strong" else world,
                            chromosomes_rank0=[u.name for u in units],
                            graph_bp_rank0=graph_bp, nodes_rank0=sum(u.g.n_nodes for u in units),
                            edges_rank0=sum(u.g.n_edges for u in units),
                            reads_rank0=[sum(len(u.s) for u in units),
                                         sum(len(u.c) if u.c is not None else len(u.s) for u in units)],
                            unique_reads=int(total_reads), read_length=R, fragment_length=F,
                            with_control=has_control, global_min=gmin,
                            peaks=n_peaks_all, events=n_events_all,
                            l2="inputs (%.0f MB reads + %.0f MB graph on rank 0) exceed the 126 MB L2"
                               % (sum(u.ds.host_bytes + (u.dc.host_bytes if u.dc is not u.ds else 0)
                                      for u in units) / 1e6,
                                  sum(u.dg.nbytes() for u in units) / 1e6)),
                pileup_bp_per_sec=pileup_bp(cfg, n_sample_all, n_control_all) / (ms_per_step / 1e3),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=total_reads / (e2e_ms / 1e3), unit="reads/s",
                         ms_per_step=e2e_ms, h2d_bytes_per_step=int(h2d),
                         d2h_bytes_per_step=d2h,
                         wire_form=units[0].pinned_s.wire_form(),
                         note="CallPeaks.run(UniqueIntervals(page-locked host reads), ...) incl. H2D "
                              "of the read sets and D2H of the peaks; the graph is resident"),
                roofline=roofline, cpu_baseline=cpu)
    if world > 1 or scaling == "strong":
        line["sharding"] = dict(
            rule="one chromosome per rank" if scaling == "weak" else
                 "longest-processing-time packing of %d chromosomes by graph size" % len(cfg["chromosomes"]),
            busy_ms_per_rank=busy, imbalance_max_over_mean=max(busy) / (sum(busy) / len(busy)),
            exchange_wait_ms_rank0=wait_ms / args.steps,
            join_check="identical joined q table on all %d ranks (sha256 %s)"
                       % (world, qhash.hex()[:16]))
    if ingest is not None:
        line["ingest"] = ingest
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD,
                    choices=["mhc_shaped", "chr22_scale", "arabidopsis_shaped", "human_wg",
                             "dense_variation"])
    ap.add_argument("--scale", type=float, default=1.0,
                    help="GPU arm: divide the workload by this factor (debugging / small boxes)")
    ap.add_argument("--cpu-reads", type=int, default=CPU_READS_PER_PROCESS,
                    help="CPU arm: reads per chromosome process and step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
