"""Benchmark of the callpeaks hot path (BASELINE.json: callpeaks reads/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one full callpeaks pass (duplicate filter -> sample pileup ->
control track -> p-values -> q-values -> threshold -> hole cleaning -> max
paths -> scored peaks) over one chromosome graph per GPU: the synthetic
chr22-scale workload of BASELINE.json configs[1] (51 Mbp backbone, 2 % variant
sites, 5 M sample + 5 M control reads of 36 bp, fragment length 200, with
control).  N > 1: one such chromosome per rank (weak scaling), joined by the
single all-gather of the p-value tables.  Prints ONE JSON line on rank 0.

Regions, in order (GPU arm):
  1. W warm-up steps;
  2. the timed region: K steps bracketed by barrier + synchronize, CUDA events on
     the launching stream, max over ranks -> `value`, `ms_per_step`;
  3. the same K steps again with an event pair around every kernel launch (the
     library's profiler) -> per-kernel times for `roofline`; kept out of (2)
     because the event records cost host time;
  4. K end-to-end steps through CallPeaks.run with page-locked HOST arrays
     (H2D of both read sets and D2H of the peaks inside the wall-clock region)
     -> `e2e`;
  5. rank 0, N == 1: the CPU port of the reference algorithm (oracle/) on a
     bounded slice -> `cpu_baseline`; and the host-side ingest rate (vg alignment
     JSON -> reads SoA, tools/ingest_rate.py) -> `ingest`.
`--impl reference` runs only the CPU port (one process per chromosome, as the
reference's README prescribes) and prints the same line with impl=reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="chr22_scale", backbone_bp=51_000_000, site_rate=0.02,
                n_sample=5_000_000, n_control=5_000_000, read_length=36,
                fragment_length=200, has_control=True)
CONFIG_INDEX = 2                      # position in BASELINE.json configs (1-based)
# CPU arm: a bounded slice of the same workload (same generator, same read
# density: 5 M reads / 51 Mbp), sized for a few seconds per step
CPU_SAMPLE = dict(backbone_bp=1_275_000, n_sample=125_000, n_control=125_000)


def make_inputs(spec, chrom):
    from graph_peak_caller_b200.control.linearmap import LinearMap
    from graph_peak_caller_b200.synthetic import config_seed, make_graph, make_reads
    seed = config_seed(CONFIG_INDEX, chrom)
    g = make_graph(spec["backbone_bp"], WORKLOAD["site_rate"], seed=seed)
    s = make_reads(g, spec["n_sample"], WORKLOAD["read_length"], WORKLOAD["fragment_length"],
                   seed=seed + 100)
    c = make_reads(g, spec["n_control"], WORKLOAD["read_length"], WORKLOAD["fragment_length"],
                   seed=seed + 200, peak_frac=0.0, dup_frac=0.0)
    return g, s, c, LinearMap.from_graph(g)


def pileup_bp(n_sample, n_control):
    """nominal deposited coverage."""
    F = WORKLOAD["fragment_length"]
    ext = [F, 1000, 10000] if WORKLOAD["has_control"] else [10000]
    return n_sample * F + sum(n_control * 2 * (e // 2) for e in ext)


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


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (Python / numpy, one
# core per chromosome process as the reference's README prescribes)
# ---------------------------------------------------------------------------

def cpu_pass(chrom):
    from oracle import control as ocontrol, pipeline as opipeline
    g, s, c, _ = make_inputs(CPU_SAMPLE, chrom)
    lm = ocontrol.linear_map_from_graph(g)
    t0 = time.perf_counter()
    su, cu = opipeline.unique_reads(s), opipeline.unique_reads(c)
    timings = {}
    out = opipeline.callpeaks(g, lm, su, cu, WORKLOAD["read_length"],
                              WORKLOAD["fragment_length"], WORKLOAD["has_control"],
                              timings=timings)
    dt = time.perf_counter() - t0
    return len(su.start_node) + len(cu.start_node), dt, timings, len(out["peaks"])


def run_cpu(n_chrom, steps, warmup):
    """Returns (reads/s aggregate, ms per step, cores, stage split of the last step)."""
    import multiprocessing as mp
    cores = min(n_chrom, os.cpu_count() or 1)
    total_reads, total_time, split = 0, 0.0, {}
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if n_chrom == 1:
            res = [cpu_pass(0)]
        else:
            with mp.get_context("fork").Pool(cores) as pool:
                res = pool.map(cpu_pass, range(n_chrom))
        dt = time.perf_counter() - t0
        if it >= warmup:
            total_reads += sum(r[0] for r in res)
            total_time += dt
            split = res[0][2]
    return total_reads / total_time, 1e3 * total_time / max(steps, 1), cores, split


def cpu_sample_text(n_chrom):
    return ("oracle port of the reference algorithm on %d x (%.1f Mbp backbone, %d+%d reads): "
            "1/%d of the per-GPU workload at the same read density, whole callpeaks pass" % (
                n_chrom, CPU_SAMPLE["backbone_bp"] / 1e6, CPU_SAMPLE["n_sample"],
                CPU_SAMPLE["n_control"],
                round(WORKLOAD["backbone_bp"] / CPU_SAMPLE["backbone_bp"])))


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum,
    averaged over its launches in one pass) from the committed ncu capture of this
    same command; None when the capture has no entry for it."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    try:
        with open(path) as fh:
            table = json.load(fh)
        e = table["kernels"][kernel]
        return e["dram_bytes_per_launch"], table.get("source")
    except (OSError, KeyError, ValueError):
        return None, None


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, split = run_cpu(args.gpus, args.steps, args.warmup)
    line = dict(metric="callpeaks_reads_per_sec", value=value, unit="reads/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD["name"], chromosomes=args.gpus,
                            read_length=WORKLOAD["read_length"],
                            fragment_length=WORKLOAD["fragment_length"], with_control=True),
                cpu_baseline=dict(value=value, unit="reads/s", cores=cores, kind="port",
                                  sample=cpu_sample_text(args.gpus), stage_seconds=split),
                e2e=dict(value=value, unit="reads/s", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

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
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    from graph_peak_caller_b200.intervals import UniqueIntervals
    from graph_peak_caller_b200.reads import ReadBatch
    from graph_peak_caller_b200.reporter import Reporter

    R, F = WORKLOAD["read_length"], WORKLOAD["fragment_length"]
    g, s, c, lm = make_inputs(WORKLOAD, rank)
    dg = DeviceGraph(g, lm.as_arrays())
    ds, dc = DeviceReads(s), DeviceReads(c)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        st = pipeline.run_to_p_values(dg, ds, dc, R, F, WORKLOAD["has_control"], None, True)
        tp, tc = pipeline.local_p_table(st.p_values, st.track_size)
        tp, tc, total_bp = multigraph.gather_p_tables(tp, tc, st.track_size)
        qp, qq = pipeline.q_table(tp, tc, total_bp)
        _, _, _, raw = pipeline.run_from_p_values(dg, st.p_values, st.touched, st.direct, qp, qq,
                                                  R, F)
        return st, pipeline.PeakSet(raw, F)

    # nvidia-smi needs ~0.1 s before its first sample, longer than a short timed
    # region: it is started before the warm-up and runs through the timed and the
    # instrumented steps (the same load throughout)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        st, peaks = step()
    n_reads = st.n_sample + st.n_control
    # ---- timed region: K steps, device time, max over ranks --------------------
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st, peaks = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (_lib.launch_count() - launches0) // max(args.steps, 1)
    # ---- the same K steps again with a CUDA-event pair around every launch (on the
    # launching stream): per-kernel durations for the roofline.  Kept out of the
    # region above because 2 x 140 event records per step cost ~0.5 ms of host time.
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
    t = torch.tensor([ms, float(n_reads)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, total_reads = float(tmax[0]), float(t[1])
    else:
        total_reads = float(n_reads)
    ms_per_step = ms / args.steps
    value = total_reads / (ms_per_step / 1e3)

    # ---- end to end through the public API, host buffers ----------------------
    config = Configuration()
    config.read_length, config.fragment_length = R, F
    config.has_control = WORKLOAD["has_control"]
    config.linear_map_name = lm
    pinned = {}
    for tag, rb in (("s", s), ("c", c)):
        pinned[tag] = [torch.from_numpy(a).pin_memory() for a in (
            rb.start_node, rb.start_off, rb.end_node, rb.end_off, rb.path_ptr.astype(np.int32), rb.path)]
    h2d = sum(t_.numel() * t_.element_size() for v in pinned.values() for t_ in v)

    def fresh(tag):
        rb = ReadBatch.__new__(ReadBatch)
        (rb.start_node, rb.start_off, rb.end_node, rb.end_off, rb.path_ptr, rb.path) = \
            [t_.numpy() for t_ in pinned[tag]]
        rb._device = None
        rb._pinned = pinned[tag]
        return rb

    def e2e_step():
        caller = CallPeaks(g, config, Reporter(None))
        if world > 1:
            caller.run_to_p_values(UniqueIntervals(fresh("s")), UniqueIntervals(fresh("c")))
            from graph_peak_caller_b200.sparsepvalues import PToQValuesMapper
            pos, p = caller.p_values_pileup.device_arrays()
            tp, tc = pipeline.local_p_table((pos, p), g.size_bp)
            tp, tc, total_bp = multigraph.gather_p_tables(tp, tc, g.size_bp)
            caller.p_to_q_values_mapping = PToQValuesMapper._from_device_table(
                tp, tc, total_bp).get_p_to_q_values()
            caller.get_q_values()
            caller.call_peaks_from_q_values()
        else:
            caller.run(UniqueIntervals(fresh("s")), UniqueIntervals(fresh("c")))
        return caller.max_path_peaks

    e2e_peaks = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_peaks = e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    d2h = int(peaks.n_all * (4 * 4 + 8 + 1) + peaks.nodes.size * 4 + 64 * 8)
    if len(e2e_peaks) != len(peaks):
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
                    top5={k: dict(ms_per_step=v[1] / args.steps,
                                  GBps=(v[2] / (v[1] / 1e3) / 1e9 if v[1] > 0 and v[2] > 0 else None))
                          for k, v in kt[:5]})
    cpu = None
    ingest = None
    if world == 1 and not args.no_cpu:
        v, cms, cores, split = run_cpu(1, 1, 0)
        cpu = dict(value=v, unit="reads/s", cores=cores, kind="port",
                   sample=cpu_sample_text(1), stage_seconds=split)
        # the row before the hot path, host only: vg alignment JSON -> reads SoA
        try:                     # a side figure: it must never cost the line above it
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ingest_rate
            ingest = ingest_rate.measure()
        except Exception as e:   # noqa: BLE001
            ingest = {"error": "%s: %s" % (type(e).__name__, e)}
    line = dict(metric="callpeaks_reads_per_sec", value=value, unit="reads/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32+f64",
                data="synthetic",
                config=dict(workload=WORKLOAD["name"], chromosomes=world,
                            graph_bp=g.size_bp, nodes=g.n_nodes, edges=g.n_edges,
                            reads_per_gpu=[len(s), len(c)], unique_reads_per_gpu=int(n_reads),
                            read_length=R, fragment_length=F, with_control=True,
                            peaks=len(peaks), events=st.n_events,
                            l2="inputs (%.0f MB reads + %.0f MB graph per GPU) exceed the 126 MB L2"
                               % ((ds.host_bytes) / 1e6 * 2, dg.nbytes() / 1e6)),
                pileup_bp_per_sec=pileup_bp(st.n_sample, st.n_control) * world / (ms_per_step / 1e3),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=total_reads / (e2e_ms / 1e3), unit="reads/s",
                         ms_per_step=e2e_ms, h2d_bytes_per_step=int(h2d),
                         d2h_bytes_per_step=d2h,
                         note="CallPeaks.run(UniqueIntervals(host SoA), ...) incl. pinned H2D of "
                              "both read sets and D2H of the peaks; the graph is resident"),
                roofline=roofline, cpu_baseline=cpu)
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
