"""Worker of tests/test_gpu_multirank.py -- launched under torch.distributed.run, one
process per GPU (NCCL).  The 24 chromosome graphs of the human_wg workload (scaled down)
are packed onto the ranks by MultipleGraphsCallpeaks (longest processing time first);
every rank checks that

  * all ranks hold a bit-identical joined (p, q) table,
  * that table equals the single-process join of ALL chromosomes (rank 0 and rank 1
    each recompute every chromosome's local p table themselves), bit for bit,
  * the peaks of its own chromosomes equal those of a single-process run with the
    single-process table.

Reference: PToQValuesMapper.from_files / get_p_to_q_values (sparsepvalues.py:76-103),
MultipleGraphsCallpeaks.run (multiplegraphscallpeaks.py:106-168).  Prints
"MULTIRANK OK ..." on rank 0; any difference raises.
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local_rank = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from graph_peak_caller_b200 import Configuration, pipeline, synthetic
    from graph_peak_caller_b200.control.linearmap import LinearMap
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
    from graph_peak_caller_b200.reporter import Reporter

    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 64.0
    cfg = synthetic.scaled(synthetic.CONFIGS["human_wg"], scale)
    R, F = cfg["read_length"], cfg["fragment_length"]
    gmin = synthetic.global_min_of(cfg)
    n = len(cfg["chromosomes"])
    inputs = [synthetic.chromosome_inputs(cfg, i) for i in range(n)]      # small: every rank makes all
    names = [c[0] for c in cfg["chromosomes"]]
    config = Configuration()
    config.read_length, config.fragment_length = R, F
    config.has_control, config.global_min = True, gmin
    lms = [LinearMap.from_graph(g) for g, _, _ in inputs]
    caller = MultipleGraphsCallpeaks(names, [g for g, _, _ in inputs], [s for _, s, _ in inputs],
                                     [c for _, _, c in inputs], lms, config, Reporter(None))
    mine = caller.my_chromosomes()
    caller.run()                                   # p-values, ONE all-gather, peaks
    mapping = caller._q_value_mapping
    tp, tq = (t.cpu().numpy() for t in mapping.device_table)
    digest = hashlib.sha256(tp.tobytes() + tq.tobytes()).digest()
    # 1. identical on every rank
    all_digests = [torch.zeros(32, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(all_digests, torch.tensor(list(digest), dtype=torch.uint8, device="cuda"))
    assert all(bytes(d.cpu().tolist()) == digest for d in all_digests), "ranks hold different q tables"
    # 2. equal to the single-process join (this process alone, all chromosomes)
    tabs, states, total_bp = [], {}, 0
    for i, (g, s, c) in enumerate(inputs):
        dg = DeviceGraph(g, lms[i].as_arrays())
        st = pipeline.run_to_p_values(dg, DeviceReads(s), DeviceReads(c), R, F, True, gmin, True)
        tabs.append(pipeline.local_p_table(st.p_values, st.track_size))
        total_bp += st.track_size
        if i in mine:
            states[i] = (dg, st)
    qp, qq = pipeline.q_table(torch.cat([t[0] for t in tabs]), torch.cat([t[1] for t in tabs]),
                              total_bp)
    assert np.array_equal(qp.cpu().numpy(), tp) and np.array_equal(qq.cpu().numpy(), tq), \
        "distributed join differs from the single-process join"
    # 3. peaks of this rank's chromosomes
    n_peaks = 0
    for i in mine:
        dg, st = states[i]
        _, _, _, raw = pipeline.run_from_p_values(dg, st.p_values, st.touched, st.direct, qp, qq, R, F)
        want = sorted((p[0], p[1], tuple(p[2])) for p in pipeline.PeakSet(raw, F).final_peaks())
        got = sorted((p.start_position.offset, p.end_position.offset, tuple(p.region_paths))
                     for p in caller.peaks()[names[i]])
        assert got == want, "peaks of chromosome %s differ" % names[i]
        n_peaks += len(got)
    counts = torch.tensor([n_peaks, len(mine)], dtype=torch.int64, device="cuda")
    dist.all_reduce(counts)
    assert int(counts[1]) == n, "every chromosome must have exactly one owner"
    if rank == 0:
        print("MULTIRANK OK world=%d chromosomes=%d peaks=%d table=%d entries sha256=%s"
              % (world, n, int(counts[0]), tp.size, digest.hex()[:16]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
