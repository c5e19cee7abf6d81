"""Stage-isolated timing.
  python tools/bench_stage.py dump            # run the pipeline once, save stage inputs to /tmp
  python tools/bench_stage.py pvalues|linear  # time one stage on the saved inputs (GPC_B200_LIB picks the library)
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from graph_peak_caller_b200 import _lib, kernels, pipeline
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads

what = sys.argv[1]
PATH = "/tmp/stage_inputs.pt"
R, F = 36, 200


def timed(fn, n=6):
    out = None
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


if what == "dump":
    g = make_graph(51_000_000, 0.02, seed=20261001)
    s = make_reads(g, 5_000_000, 36, 200, seed=1)
    c = make_reads(g, 5_000_000, 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
    lm = LinearMap.from_graph(g)
    dg = DeviceGraph(g, lm.as_arrays())
    st = pipeline.run_to_p_values(dg, DeviceReads(s), DeviceReads(c), R, F, True, None, True)
    torch.save(dict(spos=st.sample[0].cpu(), sval=st.sample[1].cpu(), cpos=st.control[0].cpu(),
                    cval=st.control[1].cpu(), ppos=st.p_values[0].cpu(), p=st.p_values[1].cpu()),
               PATH)
    print("dumped", st.sample[0].numel(), st.control[0].numel(), st.p_values[0].numel())
elif what == "pvalues":
    d = torch.load(PATH)
    dev = torch.device("cuda")
    spos, sval, cpos, cval = (d[k].to(dev) for k in ("spos", "sval", "cpos", "cval"))
    ms, (ppos, p) = timed(lambda: kernels.pvalues(spos, sval, 1.0, cpos, cval))
    same = ppos.numel() == d["ppos"].numel() and bool((p.cpu() == d["p"]).all())
    _lib.profile_enable(True)
    kernels.pvalues(spos, sval, 1.0, cpos, cval)
    rep = _lib.profile_report()
    print("pvalues stage %.3f ms, kernel %.3f ms, %d runs, identical to the dump: %s" % (
        ms, rep["pvalues"][1], ppos.numel(), same))
    print({k: round(v[1], 3) for k, v in rep.items()})
    if len(sys.argv) > 2 and sys.argv[2] == "stats":
        u = torch.unique(torch.cat([spos, cpos]))
        k = sval[(torch.searchsorted(spos, u, right=True) - 1).clamp(min=0)]
        # last control entry of a position wins
        lam = cval[(torch.searchsorted(cpos, u, right=True) - 1).clamp(min=0)]
        nz = k != 0
        pairs = torch.unique(torch.stack([k[nz].to(torch.float64), lam[nz]]), dim=1)
        print("positions %d, with count != 0: %d, distinct (count, lambda) pairs: %d, distinct lambda %d, max count %d"
              % (u.numel(), int(nz.sum()), pairs.shape[1], torch.unique(lam).numel(), int(k.max())))
elif what == "sort32":
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 16_000_000
    bits = int(sys.argv[3]) if len(sys.argv) > 3 else 26
    rng = np.random.default_rng(0)
    keys = torch.from_numpy(rng.integers(0, 2 ** bits, size=n, dtype=np.int64).astype(np.uint32)
                            .view(np.int32)).cuda()
    _lib.profile_enable(True)
    best = 1e9
    for _ in range(6):
        k = keys.clone()
        torch.cuda.synchronize()
        _lib.profile_report()
        out = kernels.sort_u32(k, 0, bits)
        rep = _lib.profile_report()
        ms = rep["rs_onesweep"][1] / rep["rs_onesweep"][0]
        best = min(best, ms)
    ok = bool((out[1:].view(torch.int32).to(torch.int64) & 0xffffffff >= out[:-1].to(torch.int64) & 0xffffffff).all())
    print("sort32 n=%d bits=%d: %d passes, %.1f us per pass, %.0f GB/s per pass (keys read+written), sorted=%s" % (
        n, bits, rep["rs_onesweep"][0], best * 1e3, 8.0 * n / (best / 1e3) / 1e9, ok))
