"""Summarises an .ncu-rep (raw page) or a saved raw-page CSV: per kernel duration, DRAM bytes, occupancy,
issue utilisation and the top warp-stall reasons.  python tools/ncu_summary.py rep [out.csv]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
if rep.endswith(".csv"):                     # a saved `--page raw --csv` export
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
col = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[col[k]].replace(",", ""))
    except Exception: return float("nan")
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
def mb(r, k):                                # ncu scales every column on its own
    return f(r, k) * TO_MB.get(rows[1][col[k]], float("nan"))
out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    stalls = sorted(((f(r, k), k.split("issue_stalled_")[1].split("_per_")[0]) for k in h
                     if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")
                     and f(r, k) == f(r, k)), reverse=True)[:3]
    rec = dict(kernel=name, ms=f(r, "gpu__time_duration.sum") / (1e6 if h and rows[1][col["gpu__time_duration.sum"]] == "ns" else 1),
               unit=rows[1][col["gpu__time_duration.sum"]],
               dram_MB=mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum"),
               dram_unit="Mbyte",
               regs=f(r, "launch__registers_per_thread"), grid=f(r, "launch__grid_size"),
               warps_active_pct=f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
               issue_pct=f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               l2_hit=f(r, "lts__t_sector_hit_rate.pct"),
               dram_pct=f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               stalls=" ".join("%s=%.1f" % (n, v) for v, n in stalls))
    out.append(rec)
    print(rec)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as fh:
        w = csv.DictWriter(fh, fieldnames=list(out[0]))
        w.writeheader(); w.writerows(out)
