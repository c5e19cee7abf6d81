"""profiles/traffic.json from an ncu launch list of bench.py:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu
  python tools/make_traffic.py gpurun_out/launches.csv "<that command>"

Per kernel label (the names bench.py's roofline uses): launches, share of the
summed kernel time, DRAM bytes per launch (read + write, averaged over the
launches).  ncu serialises launches and its times are cold-cache, so only the
SHARE is comparable with the live CUDA-event numbers."""
import collections, csv, json, re, sys

ALIAS = {"pv_eval": "pvalues", "dense_runs_block": "dense_runs", "dense_runs_pipelined": "dense_runs"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
        "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}

def label(kernel_name):
    base = re.split(r"[<(]", kernel_name)[0].split("::")[-1].replace("void ", "").strip()
    base = re.sub(r"_kernel$", "", base)
    return ALIAS.get(base, base)

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
head = rows[0]
col = {k: i for i, k in enumerate(head)}
acc = collections.defaultdict(lambda: dict(launches=0, ms=0.0, dram=0.0))
ids = collections.defaultdict(dict)
for r in rows[1:]:
    if len(r) != len(head):
        continue
    k = r[col["Kernel Name"]]
    lab = label(k)
    if "at::" in k or "nccl" in k.lower():     # not this library's
        lab = "(torch) " + lab
    ids[r[col["ID"]]][r[col["Metric Name"]]] = (float(r[col["Metric Value"]].replace(",", "")) *
                                                UNIT.get(r[col["Metric Unit"]], 1.0), lab)
for kid, m in ids.items():
    lab = next(iter(m.values()))[1]
    a = acc[lab]
    a["launches"] += 1
    a["ms"] += m.get("gpu__time_duration.sum", (0.0,))[0]
    a["dram"] += m.get("dram__bytes_read.sum", (0.0,))[0] + m.get("dram__bytes_write.sum", (0.0,))[0]
total_ms = sum(a["ms"] for a in acc.values())
out = dict(source="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                  "--clock-control none of: " + (sys.argv[2] if len(sys.argv) > 2 else "bench.py"),
           total_kernel_ms=total_ms, kernels={})
for lab, a in sorted(acc.items(), key=lambda kv: -kv[1]["ms"]):
    out["kernels"][lab] = dict(launches=a["launches"], ms=round(a["ms"], 4),
                               share_of_kernel_time=round(a["ms"] / total_ms, 4),
                               dram_bytes_per_launch=round(a["dram"] / a["launches"]))
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
for lab, e in list(out["kernels"].items())[:14]:
    print("%-24s %5d launches %8.3f ms %5.1f%%  %10.2f MB DRAM per launch" % (
        lab, e["launches"], e["ms"], 100 * e["share_of_kernel_time"], e["dram_bytes_per_launch"] / 1e6))
