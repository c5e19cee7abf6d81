#!/bin/bash
# A/B on one box: tools/ab.sh OUT_PREFIX "name1:ENV=val ENV2=val" "name2:" ...
# runs bench.py (default workload, no CPU leg) once per variant and prints one line each.
prefix=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
    name=${spec%%:*}; envs=${spec#*:}
    env $envs GPC_BENCH_ALL_KERNELS=1 timeout 400 python bench.py --steps ${AB_STEPS:-10} --warmup 3 --no-cpu \
        > gpurun_out/${prefix}_${name}.json 2> gpurun_out/${prefix}_${name}.err
    python - "$name" gpurun_out/${prefix}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]
    top = sorted(r["all_kernels"].items(), key=lambda kv: -kv[1][0])[:9]
    print("%-10s dev %.3f ms  e2e %.3f ms  kernels %.3f ms  h2d %.1f MB | %s" % (
        sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], r["kernel_ms_per_step"],
        d["e2e"]["h2d_bytes_per_step"] / 1e6, " ".join("%s=%.3f" % (k, v[0]) for k, v in top)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
