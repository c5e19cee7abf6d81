#!/bin/bash
# closing check of round 2: the whole GPU suite and the default bench line on the library as committed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2_tests_close.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_close.log
tail -n 3 gpurun_out/r2_tests_close.log
GPC_BENCH_ALL_KERNELS=1 timeout 400 python bench.py --no-cpu > gpurun_out/r2_bench_close_chr22.json 2> gpurun_out/r2_bench_close_chr22.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_close_chr22.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"])
PY
