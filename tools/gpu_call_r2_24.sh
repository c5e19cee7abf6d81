#!/bin/bash
# delta-coded wire form: the sample tests (all upload forms), then the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sample.py tests/test_cyclic.py tests/test_gpu_edge_cases.py -q > gpurun_out/r2_tests24.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests24.log
tail -n 12 gpurun_out/r2_tests24.log
export GPC_BENCH_ALL_KERNELS=1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench24_chr22.json 2> gpurun_out/r2_bench24_chr22.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench24_chr22.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"])
PY
