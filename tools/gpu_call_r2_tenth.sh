#!/bin/bash
# staged-output difference-array scan (default) against the first version (GPC_DENSE_SCAN=1),
# 16 items per thread in the exclusive scans, density switch of the linear control track,
# component-size statistics of the dense-variation workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests10.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests10.log
tail -n 5 gpurun_out/r2_tests10.log
for v in 0 1; do
  GPC_DENSE_SCAN=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench10_chr22_v$v.json 2> gpurun_out/r2_bench10_chr22_v$v.err
done
GPC_MP_STATS=1 timeout 600 python bench.py --workload dense_variation --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_bench10_dense.json 2> gpurun_out/r2_bench10_dense.err
timeout 600 python bench.py --workload human_wg --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench10_human_wg.json 2> gpurun_out/r2_bench10_human_wg.err
grep -m1 -A8 "max_paths:" gpurun_out/r2_bench10_dense.err
