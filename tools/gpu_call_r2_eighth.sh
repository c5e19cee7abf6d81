#!/bin/bash
# bulk-copy (1-D TMA) pipelined variant of the difference-array scan against the default;
# walk kernel after the spill-pool split
mkdir -p gpurun_out
GPC_DENSE_SCAN=3 timeout 300 python -m pytest tests/test_gpu_sample.py tests/test_cyclic.py -x -q > gpurun_out/r2_tests7_v3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests7_v3.log
tail -n 3 gpurun_out/r2_tests7_v3.log
timeout 300 python -m pytest tests/test_gpu_sample.py tests/test_gpu_guards.py -x -q > gpurun_out/r2_tests7_v0.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests7_v0.log
tail -n 3 gpurun_out/r2_tests7_v0.log
for v in 0 3; do
  GPC_DENSE_SCAN=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_7_v$v.json 2> gpurun_out/r2_bench_chr22_7_v$v.err
done
GPC_DENSE_SCAN=3 timeout 300 python -m pytest tests/test_gpu_digests.py -x -q -k "chr22 or dense" > gpurun_out/r2_tests7_dig_v3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests7_dig_v3.log
tail -n 3 gpurun_out/r2_tests7_dig_v3.log
