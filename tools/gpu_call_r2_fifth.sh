#!/bin/bash
# round 2, fifth GPU call: the three difference-array scan variants side by side
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sample.py -x -q > gpurun_out/r2_tests4a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests4a.log
for v in 0 1 2; do
  GPC_DENSE_SCAN=$v python -m pytest tests/test_gpu_sample.py -x -q -k "golden or random" > gpurun_out/r2_tests4_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests4_v$v.log
  GPC_DENSE_SCAN=$v python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_4_v$v.json 2> gpurun_out/r2_bench_chr22_4_v$v.err
done
tail -2 gpurun_out/r2_tests4*.log
