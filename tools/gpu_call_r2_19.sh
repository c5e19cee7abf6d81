#!/bin/bash
# two-level tile totals, prefix looked up by every tile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_sample.py tests/test_gpu_control_stats.py tests/test_gpu_postprocess.py tests/test_gpu_linearmap.py -x -q > gpurun_out/r2_tests19.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests18.log
tail -n 3 gpurun_out/r2_tests19.log
export GPC_BENCH_ALL_KERNELS=1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench19_chr22.json 2> gpurun_out/r2_bench19_chr22.err
