#!/bin/bash
# multi-block q table: tests that hold the table bit for bit, bench lines, full-size whole-genome digest
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_control_stats.py tests/test_gpu_callpeaks_api.py tests/test_gpu_multigraph.py tests/test_gpu_configs.py tests/test_gpu_digests.py tests/test_gpu_reference_known_answers.py -x -q > gpurun_out/r2_tests21.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests21.log
tail -n 3 gpurun_out/r2_tests21.log
export GPC_BENCH_ALL_KERNELS=1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench21_chr22.json 2> gpurun_out/r2_bench21_chr22.err
GPC_FULL_SIZE=1 timeout 1500 python -m pytest tests/test_gpu_digests.py -x -q -k human > gpurun_out/r2_tests21_human_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests21_human_full.log
tail -n 5 gpurun_out/r2_tests21_human_full.log
