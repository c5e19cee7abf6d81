#!/bin/bash
# reduce-then-scan in select_if / exclusive_scan above 2^19 items: whole suite + bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests17.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests16.log
tail -n 3 gpurun_out/r2_tests17.log
export GPC_BENCH_ALL_KERNELS=1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench17_chr22.json 2> gpurun_out/r2_bench17_chr22.err
timeout 600 python bench.py --workload human_wg --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench17_human_wg.json 2> gpurun_out/r2_bench17_human_wg.err
timeout 600 python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench17_dense.json 2> gpurun_out/r2_bench17_dense.err
