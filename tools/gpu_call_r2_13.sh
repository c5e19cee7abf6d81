#!/bin/bash
# look-back over 128 predecessors per round trip (all chained scans), p-table capacity
# remembered between calls: whole suite + bench lines with the full kernel list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests13.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests13.log
tail -n 3 gpurun_out/r2_tests13.log
export GPC_BENCH_ALL_KERNELS=1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench13_chr22.json 2> gpurun_out/r2_bench13_chr22.err
timeout 600 python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench13_dense.json 2> gpurun_out/r2_bench13_dense.err
timeout 600 python bench.py --workload human_wg --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench13_human_wg.json 2> gpurun_out/r2_bench13_human_wg.err
