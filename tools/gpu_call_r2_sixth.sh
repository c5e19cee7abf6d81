#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests5.log
tail -n 6 gpurun_out/r2_tests5.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_5.json 2> gpurun_out/r2_bench_chr22_5.err
python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_dense_5.json 2> gpurun_out/r2_bench_dense_5.err
python bench.py --workload human_wg --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_human_n1.json 2> gpurun_out/r2_bench_human_n1.err
