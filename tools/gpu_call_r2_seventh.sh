#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests6.log
tail -n 4 gpurun_out/r2_tests6.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_6.json 2> gpurun_out/r2_bench_chr22_6.err
python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_dense_6.json 2> gpurun_out/r2_bench_dense_6.err
python bench.py --workload human_wg --scale 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_human4_6.json 2> gpurun_out/r2_bench_human4_6.err
