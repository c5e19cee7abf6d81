#!/bin/bash
# round 2, first GPU call: the whole GPU suite, then the bench line of every single-GPU workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests0.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests0.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_chr22_0.json 2> gpurun_out/r2_bench_chr22_0.err
python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_dense_0.json 2> gpurun_out/r2_bench_dense_0.err
python bench.py --workload arabidopsis_shaped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_ara_0.json 2> gpurun_out/r2_bench_ara_0.err
python bench.py --workload human_wg --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench_human_0.json 2> gpurun_out/r2_bench_human_0.err
tail -3 gpurun_out/r2_tests0.log
