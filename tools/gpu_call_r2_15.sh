#!/bin/bash
# reduce-then-scan route of the difference-array scan against the single-pass one
mkdir -p gpurun_out
python tools/sample_stage.py prep > gpurun_out/r2_ss15.log 2>&1
python tools/sample_stage.py run "tile sums + scan + tiles" >> gpurun_out/r2_ss15.log 2>&1
GPC_DENSE_SCAN=2 python tools/sample_stage.py run "single pass, chained" >> gpurun_out/r2_ss15.log 2>&1
cat gpurun_out/r2_ss15.log
timeout 900 python -m pytest tests/test_gpu_sample.py tests/test_cyclic.py tests/test_gpu_guards.py tests/test_gpu_edge_cases.py tests/test_gpu_configs.py tests/test_gpu_digests.py -x -q > gpurun_out/r2_tests15.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests15.log
tail -n 3 gpurun_out/r2_tests15.log
GPC_DENSE_SCAN=2 timeout 600 python -m pytest tests/test_gpu_sample.py -x -q > gpurun_out/r2_tests15_v2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests15_v2.log
tail -n 2 gpurun_out/r2_tests15_v2.log
export GPC_BENCH_ALL_KERNELS=1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench15_chr22.json 2> gpurun_out/r2_bench15_chr22.err
