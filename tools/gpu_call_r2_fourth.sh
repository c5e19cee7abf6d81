#!/bin/bash
# round 2, fourth GPU call: warp-centric difference-array scan, compact uploads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sample.py tests/test_gpu_callpeaks_api.py -x -q > gpurun_out/r2_tests3a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests3a.log
tail -4 gpurun_out/r2_tests3a.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log
tail -4 gpurun_out/r2_tests3.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_3.json 2> gpurun_out/r2_bench_chr22_3.err
python tools/e2e_breakdown.py chr22_scale > gpurun_out/r2_e2e_chr22_3.log 2>&1
python bench.py --workload arabidopsis_shaped --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_ara_3.json 2> gpurun_out/r2_bench_ara_3.err
