#!/bin/bash
# What round 1 left unmeasured, in one gpurun call (run from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# 1. the GPU tests written after round 1's GPU budget was spent (entry points from files,
#    summits, count_unique_reads), then the whole GPU suite;
# 2. bench.py (plain), then its ncu launch list;
# 3. the summit kernel's time next to the host way.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zz_interface.py -q -m gpu > gpurun_out/zz_tests.log 2>&1
echo "zz rc=$?" >> gpurun_out/zz_tests.log
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "all rc=$?" >> gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tools/summits_rate.py > gpurun_out/summits_rate.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/zz_tests.log gpurun_out/gpu_tests.log gpurun_out/summits_rate.log
cat gpurun_out/bench.json
