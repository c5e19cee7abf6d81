#!/bin/bash
# p-values with the per-tile table of distinct pairs: parity tests, then A/B on one box (table on / off, wire form delta / plain)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_control_stats.py tests/test_gpu_callpeaks_api.py tests/test_gpu_configs.py tests/test_gpu_digests.py tests/test_gpu_reference_known_answers.py -q > gpurun_out/r2_tests25.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests25.log
tail -n 8 gpurun_out/r2_tests25.log
bash tools/ab.sh r2_ab25 "base:GPC_PV_DEDUPE=0 GPC_WIRE_DELTA=0" "dedupe:GPC_WIRE_DELTA=0" "both:" "base2:GPC_PV_DEDUPE=0 GPC_WIRE_DELTA=0" "both2:"
