#!/bin/bash
# batched memo probes, parallel read-back kernel: parity tests and A/B on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_control_stats.py tests/test_gpu_callpeaks_api.py tests/test_gpu_primitives.py tests/test_gpu_postprocess.py -q > gpurun_out/r2_tests27.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests27.log
tail -n 5 gpurun_out/r2_tests27.log
export AB_STEPS=6
bash tools/ab.sh r2_ab27 "old:GPC_PV_VARIANT=3 GPC_RB_SERIAL=1" "probes:GPC_RB_SERIAL=1" "rb:GPC_PV_VARIANT=3" "both:" "old2:GPC_PV_VARIANT=3 GPC_RB_SERIAL=1" "both2:"
