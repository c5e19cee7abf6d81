#!/bin/bash
# p-value kernel: batched tile loads / staged stores, A/B of the four combinations and the no-memo timing experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_control_stats.py tests/test_gpu_callpeaks_api.py tests/test_gpu_reference_known_answers.py -q > gpurun_out/r2_tests26.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests26.log
tail -n 5 gpurun_out/r2_tests26.log
export AB_STEPS=6
bash tools/ab.sh r2_ab26 "v0:GPC_PV_VARIANT=0" "v1:GPC_PV_VARIANT=1" "v2:GPC_PV_VARIANT=2" "v3:GPC_PV_VARIANT=3" "v0nomemo:GPC_PV_VARIANT=0 GPC_PV_DBG=1" "v3nomemo:GPC_PV_VARIANT=3 GPC_PV_DBG=1" "v0b:GPC_PV_VARIANT=0"
