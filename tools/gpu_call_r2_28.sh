#!/bin/bash
# p-value kernel without the block-wide miss queue; the bench line with W warm-up steps in the end-to-end loop too
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_control_stats.py tests/test_gpu_callpeaks_api.py tests/test_gpu_reference_known_answers.py tests/test_gpu_configs.py -q > gpurun_out/r2_tests28.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests28.log
tail -n 4 gpurun_out/r2_tests28.log
AB_STEPS=5 bash tools/ab.sh r2_ab28 "k5:" "k5b:"
AB_STEPS=10 bash tools/ab.sh r2_ab28 "k10:"
