#!/bin/bash
# q-value lookup through a per-call hash table: parity (digest of the chr22-scale pass, whole-path tests), A/B against the binary search
mkdir -p gpurun_out
GPC_Q_HASH=1 timeout 200 python -m pytest tests/test_gpu_digests.py tests/test_gpu_control_stats.py tests/test_gpu_callpeaks_api.py -q -k "not dense and not arabidopsis and not human" > gpurun_out/r2_tests30.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests30.log
tail -n 4 gpurun_out/r2_tests30.log
AB_STEPS=6 bash tools/ab.sh r2_ab30 "hash:GPC_Q_HASH=1" "search:"
