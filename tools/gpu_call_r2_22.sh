#!/bin/bash
# generic track kernels and sub-graphs: their GPU tests, then the callpeaks API tests (Reporter now gets sub_graphs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracks.py tests/test_gpu_postprocess.py tests/test_gpu_callpeaks_api.py -x -q > gpurun_out/r2_tests22.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests22.log
tail -n 25 gpurun_out/r2_tests22.log
