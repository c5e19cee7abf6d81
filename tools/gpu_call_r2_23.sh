#!/bin/bash
# generic track kernels, sub-graphs, device steps of the fragment-length estimate: their GPU tests and the entry-point tests around them
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracks.py tests/test_gpu_shift.py tests/test_gpu_postprocess.py tests/test_gpu_callpeaks_api.py tests/test_gpu_zz_interface.py -q > gpurun_out/r2_tests23.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests23.log
tail -n 40 gpurun_out/r2_tests23.log
