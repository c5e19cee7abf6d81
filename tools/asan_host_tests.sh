#!/bin/bash
# The host code of the library (parsers, splitter, peak model, linear map) under
# AddressSanitizer: builds an instrumented copy of the library next to the real one
# (/tmp/libgpc_asan.so) and runs the CPU test suite -- fuzz tests included -- against it.
# No GPU involved.  Round 1: 174 passed, no report.
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 \
    -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -shared \
    -o /tmp/libgpc_asan.so graph_peak_caller_b200/csrc/*.cu
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
    GPC_B200_LIB=/tmp/libgpc_asan.so \
    python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider \
    --deselect tests/test_cabi_exports.py
