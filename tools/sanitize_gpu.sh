#!/bin/bash
# compute-sanitizer over the small GPU parity tests, ONE tool per call (memcheck | racecheck |
# initcheck | synccheck).  PYTORCH_NO_CUDA_MEMORY_CACHING=1: every tensor is its own
# allocation, so an out-of-bounds access cannot hide inside torch's cached blocks.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
TESTS="tests/test_gpu_primitives.py tests/test_gpu_sample.py tests/test_cyclic.py tests/test_gpu_control_stats.py tests/test_gpu_postprocess.py tests/test_gpu_edge_cases.py"
# the plain run first: the sanitizer only makes sense on a green suite
python -m pytest $TESTS -m gpu -x -q -k "not medium and not full_size" > gpurun_out/sanitize_plain_$TOOL.log 2>&1 || { tail -n 20 gpurun_out/sanitize_plain_$TOOL.log; exit 1; }
tail -n 2 gpurun_out/sanitize_plain_$TOOL.log
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 40 \
    --log-file gpurun_out/sanitizer_$TOOL.log \
    python -m pytest $TESTS -m gpu -x -q -k "not medium and not full_size" > gpurun_out/sanitize_pytest_$TOOL.log 2>&1
echo "sanitizer rc=$?"
tail -n 3 gpurun_out/sanitize_pytest_$TOOL.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|Race reported|Uninitialized" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | head -20
