#!/bin/bash
# leaner difference-array scan (sparse emission loop) against the first version, sweep and
# cycle counters of max paths on the dense-variation workload, next-chromosome prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sample.py tests/test_cyclic.py tests/test_gpu_guards.py tests/test_gpu_edge_cases.py tests/test_gpu_configs.py tests/test_gpu_digests.py tests/test_gpu_multigraph.py -x -q > gpurun_out/r2_tests11.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests11.log
tail -n 5 gpurun_out/r2_tests11.log
for v in 0 1; do
  GPC_DENSE_SCAN=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench11_chr22_v$v.json 2> gpurun_out/r2_bench11_chr22_v$v.err
done
GPC_MP_STATS=1 timeout 600 python bench.py --workload dense_variation --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_bench11_dense.json 2> gpurun_out/r2_bench11_dense.err
timeout 600 python bench.py --workload human_wg --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench11_human_wg.json 2> gpurun_out/r2_bench11_human_wg.err
grep -m1 -A12 "max_paths:" gpurun_out/r2_bench11_dense.err
