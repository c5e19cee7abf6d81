#!/bin/bash
# round 2, second GPU call: sample-pileup routes (difference array / event list), the whole suite,
# bench lines, e2e host profile of the dense-variation workload
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sample.py tests/test_gpu_edge_cases.py tests/test_gpu_postprocess.py -x -q > gpurun_out/r2_tests1a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests1a.log
tail -5 gpurun_out/r2_tests1a.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_1.json 2> gpurun_out/r2_bench_chr22_1.err
python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_dense_1.json 2> gpurun_out/r2_bench_dense_1.err
python bench.py --workload human_wg --scale 4 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench_human4_1.json 2> gpurun_out/r2_bench_human4_1.err
python tools/e2e_breakdown.py dense_variation 4 > gpurun_out/r2_e2e_dense.log 2>&1
python tools/e2e_breakdown.py human_wg 8 0 > gpurun_out/r2_e2e_human.log 2>&1
