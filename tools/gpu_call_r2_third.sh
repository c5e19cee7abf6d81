#!/bin/bash
# round 2, third GPU call: suite, stream timeline of one chr22 pass, bench, e2e host profile,
# ncu --set full of the two sample kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests2.log
tail -4 gpurun_out/r2_tests2.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_chr22_2.json 2> gpurun_out/r2_bench_chr22_2.err
python tools/e2e_breakdown.py chr22_scale > gpurun_out/r2_e2e_chr22.log 2>&1
python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_dense_2.json 2> gpurun_out/r2_bench_dense_2.err
python tools/timeline.py chr22_scale 1 gpurun_out/r2_timeline_chr22.txt > gpurun_out/r2_timeline_chr22.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_runs|walk_reads' -s 8 -c 2 -o gpurun_out/r2_prof_sample python tools/timeline.py chr22_scale 1 /tmp/tl.txt > gpurun_out/r2_ncu_sample.log 2>&1
