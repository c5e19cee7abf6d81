#!/bin/bash
# one-pass topological max paths with class lists: whole suite, dense-variation and chr22
# bench lines; ncu --set full of the staged difference-array scan, the walk and the p-values
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests12.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests12.log
tail -n 5 gpurun_out/r2_tests12.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench12_chr22.json 2> gpurun_out/r2_bench12_chr22.err
timeout 600 python bench.py --workload dense_variation --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench12_dense.json 2> gpurun_out/r2_bench12_dense.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dense_runs|walk_reads|pv_eval' -s 12 -c 3 -o gpurun_out/r2_prof12 python tools/timeline.py chr22_scale 1 /tmp/tl.txt > gpurun_out/r2_ncu12.log 2>&1
tail -c 300 gpurun_out/r2_ncu12.log
