#!/bin/bash
# linear map on the device, warp-per-component max paths, vectorised node-sized scans,
# density switch of the linear control track: whole GPU suite + one bench line per workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests9.log
tail -n 5 gpurun_out/r2_tests9.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench9_chr22.json 2> gpurun_out/r2_bench9_chr22.err
for w in dense_variation human_wg arabidopsis_shaped; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench9_$w.json 2> gpurun_out/r2_bench9_$w.err
done
tail -c 600 gpurun_out/r2_bench9_chr22.json
