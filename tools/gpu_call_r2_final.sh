#!/bin/bash
# End of round 2: the whole GPU suite, the default bench line as the driver runs it (with the
# reference arm's child process), the ncu launch list of the same bench command and one
# --set full capture of the p-value kernel (each ncu pass only after the plain run exited 0).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2_tests_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_final.log
tail -n 4 gpurun_out/r2_tests_final.log
export GPC_BENCH_ALL_KERNELS=1
timeout 600 python bench.py > gpurun_out/r2_bench_final_chr22.json 2> gpurun_out/r2_bench_final_chr22.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_final_chr22.json; echo
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_bench_final_short.json 2> gpurun_out/r2_bench_final_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_final.log 2>&1
echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pv_eval' -s 3 -c 1 -o gpurun_out/r2_prof_final_pv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_final_pv.log 2>&1
echo "ncu full rc=$?"
