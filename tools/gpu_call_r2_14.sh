#!/bin/bash
# where the difference-array scan spends its time: the kernel with parts switched off
# (GPC_DENSE_DBG, results invalid), and ncu --set full of the complete one
mkdir -p gpurun_out
python tools/sample_stage.py prep > gpurun_out/r2_ss14.log 2>&1
python tools/sample_stage.py run "default" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_SCAN=1 python tools/sample_stage.py run "first version" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=4 python tools/sample_stage.py run "no ticket" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=1 python tools/sample_stage.py run "no look-back" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=2 python tools/sample_stage.py run "no emission" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=3 python tools/sample_stage.py run "no look-back, no emission" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=7 python tools/sample_stage.py run "+ no ticket" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=16 python tools/sample_stage.py run "load + zero only" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=20 python tools/sample_stage.py run "load + zero only, no ticket" >> gpurun_out/r2_ss14.log 2>&1
GPC_DENSE_DBG=28 python tools/sample_stage.py run "load only, no ticket" >> gpurun_out/r2_ss14.log 2>&1
cat gpurun_out/r2_ss14.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'dense_runs' -s 4 -c 1 -o gpurun_out/r2_prof14 python tools/sample_stage.py run ncu > gpurun_out/r2_ncu14.log 2>&1
tail -c 200 gpurun_out/r2_ncu14.log
