#!/bin/bash
# multi-GPU call, end of round 2: N = number of GPUs of the box.  The torchrun-launched
# correctness test of the sharded join, the human_wg workload (24 chromosomes, LPT packing,
# strong scaling) and -- at N = 8 -- the default line (one chr22-scale chromosome per rank).
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2b_multirank_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_multirank_n$N.log
tail -n 3 gpurun_out/r2b_multirank_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload human_wg --steps ${2:-10} --warmup 3 --no-cpu > gpurun_out/r2b_bench_human_n$N.json 2> gpurun_out/r2b_bench_human_n$N.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench_human_n$N.json
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_chr22_n$N.json 2> gpurun_out/r2b_bench_chr22_n$N.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench_chr22_n$N.json
fi
