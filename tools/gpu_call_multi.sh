#!/bin/bash
# multi-GPU call: N = number of GPUs of the box.  The torchrun-launched correctness test of the
# sharded join, then the human_wg workload (24 chromosomes, LPT packing, strong scaling).
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2_multirank_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multirank_n$N.log
tail -n 3 gpurun_out/r2_multirank_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload human_wg --steps ${2:-10} --warmup 3 --no-cpu > gpurun_out/r2_bench_human_n$N.json 2> gpurun_out/r2_bench_human_n$N.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_human_n$N.json
