#!/bin/bash
# GPU box with N GPUs: sharded-vs-single parity check and the weak-scaling bench line at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "OK|FAIL|Error|error" gpurun_out/multi_check_$N.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.log | cut -c1-900
