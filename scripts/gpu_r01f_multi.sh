#!/bin/bash
# GPU box with N GPUs: sharded-vs-single parity check including the sharded driver output
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/r01f_multi_check_$N.log 2>&1
echo "multi_gpu_check rc=$?"; grep -E "OK|FAIL|Error|error|assert" gpurun_out/r01f_multi_check_$N.log | tail -14
