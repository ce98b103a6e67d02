#!/bin/bash
# GPU box with N GPUs: strong scaling of cfg4 (and cfg2) sharded over the ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    scripts/bench_configs.py --configs ${2:-cfg4,cfg5,cfg2} > gpurun_out/strong_n$N.log 2>&1
echo "rc=$?"; grep '"config"' gpurun_out/strong_n$N.log | cut -c1-330; tail -3 gpurun_out/strong_n$N.log | grep -v config | cut -c1-300
