#!/bin/bash
# GPU box with N GPUs: weak-scaling bench lines for each env in $DFLO_VARIANT_ENVS
N=${1:-2}
mkdir -p gpurun_out
IFS=';' read -ra ENVS <<< "$DFLO_VARIANT_ENVS"
i=0
for e in "" "${ENVS[@]}"; do
  env $e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+i)) \
      bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_v$i.log 2>&1
  echo "v$i [$e] rc=$?"; tail -1 gpurun_out/bench_n${N}_v$i.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value %.0f ms/step %.4f b2b %.0f e2e %.0f' % (d['value'], d['ms_per_step'], d['value_back_to_back_no_flush'], d['e2e']['value']))
except Exception as ex: print('  ERR', ex)
"
  i=$((i+1))
done
