#!/bin/bash
# GPU box: parity tests after the graph-reuse change, then e2e with 2 / 3 / 4 batches in flight
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r01f2_pytest_gpu.log
for n in 2 3 4; do
  (DFLO_BENCH_E2E_CTX=$n timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r01f2_bench_ctx$n.log
done
tail -3 gpurun_out/r01f2_pytest_gpu.log
python - <<'PY'
import json
for n in (2, 3, 4):
    try:
        d = json.loads(open('gpurun_out/r01f2_bench_ctx%d.log' % n).read().strip().splitlines()[-1])
        print(n, 'value %.0f' % d['value'], 'e2e', d['e2e'])
    except Exception as e:
        print(n, 'ERR', e)
PY
