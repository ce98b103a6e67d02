#!/bin/bash
# GPU box, one call: smoke, parity tests, default bench (full line incl. clocks + cpu_baseline), per-config bench
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG"; exit 1; }
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/bench_default.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
r = d['roofline']
print('value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'e2e %.0f' % d['e2e']['value'],
      'cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
(timeout 600 python scripts/bench_configs.py 2>&1 | tail -8) > gpurun_out/bench_configs.log
cat gpurun_out/bench_configs.log | cut -c1-400
