#!/bin/bash
# GPU box, one call: smoke, parity tests, default bench (full line incl. clocks + cpu_baseline), reference arm,
# per-config bench, launch list of the bench command; "ncu": full captures of the cfg2 stage kernel and the Pk cell kernel
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG"; exit 1; }
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/bench_default.log
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1) > gpurun_out/bench_ref.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
r = d['roofline']
print('value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'e2e %.0f' % d['e2e']['value'],
      'cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d['clocks'], 'launches', d['gpu_launches'])
print('ref', json.loads(open('gpurun_out/bench_ref.log').read().strip().splitlines()[-1])['value'])
PY
(timeout 600 python scripts/bench_configs.py 2>&1 | tail -8) > gpurun_out/bench_configs.log
cat gpurun_out/bench_configs.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3.csv \
    python scripts/bench_configs.py --configs cfg3 --steps 3 > gpurun_out/ncu_launches_cfg3_run.log 2>&1
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'row_stage' -s 6 -c 1 -f -o gpurun_out/prof_row \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
fi
ls -la gpurun_out | tail -8
