#!/bin/bash
# GPU box: bench + one full ncu capture of the row stage kernel (cfg2), plus optional workload
mkdir -p gpurun_out
(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_default.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
r = d['roofline']
print('value %.0f ms/step %.4f kernel_ms %.4f warm %.4f frac %.3f e2e %.0f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_l2_warm'], r['frac'], d['e2e']['value']))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'row_stage|stage_persistent|phase_kernel' -s 6 -c 1 -f -o gpurun_out/prof_row \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out | head -20
