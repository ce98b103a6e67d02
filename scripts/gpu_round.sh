#!/bin/bash
# GPU box, one call: smoke (hang guard), parity tests, bench (default + one-tile-per-block form),
# reference arm, launch list, one full ncu capture of the dominant (stage) kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG"; exit 1; }
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py --steps 50 --warmup 5 2>&1 | tail -1) > gpurun_out/bench_default.log
(DFLO_B200_PERSISTENT=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_nonpersistent.log
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1) > gpurun_out/bench_ref.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline', {})
        print(f, 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % r.get('kernel_ms', 0), 'frac %.3f' % r.get('frac', 0), 'e2e %.0f' % d['e2e']['value'], 'cpu', d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:])
PY
if [ "$1" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'row_stage|stage_persistent|phase_kernel' -s 6 -c 2 -f -o gpurun_out/prof_stage \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
fi
ls -la gpurun_out
