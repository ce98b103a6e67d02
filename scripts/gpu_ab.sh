#!/bin/bash
# GPU box: smoke (short timeout as a hang guard), parity tests, bench persistent vs one-tile-per-block, ncu of the default.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG"; exit 1; }
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_default.log
(DFLO_B200_PERSISTENT=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_nonpersistent.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.0f' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:])
PY
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stage_persistent|phase_kernel' -s 6 -c 2 -f -o gpurun_out/prof_stage \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
fi
