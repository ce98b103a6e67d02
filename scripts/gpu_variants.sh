#!/bin/bash
# GPU box: parity tests on the default build, then the cfg2 bench on each tuned build variant
# (build/libdflo_b200_*.so) and on each env setting listed in $DFLO_VARIANT_ENVS (";"-separated).
mkdir -p gpurun_out
rm -f gpurun_out/bench_*.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_default.log
for f in build/libdflo_b200_*.so; do
  [ -f "$f" ] || continue
  n=$(basename $f .so)
  (DFLO_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_$n.log
done
IFS=';' read -ra ENVS <<< "$DFLO_VARIANT_ENVS"
i=0
for e in "${ENVS[@]}"; do
  i=$((i+1))
  (env $e timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > "gpurun_out/bench_env${i}.log"
  echo "env${i}: $e"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'warm %.4f' % d['roofline']['kernel_ms_l2_warm'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.0f' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:])
PY
