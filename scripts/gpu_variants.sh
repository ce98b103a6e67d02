#!/bin/bash
# GPU box: parity tests on the default build, then the cfg2 bench on each tuned build variant.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_default.log
for f in build/libdflo_b200_*.so; do
  n=$(basename $f .so)
  (DFLO_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_$n.log
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.0f' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:])
PY
