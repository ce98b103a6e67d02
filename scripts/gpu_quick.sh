#!/bin/bash
# GPU box: cfg2 bench only, on the default build, each build/libdflo_b200_*.so and each env in $DFLO_VARIANT_ENVS
mkdir -p gpurun_out
rm -f gpurun_out/bench_*.log
run() { (env $2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $DFLO_BENCH_ARGS 2>&1 | tail -1) > "gpurun_out/bench_$1.log"; }
run default ""
for f in build/libdflo_b200_*.so; do [ -f "$f" ] || continue; run $(basename $f .so) "DFLO_B200_LIB=$PWD/$f"; done
IFS=';' read -ra ENVS <<< "$DFLO_VARIANT_ENVS"
i=0
for e in "${ENVS[@]}"; do i=$((i+1)); run "env$i" "$e"; echo "env$i: $e"; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'warm %.4f' % d['roofline']['kernel_ms_l2_warm'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.0f' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-500:])
PY
