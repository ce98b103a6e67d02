#!/bin/bash
# GPU box: A/B of library builds (build/libdflo_b200_*.so + the in-tree one) on cfg2 (bench.py) and cfg4 (bench_configs.py)
mkdir -p gpurun_out
for f in dflo_b200/csrc/libdflo_b200.so build/libdflo_b200_*.so; do
  [ -f "$f" ] || continue
  echo "== $f"
  DFLO_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  cfg2 value %.0f ms/step %.4f kernel_ms %.4f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']))"
  DFLO_B200_LIB=$PWD/$f timeout 300 python scripts/bench_configs.py --configs ${DFLO_AB_CONFIGS:-cfg4} 2>&1 | tail -1 | cut -c1-330
done
