python scripts/bench_configs.py --configs cfg3,cfg4,cfg5 --steps 20 2>&1 | grep config | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'], d['ms_per_step'], round(d['mdof_per_s']))" | tee gpurun_out/r02i_cfgs.log
for c in cfg4 cfg5; do DFLO_B200_KTRACE=1 python scripts/bench_configs.py --configs $c --steps 6 2>&1 | grep -E "ktrace.*(Limiter)" ; done | tee gpurun_out/r02i_ktrace.log
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) | tee gpurun_out/r02i_pytest_gpu.log
