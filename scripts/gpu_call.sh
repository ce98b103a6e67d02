python scripts/bench_configs.py --configs cfg4 --steps 20 2>&1 | grep config | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'], d['ms_per_step'], round(d['mdof_per_s']))"
DFLO_B200_KTRACE=1 python scripts/bench_configs.py --configs cfg4 --steps 6 2>&1 | grep -E "ktrace.*(BcEval)"
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
