python scripts/bench_configs.py --configs cfg3,cfg4,cfg5 --steps 20 2>&1 | grep config | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'], d['ms_per_step'], round(d['mdof_per_s']))"
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
