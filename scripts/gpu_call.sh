python bench.py --configs "" --next-rows q1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for c in d['configs']: print(c['config'], c['ms_per_step'], round(c['mdof_per_s']), c['roofline']['frac'])
" | tee gpurun_out/r02zg_q1.log
