python bench.py --configs "" --next-rows q1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for c in d['configs']: print(c['config'], c['ms_per_step'], round(c['mdof_per_s']), c['roofline']['frac'], c.get('linf_vs_ref',{}).get('within_tolerance'))
print('cfg2', d['value'], d['ms_per_step'])
" | tee gpurun_out/r02zf_q1.log
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/r02zf_pytest_gpu.log
