(timeout 1500 python -m pytest tests/test_anchors.py tests/test_gpu_parity.py -m gpu -q -k "convergence or sod_vs or horizon" 2>&1 | tail -5) > gpurun_out/r02b_new_tests.log
tail -5 gpurun_out/r02b_new_tests.log
(time timeout 900 python bench.py 2> gpurun_out/r02b_bench.err | tail -1 > gpurun_out/r02b_bench.json) 2>&1 | grep real
tail -5 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench.json'))
print('value',d['value'],'k_ms',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'])
print('linf',{k:v for k,v in d['linf_vs_ref'].items() if k in('rhs','step1','step20','limiter_flips','within_tolerance')})
for c in d['configs']:
    print(c['config'],c['mdof_per_s'],c['ms_per_step'],c['roofline']['frac'],c['stage_kernel']['ms'],{k:v for k,v in c.get('linf_vs_ref',{}).items() if k in('rhs','step1','step20','limiter_flips','within_tolerance')},c.get('cpu_baseline',{}).get('value'))
PY
