(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compression_corner or local_time" 2>&1 | tail -4)
(time timeout 900 python bench.py 2> gpurun_out/r02o_bench.err | tail -1 > gpurun_out/r02o_bench.json) 2>&1 | grep real
tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02o_bench.json'))
print('value',round(d['value']),'ms/step',d['ms_per_step'],'k_ms',d['roofline']['kernel_ms'],'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value'],1),d['cpu_baseline']['cores'])
print('linf',{k:v for k,v in d['linf_vs_ref'].items() if k in('rhs','step1','step20','limiter_flips','within_tolerance')})
for c in d['configs']:
    print(c['config'],round(c['mdof_per_s']),round(c['ms_per_step'],4),round(c['roofline']['frac'],3),round(c['stage_kernel']['ms'],4),{k:v for k,v in c.get('linf_vs_ref',{}).items() if k in('rhs','step1','step20','limiter_flips','within_tolerance')},round(c.get('cpu_baseline',{}).get('value',0),1))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02o_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs "" --next-rows "" > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'row_stage' -s 7 -c 1 -f -o gpurun_out/r02o_prof_row python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs "" --next-rows "" > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/*.ncu-rep
