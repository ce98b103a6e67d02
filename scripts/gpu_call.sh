nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash scripts/gpu_sanitize.sh r02a
(timeout 400 python bench.py --steps 50 --warmup 5 2>&1 | tail -1) > gpurun_out/r02a_bench_default.json
python -c "
import json; d=json.load(open('gpurun_out/r02a_bench_default.json')); print(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
