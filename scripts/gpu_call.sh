python scripts/bench_configs.py --configs cfg3,cfg4,cfg5 --steps 20 2>&1 | grep config | cut -c1-330 | tee gpurun_out/r02zc_cfgs.jsonl
for c in cfg3 cfg4 cfg5; do DFLO_B200_KTRACE=1 python scripts/bench_configs.py --configs $c --steps 6 2>&1 | grep -E "ktrace.*(Limiter|Stage|stage)" ; done | tee gpurun_out/r02zc_ktrace.log
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/r02zc_pytest_gpu.log
