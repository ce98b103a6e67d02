python scripts/bench_configs.py --configs cfg5 --steps 20 2>&1 | grep config | cut -c1-330 | tee gpurun_out/r02zd_cfgs.jsonl
for c in cfg5; do DFLO_B200_KTRACE=1 python scripts/bench_configs.py --configs $c --steps 6 2>&1 | grep -E "ktrace.*(Limiter|Stage|stage)" ; done | tee gpurun_out/r02zd_ktrace.log
