(cd scripts/micro; for pf in 0 74 148 222 296 444; do echo -n "pf $pf: "; timeout 60 ./pk_cell_bench 1600 160 8 4 $pf | tail -1; done 2>&1 | tee ../../gpurun_out/r02v_pk_micro5.log)
python scripts/bench_configs.py --configs cfg3 --steps 30 2>&1 | tail -2 | tee gpurun_out/r02v_cfg3.jsonl
DFLO_B200_KTRACE=1 python scripts/bench_configs.py --configs cfg3 --steps 10 2>&1 | grep ktrace | tee gpurun_out/r02v_ktrace_cfg3.log
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) | tee gpurun_out/r02v_pytest_gpu.log
