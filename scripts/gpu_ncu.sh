#!/bin/bash
# GPU box: parity tests, bench, and one full ncu capture of the stage kernel (cfg2).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_default.log
cat gpurun_out/bench_default.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_kernel -s 6 -c 2 -f -o gpurun_out/prof_stage \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
