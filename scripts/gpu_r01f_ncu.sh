#!/bin/bash
# GPU box: ncu launch list of the bench command + one full capture of the row stage kernel (r01f build)
mkdir -p gpurun_out
export DFLO_BENCH_E2E_CTX=1   # no worker threads under the profiler
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches_cfg2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_ncu_launches_run.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:row_stage -s 6 -c 1 -f -o gpurun_out/prof_row_r01f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_ncu_full_run.log 2>&1
ls -la gpurun_out/prof_row_r01f.ncu-rep gpurun_out/r01f_launches_cfg2.csv; tail -2 gpurun_out/r01f_ncu_full_run.log | cut -c1-300
