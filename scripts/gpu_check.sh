#!/bin/bash
# Run on the GPU box via gpurun: parity tests, bench, launch list, one full ncu capture of the stage kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/smoke.log
(timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -5) > gpurun_out/bench.log
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -3) > gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_kernel -s 6 -c 3 -o gpurun_out/prof_stage \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.log
