#!/bin/bash
# GPU box (1 GPU): parity tests, default bench line, reference arm
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r01f3_pytest_gpu.log
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/r01f3_bench.log
tail -3 gpurun_out/r01f3_pytest_gpu.log; cut -c1-1500 gpurun_out/r01f3_bench.log
