#!/bin/bash
# GPU box (1 GPU): parity tests only (final check of the output-format change)
mkdir -p gpurun_out
(timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r01f4_pytest_gpu.log
tail -4 gpurun_out/r01f4_pytest_gpu.log
