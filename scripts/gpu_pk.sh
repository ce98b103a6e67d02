#!/bin/bash
# GPU box: Pk cell stage kernel -- parity tests of the Pk cases, cfg3 bench (cell vs tile), optional ncu capture
mkdir -p gpurun_out
if [ "$1" != "ncuonly" ]; then
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python scripts/bench_configs.py --configs cfg3 2>&1 | tail -1 | cut -c1-420
fi
if [ "$1" = "ncu" ] || [ "$1" = "ncuonly" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'PkCellStage' -s 7 -c 1 -f -o gpurun_out/prof_pkcell \
    python scripts/bench_configs.py --configs cfg3 --steps 3 > gpurun_out/ncu_pkcell_run.log 2>&1
tail -2 gpurun_out/ncu_pkcell_run.log
fi
