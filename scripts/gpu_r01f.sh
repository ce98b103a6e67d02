#!/bin/bash
# GPU box, round 1 session f: parity tests (incl. the run-loop output schedule), smoke, default bench (pipelined e2e),
# and the standalone driver writing the reference's numbered VTU files.
mkdir -p gpurun_out
rm -f gpurun_out/r01f_*.log
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r01f_pytest_gpu.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r01f_smoke.log
(timeout 400 python bench.py --steps 50 --warmup 5 2>&1 | tail -3) > gpurun_out/r01f_bench.log
mkdir -p /tmp/r01f_out
cp tests/golden/prm/cfg3_sod_P2_hllc_tvb_pos.prm /tmp/r01f_out/input.prm
printf 'subsection output\n  set iter step = 20\n  set schlieren plot = true\nend\n' >> /tmp/r01f_out/input.prm
(cd /tmp/r01f_out && timeout 120 $GRAFT_REPO_ROOT/dflo_b200/csrc/dflo_b200 input.prm \
    --mesh "sod_tube 100 10" --steps 40 --quiet --output-dir . 2>&1 | tail -8; ls -la; head -c 600 solution-001.vtu; grep -c . solution-001.vtu) > gpurun_out/r01f_cli.log 2>&1
tail -3 gpurun_out/r01f_pytest_gpu.log; cat gpurun_out/r01f_smoke.log; cat gpurun_out/r01f_bench.log; cat gpurun_out/r01f_cli.log | head -40
