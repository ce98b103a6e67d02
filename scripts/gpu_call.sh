timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02e_tests.log; tail -8 gpurun_out/r02e_tests.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --configs "" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-28s'%'$label', 'value %.0f'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'k_ms %.4f warm %.4f'%(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_l2_warm']), 'b2b %.0f'%d['value_back_to_back_no_flush'])"
}
run new X=1
run new_dbg7 DFLO_B200_DBG=7
run new_dbg4 DFLO_B200_DBG=4
run new_3blocks DFLO_B200_ROW_BLOCKS=3
run v1 DFLO_B200_LIB=$PWD/dflo_b200/csrc/libdflo_b200_v1.so
