run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --configs "" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-28s'%'$label', 'value %.0f'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'k_ms %.4f warm %.4f'%(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_l2_warm']), 'b2b %.0f'%d['value_back_to_back_no_flush'])"
}
run new X=1
run new_dbg4_noedge DFLO_B200_DBG=4
run new_dbg1_noriemann DFLO_B200_DBG=1
run new_dbg3_noflux DFLO_B200_DBG=3
run new_dbg7_nothing DFLO_B200_DBG=7
run new_pf0 DFLO_B200_PF_TILES=0
run new_pf1480 DFLO_B200_PF_TILES=1480
run new_pf370 DFLO_B200_PF_TILES=370
L=$PWD/dflo_b200/csrc/libdflo_b200_v1.so
run v1 DFLO_B200_LIB=$L
run v1_dbg7 DFLO_B200_LIB=$L DFLO_B200_DBG=7
run v1_pf0 DFLO_B200_LIB=$L DFLO_B200_PF_TILES=0
