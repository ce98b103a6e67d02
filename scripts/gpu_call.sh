P=29593
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
(timeout 600 $TR $P tests/multi_gpu_check.py 2>&1 | grep -c OK) 
P=$((P+1))
run2() { label=$1; shift
  env "$@" timeout 600 $TR $P bench.py --gpus 2 --steps 50 --warmup 5 --strong "$STRONG" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-20s'%'$label', 'N=2 value %.0f'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'b2b %.0f'%d['value_back_to_back_no_flush'], 'linf', d['linf_vs_single']['value'], [(s['config'], round(s['mdof_per_s']), round(s['mdof_per_s_1gpu_same_box']), round(s['efficiency'],3), s['linf_vs_single']) for s in d['strong']])"
  P=$((P+1))
}
STRONG="" run2 n2 X=1
STRONG="cfg4,cfg5" run2 n2 X=1
STRONG="" run2 n2_fixed_dt DFLO_BENCH_FIXED_DT=0.002
