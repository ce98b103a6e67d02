TAG=r02s NGPU=8 PORT=29660 timeout 1800 bash scripts/gpu_round.sh multi
P=29670
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
DFLO_B200_P2P_DEFER=0 timeout 600 $TR $P bench.py --gpus 8 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('DEFER=0 value', round(d['value']), 'e2e', round(d['e2e']['value']), [(s['config'], round(s['mdof_per_s']), round(s['efficiency'],3), s['linf_vs_single']) for s in d['strong']])"
