TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port"
for numa in 0 1; do
DFLO_BENCH_NUMA=$numa timeout 300 $TR 2961$numa bench.py --gpus 4 --steps 20 --warmup 3 --strong "" --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('numa $numa N=4 value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done | tee gpurun_out/r02h_numa_n4.log
nvidia-smi topo -m | head -14 >> gpurun_out/r02h_numa_n4.log
