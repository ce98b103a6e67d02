N=${NGPU:-4}
P=29613
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
(DFLO_B200_KTRACE=1 timeout 600 $TR $P scripts/bench_configs.py --configs cfg4 --steps 10 2>&1 | grep -E "ktrace|mdof" | sed 's/N4dflo//' | sort | cut -c1-160) | tee gpurun_out/r02l_ktrace_cfg4_n$N.log
P=$((P+1))
(timeout 600 $TR $P scripts/bench_configs.py --configs cfg4 --steps 20 2>&1 | grep -E "mdof" | cut -c1-400) | tee gpurun_out/r02l_cfg4_n$N.log
