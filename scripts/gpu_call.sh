P=29700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
(timeout 900 $TR $P tests/multi_gpu_check.py 2>&1 | grep -E "OK|FAIL|Error|error" | tail -14) > gpurun_out/r02u_multi_gpu_check_n2.log; cat gpurun_out/r02u_multi_gpu_check_n2.log
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
