#!/bin/bash
# The one GPU job script: what a round-end check runs on a B200 box, selected by words on the command line.
#   scripts/gpu_round.sh tests bench            one GPU: -m gpu tests, the default bench line + the reference arm
#   scripts/gpu_round.sh sanitize               compute-sanitizer memcheck / racecheck / synccheck (scripts/gpu_sanitize.sh)
#   scripts/gpu_round.sh ncu                    launch list + one `ncu --set full` capture of the row stage kernel (cfg2)
#   NGPU=8 scripts/gpu_round.sh multi           torchrun: tests/multi_gpu_check.py + bench.py --gpus $NGPU (run under gpurun --gpus N)
# Everything lands under gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by hand.
mkdir -p gpurun_out
tag=${TAG:-round}
for what in "$@"; do
case $what in
tests)
  timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 || { echo "SMOKE FAILED/HUNG"; exit 1; }
  (timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_pytest_gpu.log; tail -3 gpurun_out/${tag}_pytest_gpu.log ;;
bench)
  (timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | tail -1) > gpurun_out/${tag}_bench.json
  (timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1) > gpurun_out/${tag}_bench_ref.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench.json'))
print('value %.0f ms/step %.4f kernel_ms %.4f frac %.3f e2e %.0f cpu %.1f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline', {}).get('value', 0)))
for c in d.get('configs', []):
    print(c['config'], round(c['mdof_per_s']), round(c['roofline']['frac'], 3), {k: v for k, v in c.get('linf_vs_ref', {}).items() if k in ('rhs', 'step1', 'step20', 'within_tolerance')})
print('reference arm', json.load(open('gpurun_out/${tag}_bench_ref.json'))['value'])
PY
  ;;
sanitize) bash scripts/gpu_sanitize.sh $tag ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs "" --next-rows "" > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'row_stage' -s 7 -c 1 -f -o gpurun_out/${tag}_prof_row \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs "" --next-rows "" > gpurun_out/${tag}_ncu_run.log 2>&1
  ls -la gpurun_out/${tag}_prof_row.ncu-rep ;;
multi)
  N=${NGPU:-2}; P=${PORT:-29611}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
  (timeout 600 $TR $P tests/multi_gpu_check.py 2>&1 | grep -E "OK|FAIL|Error" | tail -12) > gpurun_out/${tag}_multi_gpu_check_n$N.log; cat gpurun_out/${tag}_multi_gpu_check_n$N.log
  (timeout 900 $TR $((P+1)) bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/${tag}_n$N.err | tail -1) > gpurun_out/${tag}_bench_n$N.json
  python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench_n$N.json'))
print('N=$N value %.0f ms/step %.4f b2b %.0f e2e %.0f linf_vs_single %s' % (d['value'], d['ms_per_step'], d['value_back_to_back_no_flush'], d['e2e']['value'], d['linf_vs_single']['value']))
for s in d['strong']:
    print(s['config'], round(s['mdof_per_s']), round(s['mdof_per_s_1gpu_same_box']), round(s['efficiency'], 3), s['linf_vs_single'])
PY
  ;;
esac
done
