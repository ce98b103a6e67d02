for c in cfg3 cfg4 cfg5; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02j_launches_$c.csv python scripts/bench_configs.py --configs $c --steps 3 > /dev/null 2>&1
done
ls -la gpurun_out/r02j_launches_*.csv
