set -x
mkdir -p gpurun_out
for c in 1 3; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/ad_launches_cfg$c.csv python bench.py --workload cfg$c --steps 4 --warmup 5 --no-cpu-baseline --graph off > gpurun_out/ad_cfg$c.log 2>&1
done
