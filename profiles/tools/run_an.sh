set -x
mkdir -p gpurun_out
for c in 2 3; do python bench.py --workload cfg$c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/an_wl_cfg$c.log 2>&1; tail -1 gpurun_out/an_wl_cfg$c.log | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('cfgX', j['config']['workload'][:5], j['ms_per_step'], j['roofline']['ms_per_launch'])"; done
python bench.py --workload cfg3 --batch 40000 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/an_wl_cfg3b.log 2>&1; tail -1 gpurun_out/an_wl_cfg3b.log | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('cfgX 40000', j['ms_per_step'], j['roofline']['ms_per_launch'])"
