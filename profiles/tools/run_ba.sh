set -x
mkdir -p gpurun_out
python bench.py --workload next > gpurun_out/ba_next.log 2>&1
tail -40 gpurun_out/ba_next.log | cut -c1-400
