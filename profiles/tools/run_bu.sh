set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bu_bench.log 2>&1
tail -1 gpurun_out/bu_bench.log | cut -c1-1600
