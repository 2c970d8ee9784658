set -x
mkdir -p gpurun_out
python profiles/tools/cmp_fast.py 1024 256 262144 3 > gpurun_out/al_cmp.log 2>&1
tail -2 gpurun_out/al_cmp.log | cut -c1-120
