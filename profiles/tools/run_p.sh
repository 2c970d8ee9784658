set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lars_fast_kernel -s 6 -c 1 -o gpurun_out/r2_lars_fast_v2 -f python profiles/tools/prof_lars.py 1024 256 262144 2 > gpurun_out/p_ncu.log 2>&1
tail -3 gpurun_out/p_ncu.log
