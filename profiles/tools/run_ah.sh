set -x
mkdir -p gpurun_out
python profiles/tools/cmp_fast.py 1024 512 65536 3 > gpurun_out/ah_cmp512.log 2>&1
python profiles/tools/cmp_fast.py 1024 192 131072 3 > gpurun_out/ah_cmp192.log 2>&1
python profiles/tools/cmp_fast.py 512 128 131072 3 > gpurun_out/ah_cmp128.log 2>&1
tail -3 gpurun_out/ah_cmp512.log | cut -c1-250; tail -3 gpurun_out/ah_cmp192.log | cut -c1-250; tail -3 gpurun_out/ah_cmp128.log | cut -c1-250
