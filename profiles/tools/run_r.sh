set -x
mkdir -p gpurun_out
python profiles/tools/cmp_fast.py 1024 256 262144 3 > gpurun_out/r_cmp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier" > gpurun_out/r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r_pytest.log
ncu --set full --clock-control none --import-source on -k regex:lars_fast_kernel -s 6 -c 1 -o gpurun_out/r2_lars_fast_v4 -f python profiles/tools/prof_lars.py 1024 256 262144 2 > gpurun_out/r_ncu.log 2>&1
cat gpurun_out/r_cmp.log; tail -5 gpurun_out/r_pytest.log
