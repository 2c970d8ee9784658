set -x
mkdir -p gpurun_out
python profiles/tools/cmp_fast.py 1024 256 262144 > gpurun_out/o_cmp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier or cfg5 or large_active or kkt" > gpurun_out/o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest.log
cat gpurun_out/o_cmp.log; tail -15 gpurun_out/o_pytest.log
