set -x
mkdir -p gpurun_out
python profiles/tools/cmp_fast.py 1024 256 262144 3 > gpurun_out/al_cmp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier or cfg5_per_step or large_active" > gpurun_out/al_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/al_pytest.log
tail -3 gpurun_out/al_cmp.log | cut -c1-200; tail -3 gpurun_out/al_pytest.log
