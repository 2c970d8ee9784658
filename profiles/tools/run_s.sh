set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/s_bench.log 2>&1
tail -6 gpurun_out/s_pytest.log; tail -1 gpurun_out/s_bench.log | cut -c1-600
