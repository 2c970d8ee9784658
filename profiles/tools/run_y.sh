set -x
mkdir -p gpurun_out
python profiles/tools/cmp_fast.py 1024 256 262144 3 > gpurun_out/y_cmp.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
cat gpurun_out/y_cmp.log; tail -4 gpurun_out/y_pytest.log
