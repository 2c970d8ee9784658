set -x
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -40 gpurun_out/b_pytest.log
