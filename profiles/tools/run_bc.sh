set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/bc_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bc_pytest.log
python bench.py --workload next > gpurun_out/bc_next.log 2>&1
tail -5 gpurun_out/bc_pytest.log; cut -c1-250 gpurun_out/bc_next.log
