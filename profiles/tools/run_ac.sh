set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "edge_columns" > gpurun_out/ac_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ac_pytest.log
tail -30 gpurun_out/ac_pytest.log
