set -x
mkdir -p gpurun_out
python profiles/tools/diag_cfg5_fp32.py > gpurun_out/c_diag_tc.log 2>&1
DIAG_TC=0 python profiles/tools/diag_cfg5_fp32.py > gpurun_out/c_diag_simt.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
cat gpurun_out/c_diag_tc.log gpurun_out/c_diag_simt.log; tail -8 gpurun_out/c_pytest.log
