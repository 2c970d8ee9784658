set -x
mkdir -p gpurun_out
DIAG_COLS=0 python profiles/tools/diag_cfg5_fp32.py > gpurun_out/f_diag_tc.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
grep -v Warn gpurun_out/f_diag_tc.log | head -14; tail -8 gpurun_out/f_pytest.log
