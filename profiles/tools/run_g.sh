set -x
mkdir -p gpurun_out
ONMF_B200_WIDE_ALWAYS=1 DIAG_COLS=0 python profiles/tools/diag_cfg5_fp32.py > gpurun_out/g_diag_wide.log 2>&1
grep -v Warn gpurun_out/g_diag_wide.log | head -12
