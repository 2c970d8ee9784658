set -x
mkdir -p gpurun_out
( for m in 2 1 2; do ONMF_BCD_MBAR=$m timeout 60 python profiles/tools/prof_bcd.py 1024 256 50; echo "mode=$m rc=$?"; done
  ONMF_BCD_MBAR=2 ONMF_BCD_REGW=0 timeout 60 python profiles/tools/prof_bcd.py 1024 256 50
  ONMF_BCD_MBAR=2 timeout 60 python profiles/tools/prof_bcd.py 2048 128 50
  ONMF_BCD_MBAR=2 timeout 60 python profiles/tools/prof_bcd.py 1500 200 50 ) > gpurun_out/bp_bcd.log 2>&1
grep -v "^+" gpurun_out/bp_bcd.log
ONMF_BCD_MBAR=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -x -k "update_dict or fallbacks or two_ranks or train_matches or fused_step" > gpurun_out/bp_pytest.log 2>&1; tail -3 gpurun_out/bp_pytest.log
