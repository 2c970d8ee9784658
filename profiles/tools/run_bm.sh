set -x
mkdir -p gpurun_out
( for m in 0 1 0 1; do ONMF_BCD_MBAR=$m timeout 60 python profiles/tools/prof_bcd.py 1024 256 50; echo "mbar=$m rc=$?"; done
  ONMF_BCD_MBAR=1 timeout 60 python profiles/tools/prof_bcd.py 2048 128 50; ONMF_BCD_MBAR=0 timeout 60 python profiles/tools/prof_bcd.py 2048 128 50
  ONMF_BCD_SMALL=0 ONMF_BCD_MBAR=1 timeout 60 python profiles/tools/prof_bcd.py 400 100 50; ONMF_BCD_SMALL=0 ONMF_BCD_MBAR=0 timeout 60 python profiles/tools/prof_bcd.py 400 100 50 ) > gpurun_out/bm_bcd.log 2>&1
cat gpurun_out/bm_bcd.log | grep -v "^+"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "update_dict or fallbacks" > gpurun_out/bm_pytest.log 2>&1; tail -3 gpurun_out/bm_pytest.log
