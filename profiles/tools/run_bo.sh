set -x
mkdir -p gpurun_out
( for m in 1 0 1; do ONMF_BCD_REGW=$m timeout 60 python profiles/tools/prof_bcd.py 1024 256 50; done
  for m in 1 0; do ONMF_BCD_REGW=$m timeout 60 python profiles/tools/prof_bcd.py 2048 128 50; done
  for m in 1 0; do ONMF_BCD_REGW=$m timeout 60 python profiles/tools/prof_bcd.py 1500 200 50; done ) > gpurun_out/bo_bcd.log 2>&1
grep -v "^+" gpurun_out/bo_bcd.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -x -k "update_dict or fallbacks or two_ranks or train_matches or fused_step or cfg5" > gpurun_out/bo_pytest.log 2>&1; tail -3 gpurun_out/bo_pytest.log
