set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cabi.py -q -x -k "spectral_norm or pgd or shipped or reconstruction or cabi or exports or symbol" > gpurun_out/am_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/am_pytest.log
tail -25 gpurun_out/am_pytest.log
