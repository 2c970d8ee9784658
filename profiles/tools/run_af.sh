set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_compat.py -m gpu -q > gpurun_out/af_multirank.log 2>&1; echo "rc=$?" >> gpurun_out/af_multirank.log
tail -5 gpurun_out/af_multirank.log
