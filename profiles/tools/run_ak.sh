set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -q -x -k "not full_length and not cfg5_full_run and not kkt_at_scale" > gpurun_out/ak_memcheck_all.log 2>&1; echo "rc=$?" >> gpurun_out/ak_memcheck_all.log
tail -8 gpurun_out/ak_memcheck_all.log
