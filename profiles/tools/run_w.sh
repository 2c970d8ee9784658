set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier or large_active_sets" > gpurun_out/w_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/w_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier" > gpurun_out/w_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/w_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier" > gpurun_out/w_synccheck.log 2>&1; echo "rc=$?" >> gpurun_out/w_synccheck.log
tail -5 gpurun_out/w_memcheck.log; tail -5 gpurun_out/w_racecheck.log; tail -5 gpurun_out/w_synccheck.log
