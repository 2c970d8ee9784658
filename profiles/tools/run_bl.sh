set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vector_gathers or thread_per_sample or sklearn_order or patch_pipeline or network or update_dict or test_gathers or gather_index or step_host or narrow" > gpurun_out/bl_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/bl_memcheck.log
tail -8 gpurun_out/bl_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vector_gathers or thread_per_sample or update_dict" > gpurun_out/bl_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/bl_racecheck.log
tail -6 gpurun_out/bl_racecheck.log
