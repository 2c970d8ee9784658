set -x
mkdir -p gpurun_out
python profiles/tools/prof_e2e.py > gpurun_out/v_e2e.log 2>&1
tail -6 gpurun_out/v_e2e.log
