set -x
mkdir -p gpurun_out
python profiles/tools/prof_storage.py > gpurun_out/v_storage.log 2>&1
cat gpurun_out/v_storage.log | tail -5
