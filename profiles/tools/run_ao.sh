set -x
mkdir -p gpurun_out
python profiles/tools/prof_steps.py > gpurun_out/ao_steps.log 2>&1; tail -27 gpurun_out/ao_steps.log
