set -x
mkdir -p gpurun_out
python profiles/tools/prof_h2d.py > gpurun_out/aj_h2d.log 2>&1; cat gpurun_out/aj_h2d.log | tail -9
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv >> gpurun_out/aj_h2d.log 2>&1; tail -2 gpurun_out/aj_h2d.log
