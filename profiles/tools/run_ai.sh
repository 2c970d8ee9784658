set -x
mkdir -p gpurun_out
rm -f gpurun_out/ai_soak.log
for a in 0.5 0.2 0.05 0.0; do echo "alpha $a" >> gpurun_out/ai_soak.log; CMP_ALPHA=$a timeout 300 python profiles/tools/cmp_fast.py 1024 256 32768 1 2>&1 | tail -3 | cut -c1-260 >> gpurun_out/ai_soak.log; done
for a in 0.3 0.0; do echo "k=100 alpha $a" >> gpurun_out/ai_soak.log; CMP_ALPHA=$a timeout 300 python profiles/tools/cmp_fast.py 400 100 32768 1 2>&1 | tail -3 | cut -c1-260 >> gpurun_out/ai_soak.log; done
cat gpurun_out/ai_soak.log
