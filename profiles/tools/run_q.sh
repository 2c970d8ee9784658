set -x
mkdir -p gpurun_out
export ONMF_B200_LIB=$PWD/variants/libonmf_b200_fx.so
for c in 1 2 3 4 5 6; do ONMF_FAST_CFG=$c python profiles/tools/prof_lars.py 1024 256 262144 3 2>&1 | tail -1 | cut -c1-140 >> gpurun_out/q_cfg.log; done
cat gpurun_out/q_cfg.log
