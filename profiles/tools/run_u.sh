set -x
mkdir -p gpurun_out
export ONMF_B200_LIB=$PWD/variants/libonmf_b200_fx.so
rm -f gpurun_out/u_cfg.log
for c in 2 3 4; do ONMF_FAST_CFG=$c timeout 120 python profiles/tools/cmp_fast.py 1024 256 262144 3 2>&1 | tail -2 | cut -c1-140 >> gpurun_out/u_cfg.log; done
cat gpurun_out/u_cfg.log
