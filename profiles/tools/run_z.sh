set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier" > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log
tail -5 gpurun_out/z_pytest.log
rm -f gpurun_out/z_cfg.log
for c in 0 1 2 3 4; do ONMF_B200_LIB=$PWD/variants/libonmf_b200_fx.so ONMF_FAST_CFG=$c python profiles/tools/cmp_fast.py 400 100 16384 5 2>&1 | tail -3 | cut -c1-120 >> gpurun_out/z_cfg.log; done
cat gpurun_out/z_cfg.log
python bench.py --workload cfg4 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/z_wl_cfg4.log 2>&1; tail -1 gpurun_out/z_wl_cfg4.log | cut -c1-250
