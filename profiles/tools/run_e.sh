set -x
mkdir -p gpurun_out
python profiles/tools/diag_cfg5_fp32.py > gpurun_out/e_diag_tc.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
for w in cfg1 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline > gpurun_out/e_wl_$w.log 2>&1; done
timeout 300 python bench.py --workload cfg1 --steps 200 --no-cpu-baseline --graph off > gpurun_out/e_wl_cfg1_nograph.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/e_bench.log 2>&1
grep -v Warn gpurun_out/e_diag_tc.log | tail -70; tail -8 gpurun_out/e_pytest.log; cat gpurun_out/e_wl_*.log | cut -c1-260
