set -x
mkdir -p gpurun_out
python profiles/tools/diag_cfg5_fp32.py > gpurun_out/d_diag_tc.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x -k "network or fused_step or cfg1 or full_length or train_matches or empty" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
for w in cfg1 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline > gpurun_out/d_wl_$w.log 2>&1; done
ONMF_B200_GRAPH=0 timeout 300 python bench.py --workload cfg1 --steps 200 --no-cpu-baseline > gpurun_out/d_wl_cfg1_nograph.log 2>&1
grep -v Warn gpurun_out/d_diag_tc.log | tail -14; tail -8 gpurun_out/d_pytest.log; cat gpurun_out/d_wl_*.log | cut -c1-260
