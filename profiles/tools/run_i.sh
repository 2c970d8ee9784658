set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/i_bench.log 2>&1
ONMF_B200_FUSED_TC=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/i_bench_presplit.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --timeline > gpurun_out/i_bench_timeline.log 2>&1
for w in cfg1 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline > gpurun_out/i_wl_$w.log 2>&1; done
tail -12 gpurun_out/i_pytest.log; for f in gpurun_out/i_bench*.log gpurun_out/i_wl*.log; do tail -1 $f | cut -c1-400; done
