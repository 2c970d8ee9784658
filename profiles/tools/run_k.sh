set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "update_dict or train_matches or fused_step or graph_replayed or cfg5 or full_length" > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bcd_kernel -c 6 --csv --log-file gpurun_out/k_bcd.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/k_bcd.log 2>&1
for w in cfg1 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline > gpurun_out/k_wl_$w.log 2>&1; done
timeout 600 python bench.py > gpurun_out/k_bench.log 2>&1
tail -4 gpurun_out/k_pytest.log; grep bcd_kernel gpurun_out/k_bcd.csv | cut -d, -f5,12- | tail -3; for f in gpurun_out/k_wl*.log gpurun_out/k_bench.log; do tail -1 $f | cut -c1-200; done
