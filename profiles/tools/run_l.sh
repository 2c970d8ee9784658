set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bcd_kernel -c 6 --csv --log-file gpurun_out/l_bcd.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/l_bcd.log 2>&1
for w in cfg1 cfg4; do timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline > gpurun_out/l_wl_$w.log 2>&1; done
tail -6 gpurun_out/l_pytest.log; grep bcd_kernel gpurun_out/l_bcd.csv | awk -F'","' '{print $NF}' | tail -3; for f in gpurun_out/l_wl*.log; do tail -1 $f | cut -c1-200; done
