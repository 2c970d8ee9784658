set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_tier" > gpurun_out/ag_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ag_pytest.log
tail -12 gpurun_out/ag_pytest.log
for c in 1 2 3; do python bench.py --workload cfg$c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/ag_wl_cfg$c.log 2>&1; tail -1 gpurun_out/ag_wl_cfg$c.log | cut -c1-200; done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/ag_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ag_pytest_all.log
tail -3 gpurun_out/ag_pytest_all.log
