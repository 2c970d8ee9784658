set -x
mkdir -p gpurun_out
for c in 1 2 3; do python bench.py --workload cfg$c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/ae_wl_cfg$c.log 2>&1; tail -1 gpurun_out/ae_wl_cfg$c.log | cut -c1-200; done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ae_pytest.log
tail -3 gpurun_out/ae_pytest.log
