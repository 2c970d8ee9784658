set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "step_host or narrow or graph or multirank or two_ranks or class_api" > gpurun_out/bt_pytest.log 2>&1; tail -3 gpurun_out/bt_pytest.log
python bench.py --workload cfg1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bt_wl_cfg1.log 2>&1; tail -1 gpurun_out/bt_wl_cfg1.log | cut -c1-900
