set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/bk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bk_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bk_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bk_smoke.log 2>&1
python bench.py --workload next > gpurun_out/bk_next.log 2>&1
for c in 1 2 3 4; do python bench.py --workload cfg$c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bk_wl_cfg$c.log 2>&1; done
tail -4 gpurun_out/bk_pytest.log; tail -1 gpurun_out/bk_bench.log | cut -c1-300; tail -2 gpurun_out/bk_smoke.log
