set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/t_bench.log 2>&1
python profiles/tools/cmp_fast.py 1024 256 262144 3 > gpurun_out/t_cmp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lars_fast_kernel -s 6 -c 1 -o gpurun_out/r2_lars_fast_final -f python profiles/tools/prof_lars.py 1024 256 262144 2 > gpurun_out/t_ncu.log 2>&1
tail -4 gpurun_out/t_pytest.log; tail -1 gpurun_out/t_bench.log | cut -c1-300; cat gpurun_out/t_cmp.log
