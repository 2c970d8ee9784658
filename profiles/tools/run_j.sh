set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "update_dict or train_matches or fused_step or graph_replayed or two_ranks_match" > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cov_fused|sur_fused|sur_reduce|bcd_kernel" -s 12 -c 4 -o gpurun_out/r2_gemm_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lars_kernel -s 16 -c 1 -o gpurun_out/r2_lars python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu_lars.log 2>&1
tail -5 gpurun_out/j_pytest.log; ls -la gpurun_out/*.ncu-rep
