set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "patch_pipeline or network or update_dict or step_host or narrow" > gpurun_out/bj_pytest.log 2>&1; tail -3 gpurun_out/bj_pytest.log
python bench.py --workload next --only motif_patches,bcd > gpurun_out/bj_next.log 2>&1
cut -c1-300 gpurun_out/bj_next.log
ncu --set full --clock-control none --import-source on -k regex:bcd_small -s 18 -c 1 -f -o gpurun_out/r2n_bcd_small python bench.py --workload next --only bcd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:motif_patches_tile -s 5 -c 1 -f -o gpurun_out/r2n_motif_tile python bench.py --workload next --only motif_patches > /dev/null 2>&1
