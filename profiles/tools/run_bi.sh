set -x
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:gather_patches_tile -s 5 -c 1 -f -o gpurun_out/r2n_gather_gray python bench.py --workload next --only gather_patches > /dev/null 2>&1
$N -k regex:gather_patches_tile -s 44 -c 1 -f -o gpurun_out/r2n_gather_32 python bench.py --workload next --only gather_patches > /dev/null 2>&1
$N -k regex:transpose64 -s 5 -c 1 -f -o gpurun_out/r2n_transpose python bench.py --workload next --only transpose > /dev/null 2>&1
$N -k regex:gather_rows_kernel -s 5 -c 1 -f -o gpurun_out/r2n_gather_rows python bench.py --workload next --only gather_rows > /dev/null 2>&1
$N -k regex:pgd_columns_tps -s 5 -c 1 -f -o gpurun_out/r2n_pgd_tps python bench.py --workload next --only pgd > /dev/null 2>&1
$N -k regex:bcd_small -s 31 -c 1 -f -o gpurun_out/r2n_bcd_small python bench.py --workload next --only bcd > /dev/null 2>&1
$N -k regex:bcd_kernel -s 5 -c 1 -f -o gpurun_out/r2n_bcd_cluster python bench.py --workload next --only bcd > /dev/null 2>&1
$N -k regex:motif_patches -s 5 -c 1 -f -o gpurun_out/r2n_motif python bench.py --workload next --only motif_patches > /dev/null 2>&1
ls -la gpurun_out/r2n_*
