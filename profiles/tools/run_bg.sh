set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/bg_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bg_pytest.log
python bench.py --workload next --only gather_patches,reconstruct_image,bcd > gpurun_out/bg_next.log 2>&1
for c in 1 2 3 4; do
  python bench.py --workload cfg$c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bg_wl_cfg$c.log 2>&1
  ONMF_BCD_SMALL=0 python bench.py --workload cfg$c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bg_wl_cfg${c}_off.log 2>&1
done
tail -4 gpurun_out/bg_pytest.log; cut -c1-330 gpurun_out/bg_next.log
for c in 1 2 3 4; do tail -1 gpurun_out/bg_wl_cfg$c.log | cut -c1-200; tail -1 gpurun_out/bg_wl_cfg${c}_off.log | cut -c1-200; done
