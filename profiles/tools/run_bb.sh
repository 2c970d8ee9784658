set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vector_gathers or thread_per_sample or test_gathers or pgd or reconstruction or patch" > gpurun_out/bb_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bb_pytest.log
python bench.py --workload next --only gather_patches,transpose,pgd,reconstruct_image > gpurun_out/bb_next.log 2>&1
tail -5 gpurun_out/bb_pytest.log; cut -c1-330 gpurun_out/bb_next.log
