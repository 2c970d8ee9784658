set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sklearn_order or pgd or reconstruction" > gpurun_out/bf_pytest.log 2>&1; tail -3 gpurun_out/bf_pytest.log
python bench.py --workload next --only gather_patches,reconstruct_image > gpurun_out/bf_next.log 2>&1
cut -c1-400 gpurun_out/bf_next.log
