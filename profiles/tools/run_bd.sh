set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vector_gathers or test_gathers or reconstruction or patch or gather_index" > gpurun_out/bd_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bd_pytest.log
python bench.py --workload next --only gather_patches > gpurun_out/bd_next.log 2>&1
tail -5 gpurun_out/bd_pytest.log; cut -c1-250 gpurun_out/bd_next.log
