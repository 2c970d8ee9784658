set -x
mkdir -p gpurun_out
DIAG_COLS=0 python profiles/tools/diag_cfg5_fp32.py > gpurun_out/h_diag.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "fused_tensor_core" > gpurun_out/h_fused.log 2>&1; echo "rc=$?" >> gpurun_out/h_fused.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_fused_tensor_core_products_vs_fp64 > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
python profiles/tools/prof_lars.py 1024 256 262144 5 > gpurun_out/h_prof.log 2>&1
grep -v Warn gpurun_out/h_diag.log | head -11; tail -15 gpurun_out/h_fused.log; tail -6 gpurun_out/h_pytest.log; grep "lars ms" gpurun_out/h_prof.log
