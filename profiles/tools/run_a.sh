set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py > gpurun_out/a_bench.log 2>&1
python profiles/tools/prof_lars.py 1024 256 262144 5 > gpurun_out/a_var_base.log 2>&1
for v in variants_tmp/*.so; do
  ONMF_B200_LIB=$PWD/$v timeout 120 python profiles/tools/prof_lars.py 1024 256 262144 5 > gpurun_out/a_var_$(basename $v .so).log 2>&1
done
for w in cfg1 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --steps 50 --no-cpu-baseline > gpurun_out/a_wl_$w.log 2>&1; done
tail -3 gpurun_out/a_pytest.log; tail -1 gpurun_out/a_bench.log | cut -c1-600; cat gpurun_out/a_var_*.log | grep "lars ms"
