set -x
mkdir -p gpurun_out
timeout 60 python profiles/tools/prof_bcd.py 1024 256 50 > gpurun_out/bq_bcd.log 2>&1; cat gpurun_out/bq_bcd.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/bq_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/bq_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bq_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bq_smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bq_ref.log 2>&1
tail -4 gpurun_out/bq_pytest.log; tail -1 gpurun_out/bq_bench.log | cut -c1-300; tail -2 gpurun_out/bq_smoke.log; tail -1 gpurun_out/bq_ref.log | cut -c1-300
