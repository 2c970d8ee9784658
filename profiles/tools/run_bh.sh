set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "step_host or narrow or multirank or two_ranks or class_api or graph or update_dict" > gpurun_out/bh_pytest.log 2>&1; tail -3 gpurun_out/bh_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 profiles/tools/prof_e2e_mr.py > gpurun_out/bh_e2e_n2.log 2>&1
grep "^N=" gpurun_out/bh_e2e_n2.log || tail -20 gpurun_out/bh_e2e_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bh_bench_n2.log 2>&1
tail -1 gpurun_out/bh_bench_n2.log | cut -c1-1200
