set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/tools/prof_e2e_mr.py > gpurun_out/be_e2e_n2.log 2>&1
grep "^N=" gpurun_out/be_e2e_n2.log || tail -20 gpurun_out/be_e2e_n2.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -x -k "sklearn_order or multirank or two_ranks or class_api" > gpurun_out/be_pytest.log 2>&1; tail -3 gpurun_out/be_pytest.log
