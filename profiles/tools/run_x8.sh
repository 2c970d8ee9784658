set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/x_bench_n8.log 2>&1
ONMF_RESERVE_SMS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/x_bench_n8_rsv8.log 2>&1
tail -1 gpurun_out/x_bench_n8.log | cut -c1-200; tail -1 gpurun_out/x_bench_n8_rsv8.log | cut -c1-200
