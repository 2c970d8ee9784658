set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/br_bench_n8.log 2>&1
tail -1 gpurun_out/br_bench_n8.log | cut -c1-1500
