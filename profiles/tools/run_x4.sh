set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/x_bench_n4.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/x_bench_n2.log 2>&1
tail -1 gpurun_out/x_bench_n4.log | cut -c1-200; tail -1 gpurun_out/x_bench_n2.log | cut -c1-200
