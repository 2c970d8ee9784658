set -x
mkdir -p gpurun_out
N=${NGPU:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bs_bench_n$N.log 2>&1
tail -1 gpurun_out/bs_bench_n$N.log | cut -c1-200
