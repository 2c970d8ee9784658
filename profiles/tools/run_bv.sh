set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bv_bench_n2.log 2>&1
tail -1 gpurun_out/bv_bench_n2.log | cut -c1-1400
