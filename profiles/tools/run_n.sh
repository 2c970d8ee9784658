set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n_topo.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n_bench_n8.log 2>&1
ONMF_RESERVE_SMS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n_bench_n8_rsv8.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/n_bench_n4.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_n1.log 2>&1
for f in gpurun_out/n_bench_*.log; do tail -1 $f | cut -c1-260; done
