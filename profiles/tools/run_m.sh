set -x
mkdir -p gpurun_out
for t in 0 2 4 8 16; do ONMF_BCD_TPR=$t python profiles/tools/prof_bcd.py 1024 256 >> gpurun_out/m_bcd.log 2>&1; done
for t in 0 2 4; do ONMF_BCD_TPR=$t python profiles/tools/prof_bcd.py 400 100 >> gpurun_out/m_bcd.log 2>&1; done
python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/m_multirank.log 2>&1; echo "rc=$?" >> gpurun_out/m_multirank.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/m_bench_n2.log 2>&1
grep bcd gpurun_out/m_bcd.log; tail -4 gpurun_out/m_multirank.log; tail -1 gpurun_out/m_bench_n2.log | cut -c1-300
