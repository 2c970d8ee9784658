"""Multi-rank version of prof_e2e.py: where does the host-minibatch loop (uint8 storage, cfg5 column shards) lose time against
device-resident steps when N > 1?   torchrun --nproc-per-node N profiles/tools/prof_e2e_mr.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from onmf_ontf_ndl_b200 import OnmfEngine
from onmf_ontf_ndl_b200.parallel import init_from_env, shard_range

rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
d, k, n_global = 1024, 256, 262144
lo, hi = shard_range(n_global, world, rank)
n = hi - lo
g = torch.Generator(device=dev); g.manual_seed(rank)
gw = torch.Generator(device=dev); gw.manual_seed(0)
W = torch.rand(d, k, device=dev, generator=gw)
p8 = torch.randint(0, 256, (n, d), dtype=torch.uint8, device=dev, generator=g)
host8 = [p8.cpu().pin_memory(), p8.flip(0).cpu().pin_memory()]
dev8 = [h.to(dev) for h in host8]
W_host = torch.empty(d, k).pin_memory()
pg = dist.group.WORLD if world > 1 else None
steps = int(os.environ.get("STEPS", "30"))
for mode in ("resident", "host", "host_noW", "host_sync_each", "resident"):
    eng = OnmfEngine(d, k, alpha=1.0, dtype=torch.float32, device=dev, lars_timing=True, process_group=pg, collect_stats=True)
    eng.set_state(W)
    t = 0

    def one(i):
        global t
        t += 1
        if mode == "resident":
            eng.step_pool(dev8[i & 1], None, float(t), n=n, scale=1 / 255.0)
        elif mode == "host":
            eng.step_host(host8[i & 1], float(t), W_host)
        elif mode == "host_noW":
            eng.step_host(host8[i & 1], float(t), None)
        else:
            eng.step_host(host8[i & 1], float(t), W_host)
            torch.cuda.synchronize()
    for i in range(8):
        one(i)
    eng.flush(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.reset_lars_timing()
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.main)
    c0 = time.perf_counter()
    for i in range(steps):
        one(i)
    c1 = time.perf_counter()
    eng.flush(); e1.record(eng.main); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    lm = eng._plan.lars_ms()
    if rank == 0:
        print("N=%d %-15s %.3f ms/step (max over ranks); host enqueue %.3f ms/step; coder ms: mean %.3f min %.3f max %.3f" %
              (world, mode, float(ms.item()), (c1 - c0) * 1e3 / steps, sum(lm) / max(len(lm), 1), min(lm), max(lm)), flush=True)
    del eng
if world > 1:
    dist.destroy_process_group()
