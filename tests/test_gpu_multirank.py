"""N>1 on the GPU: two ranks (gloo process group over CUDA tensors, both on cuda:0 -- NCCL refuses two ranks on
one device) shard a minibatch by columns, all-reduce the packed partial sums inside OnmfEngine.step, and must
end with the same W, A, B as a single rank that saw the whole minibatch (up to fp32 summation order), and with
bit-identical state on both ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _run(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from onmf_ontf_ndl_b200 import OnmfEngine
    from onmf_ontf_ndl_b200.parallel import shard_range
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    d, k, n = 64, 32, 1001                       # TC-eligible shape, ragged shard sizes
    X = torch.from_numpy(rng.random((n, d)).astype(np.float32)).to(dev)
    W0 = rng.random((d, k))
    res = {}
    for dt in (torch.float32, torch.float64):
        eng = OnmfEngine(d, k, alpha=0.5, dtype=dt, device=dev, process_group=dist.group.WORLD)
        eng.set_state(W0)
        lo, hi = shard_range(n, world, rank)
        Xs = X[lo:hi].to(dt).contiguous()
        for t in (1, 2, 3, 4):
            eng.step(Xs, float(t))
        W, A, B, _ = eng.state()
        torch.cuda.synchronize()
        gathered = [torch.zeros_like(W.cpu()) for _ in range(world)]
        dist.all_gather(gathered, W.cpu())
        same = all(torch.equal(gathered[0], g) for g in gathered)
        if rank == 0:
            ref = OnmfEngine(d, k, alpha=0.5, dtype=dt, device=dev)
            ref.set_state(W0)
            Xf = X.to(dt).contiguous()
            for t in (1, 2, 3, 4):
                ref.step(Xf, float(t))
            Wr, Ar, Br, _ = ref.state()
            torch.cuda.synchronize()
            tol = 1e-11 if dt == torch.float64 else 2e-4
            errs = [float((a - b).abs().max()) / float(b.abs().max()) for a, b in ((W, Wr), (A, Ar), (B, Br))]
            ok = all(e <= tol for e in errs)
            out["detail_%s" % dt] = (errs, same)
            res[str(dt)] = bool(ok and same)
        else:
            res[str(dt)] = bool(same)
    out[rank] = all(res.values())
    dist.destroy_process_group()


def test_two_ranks_match_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_run, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1], dict(out)


# ---------------------------------------------------------------------------------------------- through the reference-facing classes
def test_class_api_two_ranks_gloo_one_device(golden_dir):
    """Online_NTF.train_dict_single / Online_NMF.train_dict with torch.distributed initialised (two gloo ranks sharing
    cuda:0): rank 0's draws are broadcast, minibatches shard by columns, every rank returns the golden single-process
    result and bit-identical dictionaries."""
    import _mr_class_worker as wk
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(wk._spawn_entry, args=(2, port, golden_dir, out), nprocs=2, join=True)
    for r in (0, 1):
        assert all(v["ok"] for v in out[r].values()), dict(out)


def test_class_api_two_ranks_nccl_two_devices(golden_dir, tmp_path):
    """the same under torchrun with NCCL on two devices (skipped on a one-GPU box)."""
    import subprocess
    import sys
    import json
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "_mr_class_worker.py"), golden_dir, str(tmp_path)]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for rank in (0, 1):
        res = json.load(open(os.path.join(str(tmp_path), "rank%d.json" % rank)))
        assert all(v["ok"] for v in res.values()), res
