"""Worker of tests/test_gpu_multirank.py: multi-GPU training THROUGH the reference-facing classes.

Every rank builds Online_NTF / Online_NMF on the same data with a DIFFERENT numpy seed except rank 0, which uses the
golden run's seed: the classes broadcast rank 0's W0 and minibatch indices, shard every minibatch by columns, all-reduce
the packed partial sums, and must return the reference's single-process result on every rank.
Launched either by torchrun (NCCL, one device per rank) or by mp.spawn (gloo, all ranks on cuda:0)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def per_atom(W, Wref):
    return float(np.max(np.linalg.norm(W - Wref, axis=0) / np.maximum(np.linalg.norm(Wref, axis=0), 1e-30)))


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run(rank, world, backend, golden_dir):
    import torch
    import torch.distributed as dist
    from onmf_ontf_ndl_b200 import Online_NMF, Online_NTF
    g = np.load(os.path.join(golden_dir, "cfg1_renoir_gray.npz"))
    ns, k, batch = int(g["n_steps"]), g["W0"].shape[1], g["idx"].shape[1]
    res = {}
    for prec, tol in (("fp64", 1e-8), ("fp32", 2e-4)):
        np.random.seed(11 if rank == 0 else 1000 + rank)            # only rank 0 holds the golden seed
        m = Online_NTF(g["X"][:, :, None], n_components=k, iterations=ns + 1, batch_size=batch, alpha=1.0, mode=0,
                       learn_joint_dict=False, precision=prec)      # process_group: picked up from torch.distributed
        W, A, B, _ = m.train_dict_single()
        errs = (per_atom(W, g["W_final"]), rel(A, g["A_final"]), rel(B, g["B_final"]))
        Wt = torch.from_numpy(W).cuda()
        gathered = [torch.empty_like(Wt) for _ in range(world)]
        dist.all_gather(gathered, Wt)
        same = all(torch.equal(gathered[0], t) for t in gathered)   # bit-identical dictionaries on every rank
        res[prec] = dict(errs=errs, ok=bool(max(errs) < tol and same and float(m.history) == float(g["history_out"])), same=same)
    # Online_NMF (driver-style 5-tuple incl. the all-reduced d x d aggregate C), all columns per step (subsample=False)
    X = g["X"][:, :301]                                             # ragged shards
    np.random.seed(5 if rank == 0 else 77)
    mm = Online_NMF(X, n_components=25, iterations=4, batch_size=100, alpha=1, subsample=False, precision="fp64")
    W, At, Bt, Ct, H = mm.train_dict()
    from oracle import c_oracle
    rs = np.random.RandomState(5)
    Wr, Ar, Br, Cr = rs.rand(100, 25), np.zeros((25, 25)), np.zeros((25, 100)), np.zeros((100, 100))
    for i in (1, 2, 3):
        Hh, A1, B1, W1 = c_oracle.step(X, Ar, Br, Wr, float(i), 1.0)
        Cr = (1 - 1.0 / i) * Cr + (1.0 / i) * X @ X.T
        Wr, Ar, Br = W1, A1, B1
    e = (per_atom(W, Wr), rel(At, Ar), rel(Bt, Br), rel(Ct, Cr))
    res["nmf"] = dict(errs=e, ok=bool(max(e) < 1e-8))
    return res


def _spawn_entry(rank, world, port, golden_dir, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    out[rank] = run(rank, world, "gloo", golden_dir)
    dist.destroy_process_group()


if __name__ == "__main__":          # torchrun entry: NCCL, one device per rank
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    out = run(rank, world, "nccl", sys.argv[1])
    with open(os.path.join(sys.argv[2], "rank%d.json" % rank), "w") as f:
        json.dump(out, f)
    dist.destroy_process_group()
