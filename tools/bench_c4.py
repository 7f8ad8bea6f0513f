"""BASELINE config 4 on N GPUs (torchrun): synthetic implicit 10M users x 1M items, 500M nnz,
identity features, WARP-kOS (k=5, n=10), d=64, item-sharded.

Each rank generates ITS shard directly (1M/N local items, 500M/N interactions over all 10M users;
a hash shard of a synthetic matrix is itself a synthetic matrix of that shape), keeps the full user
table replicated and all-reduces its epoch delta with NCCL.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29513 tools/bench_c4.py [epochs]
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from lightfm_b200 import _lightfm_fast as fast  # noqa: E402
from lightfm_b200 import sharding  # noqa: E402

_SCALE = int(os.environ.get("C4_SCALE", "1"))  # debug: shrink every dimension
N_USERS, N_ITEMS, NNZ, D = 10_000_000 // _SCALE, 1_000_000 // _SCALE, 500_000_000 // _SCALE, 64
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
fast._lib.lfm_set_device(local)
fast.set_mode("hogwild")

n_local_items, nnz_local = N_ITEMS // world, NNZ // world
t0 = time.time()
rows, cols = B.gen_interactions(N_USERS, n_local_items, nnz_local, seed=40 + rank, device=device)
pos = sp.csr_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(N_USERS, n_local_items))
pos.sort_indices()
torch.cuda.empty_cache()
st = []
for n, seed in ((n_local_items, 7 + rank), (N_USERS, 12345)):
    rs = np.random.RandomState(seed)
    emb = ((rs.rand(n, D).astype(np.float32) - 0.5) / D).astype(np.float32)
    st += [emb, np.ones_like(emb), np.zeros((n, D), np.float32), np.zeros(n, np.float32),
           np.ones(n, np.float32), np.zeros(n, np.float32)]
holder = fast.FastLightFM(*st, D, 0, 0.05, 0.95, 1e-6, 10)
plan = fast.ResidentPlan("warp-kos", fast.CSRMatrix(sp.identity(n_local_items, dtype=np.float32, format="csr")),
                         fast.CSRMatrix(sp.identity(N_USERS, dtype=np.float32, format="csr")), fast.CSRMatrix(pos),
                         rows, None, None, None, holder, 0.0, 0.0, 5, 10)
plan.set_global_items(N_ITEMS)
views = [torch.as_tensor(sharding.CudaArrayView(*plan.table(w)), device=device) for w in (6, 7, 9, 10)]
prep = time.time() - t0


def step(seed):
    snaps = [v.clone() for v in views]
    torch.cuda.synchronize(device)
    c = plan.epoch(seed=seed * 977 + rank, num_threads=8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world > 1:
        sharding.allreduce_deltas(views, snaps)
    e1.record()
    torch.cuda.synchronize(device)
    c["allreduce_ms"] = e0.elapsed_time(e1)
    return c


step(1)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(device)
t0 = time.perf_counter()
cs = [step(10 + e) for e in range(epochs)]
torch.cuda.synchronize(device)
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
dev_ms = sum(c["kernel_ms"] + c["allreduce_ms"] for c in cs)
stats = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=device)
sums = torch.tensor([sum(c["positives"] for c in cs), sum(c["negatives_drawn"] for c in cs),
                     sum(c["updates"] for c in cs), sum(c["train_kernel_ms"] for c in cs),
                     sum(c["allreduce_ms"] for c in cs)], dtype=torch.float64, device=device)
if world > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
ok = plan.all_finite()
if rank == 0:
    P, S, U = sums[0].item(), sums[1].item(), sums[2].item()
    R = 4 * D + 4
    # SURVEY 8(d) for k-OS, identity features: user gather + min(n, nnz_u) positive gathers + the
    # chosen positive again + S negative gathers + update
    abytes = P * (20 + (16 + R) + 10 * (16 + R) + (16 + R)) + S * (16 + R) + U * 40 + U * 9 * R
    train_ms = sums[3].item() / world
    print(json.dumps({"config": "C4", "n_gpus": world, "epochs": epochs,
                      "interactions_per_s": P / (stats[0].item() / 1e3),
                      "ms_per_epoch": stats[0].item() / epochs, "wall_ms_per_epoch": stats[1].item() / epochs,
                      "sgd_kernel_ms_per_epoch": train_ms / epochs, "allreduce_ms_per_epoch": sums[4].item() / world / epochs,
                      "S_per_positive": S / P, "U_per_positive": U / P,
                      "per_gpu_algorithmic_GBps": abytes / world / (train_ms / 1e3) / 1e9,
                      "per_gpu_frac_of_measured_hbm_peak": abytes / world / (train_ms / 1e3) / 1e9 / 6567.1,
                      "finite": ok, "prep_s": round(prep, 1)}))
plan.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
