"""bench.py's C4 block: BASELINE config 4 -- synthetic implicit 10 M users x 1 M items, 500 M
interactions, identity features, WARP-kOS (k=5, n=10), d=64 -- STRONG scaling over N GPUs
(the 500 M interactions are fixed, each rank holds 1/N of them).

  N = 1          the whole problem on one GPU (tables 5.6 GB, tuples 8 GB: fits 180 GB), no exchange
  N > 1, item    north-star partition: item rows hash-sharded (each rank: all 10 M users x 1 M / N
                 items), negatives from the local shard with the global catalogue size in the rank
                 estimate, the 10 M-row user table replicated and its epoch delta all-reduced (5.1 GB)
  N > 1, user    the mirror image: users sharded, the 1 M-row item table replicated (0.5 GB
                 all-reduce), negative sampling over the whole catalogue as in the reference

A hash shard of a synthetic matrix is itself a synthetic matrix of the shard's shape, so every rank
generates its own shard directly on its GPU.  Exchange per epoch: lfm_plan_delta_begin / make /
apply (two fused sweeps of our kernels over one packed buffer) around ONE NCCL all-reduce.
Timing: CUDA events, max over ranks.  1 warm-up + `epochs` timed epochs per variant (fixed and
small, independent of --steps, so that the 1 -> 8 GPU driver run stays within minutes).
"""
import json
import os
import time

import numpy as np
import scipy.sparse as sp

N_USERS, N_ITEMS, NNZ, D, K, N_KOS = 10_000_000, 1_000_000, 500_000_000, 64, 5, 10
_SCALE = int(os.environ.get("C4_SCALE", "1"))  # debug: shrink every dimension


def gen_sorted_csr(n_users, n_items, nnz, seed, device):
    """Synthetic interactions (same marginals as bench.gen_interactions) as a CSR built on the
    GPU: sorted unique keys ARE the CSR order, so no host-side sort is needed."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    draw = int(nnz * 1.2) + 16
    while keys.numel() < nnz:
        u = torch.floor(n_users * torch.rand(draw, generator=g, device=device, dtype=torch.float64) ** 1.5).long()
        i = torch.floor(n_items * torch.rand(draw, generator=g, device=device, dtype=torch.float64) ** 2.0).long()
        new = u * n_items + i
        del u, i
        keys = torch.unique(torch.cat([keys, new]))
        del new
        draw = max(int((nnz - keys.numel()) * 2.0) + 16, 1024)
    if keys.numel() > nnz:
        keep = torch.randperm(keys.numel(), generator=g, device=device)[:nnz]
        keys = keys[torch.sort(keep).values]
        del keep
    rows = (keys // n_items).to(torch.int32)
    cols = (keys % n_items).to(torch.int32)
    del keys
    counts = torch.bincount(rows.long(), minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    out = rows.cpu().numpy(), cols.cpu().numpy(), indptr.to(torch.int32).cpu().numpy()
    del rows, cols, counts, indptr
    torch.cuda.empty_cache()
    return out


def table_state(n, d, seed, device):
    """[w, g, m, b, bg, bm] of one side: w drawn on the GPU (numpy takes ~7 s for 10 M x 64)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w = ((torch.rand(n, d, generator=g, device=device, dtype=torch.float32) - 0.5) / d).cpu().numpy()
    return [w, np.ones((n, d), np.float32), np.zeros((n, d), np.float32), np.zeros(n, np.float32),
            np.ones(n, np.float32), np.zeros(n, np.float32)]


def run_variant(fast, sharding, dist, rank, world, device, axis, epochs, peak):
    import torch
    n_users, n_items, nnz = N_USERS // _SCALE, N_ITEMS // _SCALE, NNZ // _SCALE
    lu, li = n_users, n_items
    if world > 1 and axis == "item":
        li = n_items // world
    elif world > 1 and axis == "user":
        lu = n_users // world
    t0 = time.time()
    rows, cols, indptr = gen_sorted_csr(lu, li, nnz // world, seed=400 + 17 * rank + (0 if axis == "item" else 7),
                                        device=device)
    pos = sp.csr_matrix((lu, li), dtype=np.float32)
    pos.indices, pos.indptr, pos.data = cols, indptr, np.ones(len(cols), np.float32)
    # the replicated table starts identical on every rank (same seed); the sharded one is per rank
    item_seed = 12345 if (axis == "user" or world == 1) else 7 + rank
    user_seed = 12345 if (axis == "item" or world == 1) else 9 + rank
    st = table_state(li, D, item_seed, device) + table_state(lu, D, user_seed, device)
    holder = fast.FastLightFM(*st, D, 0, 0.05, 0.95, 1e-6, 10)
    plan = fast.ResidentPlan("warp-kos", fast.CSRMatrix(sp.identity(li, dtype=np.float32, format="csr")),
                             fast.CSRMatrix(sp.identity(lu, dtype=np.float32, format="csr")), fast.CSRMatrix(pos),
                             rows, None, None, None, holder, 0.0, 0.0, K, N_KOS)
    if axis == "item" and world > 1:
        plan.set_global_items(n_items)
    side = 1 if axis == "item" else 0   # the replicated table
    prep = time.time() - t0

    def step(seed):
        begin_ms = plan.delta_begin(side) if world > 1 else 0.0
        c = plan.epoch(seed=seed * 977 + rank, num_threads=8)
        c["allreduce_ms"] = 0.0
        c["exchange_ms"] = 0.0
        if world > 1:
            t = sharding.exchange_replicated(plan, side, device, None, world)
            c["allreduce_ms"] = t["allreduce_ms"]
            c["exchange_ms"] = begin_ms + t["make_ms"] + t["allreduce_ms"] + t["apply_ms"]
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
    ok = plan.all_finite()
    plan.close()
    fast.release_cache()
    del st, holder, plan, rows, cols, indptr, pos
    torch.cuda.empty_cache()
    dev_ms = sum(c["kernel_ms"] + c["exchange_ms"] for c in cs)
    stats = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=device)
    sums = torch.tensor([sum(c["positives"] for c in cs), sum(c["negatives_drawn"] for c in cs),
                         sum(c["updates"] for c in cs), sum(c["train_kernel_ms"] for c in cs),
                         sum(c["allreduce_ms"] for c in cs), sum(c["exchange_ms"] for c in cs),
                         float(ok)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    P, S, U = sums[0].item(), sums[1].item(), sums[2].item()
    R = 4 * D + 4
    # SURVEY 8(d) for k-OS, identity features: user gather + min(n, nnz_u) sampled positives + the
    # chosen positive again + S negative gathers + update
    abytes = P * (20 + (16 + R) + N_KOS * (16 + R) + (16 + R)) + S * (16 + R) + U * 40 + U * 9 * R
    train_ms = sums[3].item() / world
    gbps = abytes / world / (train_ms / 1e3) / 1e9
    return {"axis": axis if world > 1 else "single GPU", "epochs": epochs,
            "interactions_per_s": P / (stats[0].item() / 1e3),
            "ms_per_epoch": stats[0].item() / epochs, "wall_ms_per_epoch": stats[1].item() / epochs,
            "sgd_kernel_ms_per_epoch": train_ms / epochs,
            "allreduce_ms_per_epoch": sums[4].item() / world / epochs,
            "exchange_ms_per_epoch": sums[5].item() / world / epochs,
            "exchange_frac_of_epoch": (sums[5].item() / world) / max(stats[0].item(), 1e-9),
            "allreduce_bytes": int(2 * (lu if axis == "item" else li) * (D + 1) * 4) if world > 1 else 0,
            "S_per_positive": S / max(P, 1), "U_per_positive": U / max(P, 1),
            "per_gpu_algorithmic_GBps": gbps, "per_gpu_frac_of_hbm_peak": gbps / peak,
            "finite": bool(sums[6].item() == world), "prep_s": round(prep, 1)}


def run(fast, sharding, dist, rank, world, device, epochs=2, peak=6567.1):
    """Returns the `c4` block of the bench line (rank 0; None elsewhere)."""
    out = {"workload": "C4: synthetic implicit %d users x %d items, %d interactions, identity features, "
                       "WARP-kOS k=%d n=%d, d=%d; strong scaling (total fixed)"
                       % (N_USERS // _SCALE, N_ITEMS // _SCALE, NNZ // _SCALE, K, N_KOS, D),
           "n_gpus": world, "peak_GBps": peak}
    if world == 1:
        out["single"] = run_variant(fast, sharding, dist, rank, world, device, "item", epochs, peak)
        out["interactions_per_s"] = out["single"]["interactions_per_s"]
    else:
        for axis in ("item", "user"):
            out[axis + "_sharded"] = run_variant(fast, sharding, dist, rank, world, device, axis, epochs, peak)
        best = max(("item_sharded", "user_sharded"), key=lambda k: out[k]["interactions_per_s"])
        out["interactions_per_s"] = out[best]["interactions_per_s"]
        out["interactions_per_s_from"] = best
    return out if rank == 0 else None


if __name__ == "__main__":   # stand-alone: torchrun ... bench_c4.py [epochs]
    import sys
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from lightfm_b200 import _lightfm_fast as fast
    from lightfm_b200 import sharding
    fast.set_device(local)
    fast.set_mode("hogwild")
    res = run(fast, sharding, dist, rank, world, dev, epochs=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
