"""bench.py's N > 1 leg: one process per GPU (torchrun), weak scaling of the headline workload.

Global problem at N ranks: 138 493 users x (26 744 * N) items, 20 M * N interactions, WARP d=64.
Items are hash-sharded (lightfm_b200.sharding): every rank owns ~26 744 item rows and trains the
~20 M interactions whose positive item it owns, drawing negatives from its own shard while the
WARP rank estimate keeps the global catalogue size; the user table is replicated and its epoch
delta is all-reduced with NCCL once per epoch.  Per-GPU work is therefore the N=1
workload plus one ~70 MB all-reduce -- "scaling": "weak".
"""
import json
import os
import time

import numpy as np
import scipy.sparse as sp

import bench as B

_MUL = 0x9E3779B97F4A7C15 - (1 << 64)  # the multiplier of sharding.shard_of as a signed int64


def torch_shard_of(ids, world):
    """lightfm_b200.sharding.shard_of on a torch int64 tensor (bit-identical)."""
    h = (ids + 1) * _MUL
    return ((h >> 33) & 0x7FFFFFFF) % world


def local_problem(rank, world, nnz_per_gpu, seed, device):
    """Generate the GLOBAL interaction list on the GPU (same seed on every rank), keep this
    rank's item shard, remap items to local ids."""
    import torch
    from lightfm_b200 import sharding
    n_items_global = B.N_ITEMS * world
    rows, cols = B.gen_interactions(B.N_USERS, n_items_global, nnz_per_gpu * world, seed, device)
    cols_t = torch.from_numpy(cols.astype(np.int64)).to(device)
    mine = (torch_shard_of(cols_t, world) == rank).cpu().numpy()
    smap = sharding.ShardMap(n_items_global, rank, world)
    return rows[mine], smap.local_of[cols[mine]], smap, n_items_global


class LocalProblem(B.Problem):
    def __init__(self, rows, cols, n_users, n_local_items, d, seed):
        self.keep = []
        data = np.ones(len(rows), dtype=np.float32)
        csr = sp.csr_matrix((data, (rows, cols)), shape=(n_users, n_local_items))
        csr.sort_indices()
        self.n_users, self.n_items, self.nnz, self.d = n_users, n_local_items, len(rows), d
        P = self._pin
        self.row, self.col, self.data = P(rows.astype(np.int32)), P(cols.astype(np.int32)), P(data)
        self.pos = sp.csr_matrix((P(csr.data), P(csr.indices.astype(np.int32)),
                                  P(csr.indptr.astype(np.int32))), shape=csr.shape)
        self.itf = sp.identity(n_local_items, dtype=np.float32, format="csr")
        self.usf = sp.identity(n_users, dtype=np.float32, format="csr")
        for m in (self.itf, self.usf):
            m.indices, m.indptr, m.data = P(m.indices.astype(np.int32)), P(m.indptr.astype(np.int32)), P(m.data)
        self.state = {}
        for side, n, s in (("item", n_local_items, seed), ("user", n_users, 12345)):
            rs = np.random.RandomState(s)  # user table: same seed everywhere (replicated block)
            emb = ((rs.rand(n, d) - 0.5) / d).astype(np.float32)
            self.state[side + "_w"] = P(emb)
            self.state[side + "_g"] = P(np.ones_like(emb))
            self.state[side + "_m"] = P(np.zeros_like(emb))
            self.state[side + "_b"] = P(np.zeros(n, np.float32))
            self.state[side + "_bg"] = P(np.ones(n, np.float32))
            self.state[side + "_bm"] = P(np.zeros(n, np.float32))
        self.shuffle = P(np.arange(self.nnz, dtype=np.int32))


USER_TABLES = (6, 7, 9, 10)  # lfm_plan_table indices: user w, g, b, bg


def user_views(plan, device):
    import torch
    from lightfm_b200.sharding import CudaArrayView
    views = []
    for which in USER_TABLES:
        ptr, cnt = plan.table(which)
        views.append(torch.as_tensor(CudaArrayView(ptr, cnt), device=device))
    return views


def run(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from lightfm_b200 import _lightfm_fast as fast
    from lightfm_b200 import sharding
    device = torch.device("cuda", local)
    threads = max(2, os.cpu_count() or 2)
    rows, cols, smap, n_items_global = local_problem(rank, world, args.nnz, seed=2, device=device)
    prob = LocalProblem(rows, cols, B.N_USERS, smap.n_local, B.D, seed=100 + rank)
    itf, usf, pos = fast.CSRMatrix(prob.itf), fast.CSRMatrix(prob.usf), fast.CSRMatrix(prob.pos)
    holder = prob.holder(fast)
    clocks = B.ClockSampler(local)
    clocks.start()

    views = None

    def make_plan():
        p = fast.ResidentPlan("warp", itf, usf, pos, prob.row, prob.col, prob.data, prob.data, holder, 0.0, 0.0)
        p.set_global_items(n_items_global)
        return p

    def step(plan, views, seed):
        # exchange: snapshot sweep, local epoch, delta sweep, ONE all-reduce of the packed user-table
        # delta (w, g, b, bg), add-back sweep -- all timed on the device (CUDA events)
        begin_ms = plan.delta_begin(1)
        c = plan.epoch(seed=seed * 977 + rank, num_threads=threads)
        t = sharding.exchange_replicated(plan, 1, device, None, world)
        c["allreduce_ms"] = t["allreduce_ms"]
        c["exchange_ms"] = begin_ms + t["make_ms"] + t["allreduce_ms"] + t["apply_ms"]
        return c

    # -- value: resident plan; step = local epoch + delta all-reduce ----------------------------
    plan = make_plan()
    for w in range(args.warmup):
        step(plan, views, 1000 + w)
    dist.barrier()
    torch.cuda.synchronize(device)
    clocks.mark(True)
    t0 = time.perf_counter()
    counters = [step(plan, views, 2000 + s) for s in range(args.steps)]
    torch.cuda.synchronize(device)
    dist.barrier()
    wall = time.perf_counter() - t0
    clocks.mark(False)
    clk = clocks.stop()
    dev_ms = sum(c["kernel_ms"] + c["exchange_ms"] for c in counters)
    plan.download()
    plan.close()

    # -- e2e: host buffers in, host buffers out, every step -------------------------------------
    e2e_ms, h2d = [], 0
    p = make_plan()
    v = None
    for s in range(args.warmup + args.steps):
        dist.barrier()
        t0 = time.perf_counter()
        p.refresh()                          # H2D of the local inputs and state (pinned), same buffers
        c = step(p, v, 3000 + s)
        p.download()                         # D2H of the state
        torch.cuda.synchronize(device)
        if s >= args.warmup:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
    p.close()
    h2d = 4 * (3 * prob.nnz + prob.nnz + prob.n_users + 2 * prob.n_items + 2 * prob.n_users) + \
        8 * (prob.n_items + prob.n_users) * (B.D + 1)
    d2h = 8 * (prob.n_items + prob.n_users) * (B.D + 1)

    stats = torch.tensor([dev_ms, sum(e2e_ms), wall * 1e3], dtype=torch.float64, device=device)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    sums = torch.tensor([sum(c["positives"] for c in counters), sum(c["negatives_drawn"] for c in counters),
                         sum(c["updates"] for c in counters), sum(c["train_kernel_ms"] for c in counters),
                         sum(B.algorithmic_bytes(c, B.D) for c in counters),
                         sum(c["kernel_launches"] + 3 for c in counters), sum(c["allreduce_ms"] for c in counters),
                         sum(c["exchange_ms"] for c in counters)],
                        dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    # C4 runs AFTER every collective the headline numbers depend on, under try/except on every rank:
    # a failure there (or a rank that cannot follow: the process-group timeout turns a hang into an
    # exception) costs the c4 block, not the line.
    c4 = None
    if not getattr(args, "no_c4", False):
        try:
            import bench_c4
            pk = 6567.1
            try:
                pk = float(json.load(open(os.path.join(B.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                pass
            c4 = bench_c4.run(fast, sharding, dist, rank, world, device, epochs=2, peak=pk)
        except Exception as exc:  # pragma: no cover
            c4 = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        positives = sums[0].item()
        value = positives / (stats[0].item() / 1e3)
        e2e_value = positives / (stats[1].item() / 1e3)
        peak = 6650.0
        try:
            peak = float(json.load(open(os.path.join(B.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        achieved = (sums[4].item() / world) / ((sums[3].item() / world) / 1e3) / 1e9
        print(json.dumps({
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": stats[0].item() / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": B.base_config(args.nnz),
            "detail": {"per_gpu_workload": "C2 per GPU, item-sharded: 138493 users x %d items, %d nnz, WARP, d=64"
                                           % (n_items_global, int(positives / args.steps)),
                       "parallelism": "item-shard x%d + one NCCL all-reduce of the packed user-table delta per epoch, "
                                      "subtract / add-back fused into lfm_plan_delta_* sweeps" % world,
                       "wall_ms_per_step": stats[2].item() / args.steps,
                       "allreduce_ms_per_step": sums[6].item() / world / args.steps,
                       "exchange_ms_per_step": sums[7].item() / world / args.steps},
            "c4": c4,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": B.UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": stats[1].item() / args.steps,
                    "call": "per rank: lfm_plan_create/refresh(host pinned buffers) + lfm_plan_epoch + NCCL all-reduce + lfm_plan_download"},
            "gpu_launches": int(sums[5].item()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "kernel": fast.warp_kernel_name(B.D),
                         "note": "per-GPU average"},
            "cpu_baseline": None,
        }))
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        pass
