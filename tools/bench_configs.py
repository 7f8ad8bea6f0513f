"""GPU: one timing line per BASELINE.json config (C1..C5) -- coverage evidence for DESIGN.md.
Not the headline bench (bench.py); same resident-plan timing method (CUDA events, 2 warm-up +
3 timed epochs).  Usage: python tools/bench_configs.py [C1,C3,C4s,C5]"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from lightfm_b200 import _lightfm_fast as fast  # noqa: E402
from lightfm_b200 import synthetic  # noqa: E402

which = (sys.argv[1] if len(sys.argv) > 1 else "C1,C2,C3,C4s,C5").split(",")
fast.set_mode("hogwild")


def state(n_items, n_users, d, seed=0):
    rs = np.random.RandomState(seed)
    out = []
    for n in (n_items, n_users):
        emb = ((rs.rand(n, d) - 0.5) / d).astype(np.float32)
        out += [emb, np.ones_like(emb), np.zeros_like(emb), np.zeros(n, np.float32),
                np.ones(n, np.float32), np.zeros(n, np.float32)]
    return out


def run(name, loss, n_users, n_items, nnz, d, itf=None, signed=False, weights=False, k=5, n=10, note=""):
    t0 = time.time()
    rows, cols = B.gen_interactions(n_users, n_items, nnz, seed=abs(hash(name)) % 1000, device="cuda")
    rng = np.random.default_rng(1)
    data = np.where(rng.random(len(rows)) < 0.5, 1.0, -1.0).astype(np.float32) if signed \
        else np.ones(len(rows), np.float32)
    w = (0.5 + rng.random(len(rows))).astype(np.float32) if weights else data
    itf = itf if itf is not None else sp.identity(n_items, dtype=np.float32, format="csr")
    usf = sp.identity(n_users, dtype=np.float32, format="csr")
    pos = None
    if loss != "logistic":
        pos = sp.csr_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(n_users, n_items))
        pos.sort_indices()
    st = state(itf.shape[1], n_users, d)
    holder = fast.FastLightFM(*st, d, 0, 0.05, 0.95, 1e-6, 10)
    kos = loss == "warp-kos"
    plan = fast.ResidentPlan(loss, fast.CSRMatrix(itf), fast.CSRMatrix(usf),
                             fast.CSRMatrix(pos) if pos is not None else None, rows,
                             None if kos else cols, None if kos else data, None if kos else w,
                             holder, 0.0, 0.0, k, n)
    prep = time.time() - t0
    for i in range(2):
        plan.epoch(seed=10 + i, num_threads=8)
    cs = [plan.epoch(seed=20 + i, num_threads=8) for i in range(3)]
    ok = plan.all_finite()
    plan.close()
    ms = sum(c["train_kernel_ms"] for c in cs) / 3
    allms = sum(c["kernel_ms"] for c in cs) / 3
    c = cs[-1]
    print(json.dumps({"config": name, "loss": loss, "shape": [n_users, n_items], "nnz": len(rows), "d": d,
                      "train_kernel_ms": round(ms, 3), "epoch_device_ms": round(allms, 3),
                      "M_interactions_per_s": round(c["positives"] / allms / 1e3, 1),
                      "S": round(c["negatives_drawn"] / max(1, c["positives"]), 3),
                      "U": round(c["updates"] / max(1, c["positives"]), 3), "finite": ok,
                      "prep_s": round(prep, 1), "note": note}), flush=True)


if "C1" in which:
    run("C1", "bpr", 943, 1682, 100_000, 16, note="ML-100k shape; hogwild fast_pair_kernel; in-flight cap 781")
if "C2" in which:
    run("C2", "warp", 138_493, 26_744, 20_000_000, 64, note="headline (see bench.py)")
if "C3" in which:
    itf = synthetic.tag_features(26_744, 1000, 8, seed=3)
    run("C3", "warp", 138_493, 26_744, 20_000_000, 128, itf=itf,
        note="item features = [I | 1000 tags], 8 tags/item, L1-normalised; generic hogwild kernel")
if "C4s" in which:
    run("C4s", "warp-kos", 10_000_000, 125_000, 62_500_000, 64, k=5, n=10,
        note="one GPU's share of C4 (1/8 of the items and interactions, all 10M users)")
if "C5" in which:
    run("C5", "logistic", 1_000_000, 100_000, 100_000_000, 32, signed=True, weights=True,
        note="explicit +-1 with sample_weight ~ U(0.5,1.5)")
