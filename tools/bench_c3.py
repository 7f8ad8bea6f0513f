"""GPU: BASELINE config 3 (C2 interactions + item features [I | 1000 tags], WARP, d=128) on the
feature path, with and without the per-CTA hot-row aggregation.  One JSON line per variant.
Usage: python tools/bench_c3.py [nnz] [variants: hot,nohot] [epochs]"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from lightfm_b200 import _lightfm_fast as fast  # noqa: E402
from lightfm_b200 import synthetic  # noqa: E402

nnz = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
variants = (sys.argv[2] if len(sys.argv) > 2 else "hot,nohot").split(",")
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N_USERS, N_ITEMS, D = 138_493, 26_744, 128
fast.set_mode("hogwild")
rows, cols = B.gen_interactions(N_USERS, N_ITEMS, nnz, seed=3, device="cuda")
data = np.ones(len(rows), np.float32)
itf = synthetic.tag_features(N_ITEMS, 1000, 8, seed=3)
usf = sp.identity(N_USERS, dtype=np.float32, format="csr")
pos = sp.csr_matrix((data, (rows, cols)), shape=(N_USERS, N_ITEMS))
pos.sort_indices()
f_item = itf.nnz / itf.shape[0]
peak = 6567.1
try:
    peak = float(json.load(open(os.path.join(B.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

for v in variants:
    fast.set_hot_rows(v == "hot")
    rs = np.random.RandomState(0)
    st = []
    for n in (itf.shape[1], N_USERS):
        emb = ((rs.rand(n, D) - 0.5) / D).astype(np.float32)
        st += [emb, np.ones_like(emb), np.zeros_like(emb), np.zeros(n, np.float32),
               np.ones(n, np.float32), np.zeros(n, np.float32)]
    holder = fast.FastLightFM(*st, D, 0, 0.05, 0.95, 1e-6, 10)
    plan = fast.ResidentPlan("warp", fast.CSRMatrix(itf), fast.CSRMatrix(usf), fast.CSRMatrix(pos), rows, cols,
                             data, data, holder, 0.0, 0.0)
    plan.epoch(seed=1, num_threads=8)
    cs = [plan.epoch(seed=10 + i, num_threads=8) for i in range(epochs)]
    ok = plan.all_finite()
    plan.close()
    ms = sum(c["train_kernel_ms"] for c in cs) / epochs
    P = sum(c["positives"] for c in cs) / epochs
    S = sum(c["negatives_drawn"] for c in cs) / epochs
    U = sum(c["updates"] for c in cs) / epochs
    R = 4 * D + 4
    gather = lambda f: 8 + f * (8 + R)
    abytes = P * (20 + gather(1) + gather(f_item)) + S * gather(f_item) + U * 40 + U * (1 + 2 * f_item) * 3 * R
    print(json.dumps({"config": "C3", "variant": v, "nnz": int(P), "d": D, "features_per_item": round(f_item, 2),
                      "sgd_kernel_ms": round(ms, 2), "M_interactions_per_s": round(P / ms / 1e3, 1),
                      "S": round(S / P, 3), "U": round(U / P, 3),
                      "algorithmic_KB_per_interaction": round(abytes / P / 1e3, 2),
                      "algorithmic_GBps": round(abytes / ms / 1e6, 1), "frac_of_hbm_peak": round(abytes / ms / 1e6 / peak, 3),
                      "finite": ok}), flush=True)
fast.set_hot_rows(True)
