"""GPU: predict_rank throughput at config-5 shape (1M x 100k, d=32) on a user slice."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightfm_b200 import _lightfm_fast as fast  # noqa: E402

n_users, n_items, d = 1_000_000, 100_000, 32
slice_users = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
rng = np.random.default_rng(0)
st = []
for n in (n_items, n_users):
    st += [rng.normal(size=(n, d)).astype(np.float32) * 0.1, np.ones((n, d), np.float32), np.zeros((n, d), np.float32),
           rng.normal(size=n).astype(np.float32) * 0.1, np.ones(n, np.float32), np.zeros(n, np.float32)]
holder = fast.FastLightFM(*st, d, 0, 0.05, 0.95, 1e-6, 10)
per_user_test, per_user_train = 10, 100
rows = np.repeat(np.arange(slice_users), per_user_test)
test = sp.csr_matrix((np.ones(rows.size, np.float32), (rows, rng.integers(0, n_items, rows.size))), shape=(n_users, n_items))
rows = np.repeat(np.arange(slice_users), per_user_train)
train = sp.csr_matrix((np.ones(rows.size, np.float32), (rows, rng.integers(0, n_items, rows.size))), shape=(n_users, n_items))
for m in (test, train):
    m.sum_duplicates()
    m.sort_indices()
itf = sp.identity(n_items, dtype=np.float32, format="csr")
usf = sp.identity(n_users, dtype=np.float32, format="csr")
ci, cu, ct, ctr = fast.CSRMatrix(itf), fast.CSRMatrix(usf), fast.CSRMatrix(test), fast.CSRMatrix(train)
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fast.set_rank_groups(groups)
for rep in range(2):
    ranks = np.zeros_like(test.data)
    t0 = time.perf_counter()
    fast.predict_ranks(ci, cu, ct, ctr, ranks, holder, 8)
    dt = time.perf_counter() - t0
scores = slice_users * n_items
kms = fast.last_scoring_ms()
print(json.dumps({"users": slice_users, "groups_per_cta": groups, "kernel_ms": round(kms, 3),
                  "G_scores_per_s_kernel": round(scores / kms / 1e6, 1),
                  "TFLOPs_kernel": round(scores * (2 * d + 1) / kms / 1e9, 2), "items": n_items, "d": d, "test_nnz": int(test.nnz), "wall_s": round(dt, 3),
                  "G_user_item_scores_per_s": round(scores / dt / 1e9, 2),
                  "full_C5_estimate_s": round(dt * n_users / slice_users, 1), "rank_mean": float(ranks.mean())}))
