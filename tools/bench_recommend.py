"""GPU: top-k recommend at config-5 shape (1M x 100k, d=32) on a user slice; kernel time of the whole
call (lfm_last_scoring_ms).  Under `ncu --metrics gpu__time_duration.sum` it gives the per-kernel split."""
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
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(0)
st = []
for n in (n_items, n_users):
    st += [rng.normal(size=(n, d)).astype(np.float32) * 0.1, np.ones((n, d), np.float32), np.zeros((n, d), np.float32),
           rng.normal(size=n).astype(np.float32) * 0.1, np.ones(n, np.float32), np.zeros(n, np.float32)]
holder = fast.FastLightFM(*st, d, 0, 0.05, 0.95, 1e-6, 10)
rows = np.repeat(np.arange(slice_users), 100)
train = sp.csr_matrix((np.ones(rows.size, np.float32), (rows, rng.integers(0, n_items, rows.size))), shape=(n_users, n_items))
train.sum_duplicates()
train.sort_indices()
ci = fast.CSRMatrix(sp.identity(n_items, dtype=np.float32, format="csr"))
cu = fast.CSRMatrix(sp.identity(n_users, dtype=np.float32, format="csr"))
ctr = fast.CSRMatrix(train)
users = np.arange(slice_users, dtype=np.int32)
for rep in range(2):
    t0 = time.perf_counter()
    items, sc = fast.recommend(ci, cu, ctr, users, n_items, k, holder)
    wall = time.perf_counter() - t0
print(json.dumps({"users": slice_users, "k": k, "kernel_ms": round(fast.last_scoring_ms(), 3), "call_wall_ms": round(1e3 * wall, 2),
                  "users_per_s_kernel": round(slice_users / fast.last_scoring_ms() * 1e3)}))
