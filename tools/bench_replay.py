"""GPU: throughput of the deterministic replay mode (num_threads=1) through the boundary call."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightfm_b200 import _lightfm_fast as fast  # noqa: E402
from lightfm_b200 import synthetic  # noqa: E402

CONFIGS = (("C1", "bpr", 943, 1682, 100_000, 16), ("C2-slice", "warp", 138_493, 26_744, 400_000, 64),
           ("logistic", "logistic", 20_000, 5_000, 300_000, 32),
           ("C2-shape-bpr", "bpr", 138_493, 26_744, 2_000_000, 64),
           ("C5-slice-logistic", "logistic", 100_000, 100_000, 4_000_000, 32))
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
for name, loss, nu, ni, nnz, d in CONFIGS:
    if only and name not in only:
        continue
    inter = synthetic.interactions(nu, ni, nnz, seed=1, signed=(loss == "logistic"))
    rs = np.random.RandomState(0)
    st = []
    for n in (ni, nu):
        emb = ((rs.rand(n, d) - 0.5) / d).astype(np.float32)
        st += [emb, np.ones_like(emb), np.zeros_like(emb), np.zeros(n, np.float32), np.ones(n, np.float32), np.zeros(n, np.float32)]
    holder = fast.FastLightFM(*st, d, 0, 0.05, 0.95, 1e-6, 10)
    itf = fast.CSRMatrix(sp.identity(ni, dtype=np.float32, format="csr"))
    usf = fast.CSRMatrix(sp.identity(nu, dtype=np.float32, format="csr"))
    pos = inter.tocsr(); pos.sort_indices()
    shuffle = np.arange(inter.nnz, dtype=np.int32); rs.shuffle(shuffle)
    w = inter.data if loss != "logistic" else np.ones_like(inter.data)
    for rep, fastpath in ((0, 1), (1, 0), (2, 1)):
        fast.set_replay_fast(fastpath)
        t0 = time.perf_counter()
        if loss == "logistic":
            fast.fit_logistic(itf, usf, inter.row, inter.col, inter.data, w, shuffle, holder, 0.05, 0.0, 0.0, 1)
        else:
            getattr(fast, "fit_" + loss)(itf, usf, fast.CSRMatrix(pos), inter.row, inter.col, inter.data, w, shuffle,
                                         holder, 0.05, 0.0, 0.0, 1, rs)
        dt = time.perf_counter() - t0
        c = fast.last_counters["fit"]
        if rep == 1:
            general = round(c["positives"] / c["train_kernel_ms"], 1)
    df = fast.last_replay_dataflow() if loss in ("bpr", "logistic") else None
    print(json.dumps({"config": name, "dataflow_scheduler_ms_kernel_ms_tasks": df, "general_kernel_k_interactions_per_s": general, "loss": loss, "d": d, "nnz": inter.nnz, "mode": c["mode"], "replay_kernel_ms": round(c["train_kernel_ms"], 1),
                      "k_interactions_per_s": round(c["positives"] / c["train_kernel_ms"], 1), "wall_s": round(dt, 3)}), flush=True)
