"""One replay-mode WARP call (for ncu): 138 493 x 26 744 tables, 100 k interactions, d = 64."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightfm_b200 import _lightfm_fast as fast, synthetic
nu, ni, nnz, d = 138_493, 26_744, 100_000, 64
inter = synthetic.interactions(nu, ni, nnz, seed=1)
rs = np.random.RandomState(0)
st = []
for n in (ni, nu):
    emb = ((rs.rand(n, d) - 0.5) / d).astype(np.float32)
    st += [emb, np.ones_like(emb), np.zeros_like(emb), np.zeros(n, np.float32), np.ones(n, np.float32), np.zeros(n, np.float32)]
holder = fast.FastLightFM(*st, d, 0, 0.05, 0.95, 1e-6, 10)
pos = inter.tocsr(); pos.sort_indices()
shuffle = np.arange(inter.nnz, dtype=np.int32); rs.shuffle(shuffle)
fast.fit_warp(fast.CSRMatrix(sp.identity(ni, dtype=np.float32, format="csr")), fast.CSRMatrix(sp.identity(nu, dtype=np.float32, format="csr")),
              fast.CSRMatrix(pos), inter.row, inter.col, inter.data, inter.data, shuffle, holder, 0.05, 0.0, 0.0, 1, rs)
print(fast.last_counters["fit"])
