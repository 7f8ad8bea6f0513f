"""GPU diagnostic: the reference's tests/test_api.py::test_predict_ranks scenario, GPU ranks vs oracle."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from lightfm_b200 import LightFM  # noqa: E402

orc, cu = H.oracle_native(), H.cuda_native()
no_users, no_items = 10, 100
train = sp.rand(no_users, no_items, format="csr", random_state=42)
model = LightFM()
model.fit_partial(train)
rank_input = sp.csr_matrix(np.ones((no_users, no_items)))
for rep in range(3):
    ranks = model.predict_rank(rank_input, num_threads=2)
    test = sp.csr_matrix(rank_input, dtype=np.float32)
    arr = {k: getattr(model, k) for k in H.MODEL_ARRAYS}
    hp = H.Hyper(d=model.no_components)
    want = np.zeros(test.nnz, np.float32)
    empty = sp.csr_matrix((no_users, no_items), dtype=np.float32)
    ii = sp.identity(no_items, dtype=np.float32, format="csr")
    iu = sp.identity(no_users, dtype=np.float32, format="csr")
    orc.predict_ranks(orc.CSRMatrix(ii), orc.CSRMatrix(iu), orc.CSRMatrix(test), orc.CSRMatrix(empty), want,
                      H.holder(orc, arr, hp), 1)
    got = ranks.data
    bad = np.flatnonzero(got != want)
    print("rep", rep, "mismatches", len(bad), "of", len(want), "| oracle permutation ok:",
          all(np.array_equal(np.sort(want[r * 100:(r + 1) * 100]), np.arange(100)) for r in range(10)))
    for b in bad[:10]:
        u, it = b // 100, b % 100
        s = model.predict(int(u), np.arange(no_items, dtype=np.int32))
        print("  user", u, "item", it, "gpu", got[b], "oracle", want[b], "score", repr(s[it]),
              "n_equal_scores", int((s == s[it]).sum()), "user_emb_norm", float(np.abs(model.user_embeddings[u]).sum()))
print("user biases", model.user_biases[:10], "nan anywhere:", any(np.isnan(v).any() for v in arr.values()))
