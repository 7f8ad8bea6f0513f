"""Deterministic synthetic interaction data in the shapes of BASELINE.json's configs
(SURVEY 8(d)): skewed user activity ~ floor(U r^1.5), item popularity ~ floor(I r^2),
de-duplicated COO with int32 indices and float32 values."""
import numpy as np
import scipy.sparse as sp

CONFIGS = {
    # name: (users, items, nnz, loss, no_components)
    "C1": (943, 1682, 100_000, "bpr", 16),
    "C2": (138_493, 26_744, 20_000_000, "warp", 64),
    "C3": (138_493, 26_744, 20_000_000, "warp", 128),
    "C4": (10_000_000, 1_000_000, 500_000_000, "warp-kos", 64),
    "C5": (1_000_000, 100_000, 100_000_000, "logistic", 32),
}


def interactions(n_users, n_items, nnz, seed, signed=False):
    """COO float32 matrix with exactly min(nnz, available) distinct entries, in random order."""
    rng = np.random.default_rng(seed)
    keys = np.empty(0, dtype=np.int64)
    want = int(nnz)
    draw = int(want * 1.25) + 16
    while keys.size < want:
        u = np.floor(n_users * rng.random(draw) ** 1.5).astype(np.int64)
        i = np.floor(n_items * rng.random(draw) ** 2.0).astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
        draw = max(int((want - keys.size) * 2.0) + 16, 1024)
        if keys.size >= n_users * n_items:
            break
    rng.shuffle(keys)
    keys = keys[:want]
    rows = (keys // n_items).astype(np.int32)
    cols = (keys % n_items).astype(np.int32)
    if signed:
        data = np.where(rng.random(keys.size) < 0.5, 1.0, -1.0).astype(np.float32)
    else:
        data = np.ones(keys.size, dtype=np.float32)
    return sp.coo_matrix((data, (rows, cols)), shape=(n_users, n_items), dtype=np.float32)


def tag_features(n_rows, n_tags, per_row, seed):
    """hstack([I, tags]) with Zipf-like tags, rows L1-normalised (as Dataset.build_item_features)."""
    rng = np.random.default_rng(seed)
    cols = np.minimum((n_tags * rng.random((n_rows, per_row)) ** 2.5).astype(np.int64), n_tags - 1)
    r = np.repeat(np.arange(n_rows), per_row)
    tags = sp.coo_matrix((np.ones(r.size, np.float32), (r, cols.ravel())), shape=(n_rows, n_tags)).tocsr()
    tags.sum_duplicates()
    tags.data[:] = 1.0
    mat = sp.hstack([sp.identity(n_rows, dtype=np.float32, format="csr"), tags]).tocsr()
    s = np.asarray(mat.sum(axis=1)).ravel()
    s[s == 0] = 1.0
    mat = sp.diags((1.0 / s).astype(np.float32)).dot(mat).tocsr().astype(np.float32)
    mat.sort_indices()
    return mat
