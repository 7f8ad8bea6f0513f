"""Shared test helpers: synthetic problems + a uniform driver for any backend
exposing the reference's native-module surface (CSRMatrix, FastLightFM, fit_*)."""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

MODEL_ARRAYS = ["item_embeddings", "item_embedding_gradients", "item_embedding_momentum",
                "item_biases", "item_bias_gradients", "item_bias_momentum",
                "user_embeddings", "user_embedding_gradients", "user_embedding_momentum",
                "user_biases", "user_bias_gradients", "user_bias_momentum"]


def synthetic_interactions(n_users, n_items, nnz, seed, signed=False):
    """De-duplicated COO with skewed user activity / item popularity."""
    rng = np.random.default_rng(seed)
    u = np.floor(n_users * rng.random(int(nnz * 1.3)) ** 1.5).astype(np.int64)
    i = np.floor(n_items * rng.random(int(nnz * 1.3)) ** 2.0).astype(np.int64)
    key = np.unique(u * n_items + i)
    rng.shuffle(key)
    key = key[:nnz]
    rows = (key // n_items).astype(np.int32)
    cols = (key % n_items).astype(np.int32)
    if signed:
        data = np.where(rng.random(len(key)) < 0.5, 1.0, -1.0).astype(np.float32)
    else:
        data = np.ones(len(key), dtype=np.float32)
    return sp.coo_matrix((data, (rows, cols)), shape=(n_users, n_items), dtype=np.float32)


def tag_features(n_rows, n_tags, per_row, seed, identity=True, normalize=True):
    """hstack([I, tags]) CSR float32 with Zipf-ish tags, rows L1-normalised."""
    rng = np.random.default_rng(seed)
    cols = np.minimum((n_tags * rng.random((n_rows, per_row)) ** 2.5).astype(np.int64), n_tags - 1)
    r = np.repeat(np.arange(n_rows), per_row)
    tags = sp.coo_matrix((np.ones(r.size, np.float32), (r, cols.ravel())),
                         shape=(n_rows, n_tags)).tocsr()
    tags.sum_duplicates()
    tags.data[:] = 1.0
    mat = sp.hstack([sp.identity(n_rows, dtype=np.float32, format="csr"), tags]).tocsr() \
        if identity else tags
    if normalize:
        s = np.asarray(mat.sum(axis=1)).ravel()
        s[s == 0] = 1.0
        mat = sp.diags((1.0 / s).astype(np.float32)).dot(mat).tocsr()
    mat = mat.astype(np.float32)
    mat.sort_indices()
    return mat


def init_arrays(random_state, n_item_features, n_user_features, d, schedule="adagrad"):
    """Model initialisation, same RNG consumption order as lightfm.py:281-312."""
    a = {}
    a["item_embeddings"] = ((random_state.rand(n_item_features, d) - 0.5) / d).astype(np.float32)
    a["item_embedding_gradients"] = np.zeros_like(a["item_embeddings"])
    a["item_embedding_momentum"] = np.zeros_like(a["item_embeddings"])
    a["item_biases"] = np.zeros(n_item_features, dtype=np.float32)
    a["item_bias_gradients"] = np.zeros_like(a["item_biases"])
    a["item_bias_momentum"] = np.zeros_like(a["item_biases"])
    a["user_embeddings"] = ((random_state.rand(n_user_features, d) - 0.5) / d).astype(np.float32)
    a["user_embedding_gradients"] = np.zeros_like(a["user_embeddings"])
    a["user_embedding_momentum"] = np.zeros_like(a["user_embeddings"])
    a["user_biases"] = np.zeros(n_user_features, dtype=np.float32)
    a["user_bias_gradients"] = np.zeros_like(a["user_biases"])
    a["user_bias_momentum"] = np.zeros_like(a["user_biases"])
    if schedule == "adagrad":
        for k in ("item_embedding_gradients", "item_bias_gradients",
                  "user_embedding_gradients", "user_bias_gradients"):
            a[k] += 1
    return a


def copy_arrays(a):
    return {k: v.copy() for k, v in a.items()}


class Hyper(object):
    def __init__(self, d=16, schedule="adagrad", lr=0.05, rho=0.95, eps=1e-6, max_sampled=10,
                 item_alpha=0.0, user_alpha=0.0, k=5, n=10):
        self.d, self.schedule, self.lr, self.rho, self.eps = d, schedule, lr, rho, eps
        self.max_sampled, self.item_alpha, self.user_alpha = max_sampled, item_alpha, user_alpha
        self.k, self.n = k, n


def holder(api, arrays, hp):
    return api.FastLightFM(*[arrays[k] for k in MODEL_ARRAYS], hp.d,
                           int(hp.schedule == "adadelta"), hp.lr, hp.rho, hp.eps, hp.max_sampled)


def run_epoch(api, loss, interactions, arrays, hp, random_state, item_features=None,
              user_features=None, sample_weight=None, num_threads=1):
    """One epoch exactly as lightfm.py:668-759 drives the native module."""
    inter = interactions.tocoo()
    n_users, n_items = inter.shape
    itf = item_features if item_features is not None else \
        sp.identity(n_items, dtype=np.float32, format="csr")
    usf = user_features if user_features is not None else \
        sp.identity(n_users, dtype=np.float32, format="csr")
    sw = sample_weight if sample_weight is not None else \
        (inter.data if np.array_equiv(inter.data, 1.0) else np.ones_like(inter.data))
    if loss in ("warp", "bpr", "warp-kos"):
        lookup = inter.tocsr()
        if not lookup.has_sorted_indices:
            lookup = lookup.sorted_indices()
        positives = api.CSRMatrix(lookup)
    shuffle = np.arange(len(inter.data), dtype=np.int32)
    random_state.shuffle(shuffle)
    h = holder(api, arrays, hp)
    ci, cu = api.CSRMatrix(itf), api.CSRMatrix(usf)
    if loss == "warp":
        api.fit_warp(ci, cu, positives, inter.row, inter.col, inter.data, sw, shuffle, h, hp.lr,
                     hp.item_alpha, hp.user_alpha, num_threads, random_state)
    elif loss == "bpr":
        api.fit_bpr(ci, cu, positives, inter.row, inter.col, inter.data, sw, shuffle, h, hp.lr,
                    hp.item_alpha, hp.user_alpha, num_threads, random_state)
    elif loss == "warp-kos":
        api.fit_warp_kos(ci, cu, positives, inter.row, shuffle, h, hp.lr, hp.item_alpha,
                         hp.user_alpha, hp.k, hp.n, num_threads, random_state)
    else:
        api.fit_logistic(ci, cu, inter.row, inter.col, inter.data, sw, shuffle, h, hp.lr,
                         hp.item_alpha, hp.user_alpha, num_threads)


def reference_native(variant="strict"):
    """Namespace over the REAL reference's native module (oracle/_ref)."""
    import oracle
    ref = oracle.load_reference(variant)
    import lightfm._lightfm_fast as fast
    return types.SimpleNamespace(
        CSRMatrix=fast.CSRMatrix, FastLightFM=fast.FastLightFM, fit_warp=fast.fit_warp,
        fit_bpr=fast.fit_bpr, fit_warp_kos=fast.fit_warp_kos, fit_logistic=fast.fit_logistic,
        predict_lightfm=fast.predict_lightfm, predict_ranks=fast.predict_ranks,
        calculate_auc_from_rank=fast.calculate_auc_from_rank,
        test_in_positives=fast.__test_in_positives, package=ref)


def oracle_native():
    import oracle
    return oracle


def cuda_native():
    from lightfm_b200 import _lightfm_fast as fast
    return types.SimpleNamespace(
        CSRMatrix=fast.CSRMatrix, FastLightFM=fast.FastLightFM, fit_warp=fast.fit_warp,
        fit_bpr=fast.fit_bpr, fit_warp_kos=fast.fit_warp_kos, fit_logistic=fast.fit_logistic,
        predict_lightfm=fast.predict_lightfm, predict_ranks=fast.predict_ranks,
        calculate_auc_from_rank=fast.calculate_auc_from_rank,
        test_in_positives=getattr(fast, "__test_in_positives"), module=fast)


def max_rel_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))) if a.size else 0.0


# ---- golden fixtures (tests/golden/*.npz, generated from the real reference) -------------
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(f[len("golden_"):-4] for f in os.listdir(GOLDEN_DIR)
                  if f.startswith("golden_") and f.endswith(".npz"))


class ReplayState(object):
    """Stands in for numpy RandomState: replays the shuffles / seeds the reference drew."""

    def __init__(self, shuffles, seeds):
        self.shuffles, self.seeds = list(shuffles), list(seeds)

    def shuffle(self, arr):
        arr[:] = self.shuffles.pop(0)

    def randint(self, *a, **k):
        return self.seeds.pop(0).astype(np.int64)


def _csr(g, prefix):
    shape = tuple(int(x) for x in g[prefix + "_shape"])
    return sp.csr_matrix((g[prefix + "_data"].copy(), g[prefix + "_indices"].copy(),
                          g[prefix + "_indptr"].copy()), shape=shape)


def run_golden(api, name, num_threads=1):
    """Replay one golden case through `api`; returns (outputs, golden) dicts."""
    g = np.load(os.path.join(GOLDEN_DIR, "golden_%s.npz" % name))
    hy = g["hyper"]
    hp = Hyper(d=int(hy[0]), schedule="adadelta" if hy[1] else "adagrad", lr=float(hy[2]),
               rho=float(hy[3]), eps=float(hy[4]), max_sampled=int(hy[5]), item_alpha=float(hy[6]),
               user_alpha=float(hy[6]), k=int(hy[7]), n=int(hy[8]))
    loss = str(g["loss"])
    shape = tuple(int(x) for x in g["shape"])
    inter = sp.coo_matrix((g["data"].copy(), (g["row"].copy(), g["col"].copy())), shape=shape)
    itf, usf = _csr(g, "itf"), _csr(g, "usf")
    sw = g["sample_weight"].copy() if "sample_weight" in g.files else None
    arrays = {k: g["init_" + k].copy() for k in MODEL_ARRAYS}
    rs = ReplayState(g["shuffles"], g["seeds"])
    for _ in range(int(hy[9])):
        run_epoch(api, loss, inter, arrays, hp, rs, itf, usf, sw, num_threads=num_threads)
    out = {"final_" + k: v for k, v in arrays.items()}
    # scoring runs on the GOLDEN final state so that each function is pinned on its own
    gold_arrays = {k: g["final_" + k].copy() for k in MODEL_ARRAYS}
    h = holder(api, gold_arrays, hp)
    pred = np.empty(len(g["pred_users"]), dtype=np.float32)
    api.predict_lightfm(api.CSRMatrix(itf), api.CSRMatrix(usf), g["pred_users"].copy(),
                        g["pred_items"].copy(), pred, h, 1)
    out["pred"] = pred
    test, train = _csr(g, "test"), _csr(g, "train")
    ranks = np.zeros_like(test.data)
    api.predict_ranks(api.CSRMatrix(itf), api.CSRMatrix(usf), api.CSRMatrix(test),
                      api.CSRMatrix(train), ranks, h, 1)
    out["ranks"] = ranks.copy()
    auc = np.zeros(test.shape[0], dtype=np.float32)
    rk = sp.csr_matrix((g["ranks"].copy(), test.indices, test.indptr), shape=test.shape)
    api.calculate_auc_from_rank(api.CSRMatrix(rk), g["num_train_positives"].copy(), rk.data, auc, 1)
    out["auc"], out["ranks_sorted"] = auc, rk.data.copy()
    return out, g


def planted_interactions(n_users, n_items, per_user, seed, rank=4, temperature=1.5):
    """Interactions with learnable low-rank structure (for statistical parity tests)."""
    rng = np.random.default_rng(seed)
    ul = rng.normal(size=(n_users, rank))
    il = rng.normal(size=(n_items, rank))
    pop = rng.normal(size=n_items) * 0.5
    scores = temperature * (ul @ il.T) + pop[None, :]
    g = rng.gumbel(size=scores.shape)
    top = np.argsort(-(scores + g), axis=1)[:, :per_user]
    rows = np.repeat(np.arange(n_users), per_user).astype(np.int32)
    cols = top.ravel().astype(np.int32)
    return sp.coo_matrix((np.ones(rows.size, np.float32), (rows, cols)), shape=(n_users, n_items))


def split(inter, seed, frac=0.8):
    rs = np.random.RandomState(seed)
    order = np.arange(inter.nnz)
    rs.shuffle(order)
    cut = int(frac * inter.nnz)
    mk = lambda ix: sp.coo_matrix((inter.data[ix], (inter.row[ix], inter.col[ix])), shape=inter.shape)
    return mk(order[:cut]), mk(order[cut:])


def eval_arrays(arrays, d, train, test, k=10):
    """Held-out precision@k and AUC of a weight set, computed in numpy (identity features)."""
    scores = arrays["user_embeddings"].astype(np.float64) @ arrays["item_embeddings"].astype(np.float64).T
    scores += arrays["user_biases"][:, None] + arrays["item_biases"][None, :]
    tr = train.tocsr()
    te = test.tocsr()
    precs, aucs = [], []
    for u in range(te.shape[0]):
        t = te[u].indices
        if len(t) == 0:
            continue
        s = scores[u].copy()
        s[tr[u].indices] = -np.inf
        order = np.argsort(-s)
        topk = set(order[:k].tolist())
        precs.append(len(topk & set(t.tolist())) / k)
        neg = np.ones(len(s), bool)
        neg[tr[u].indices] = False
        neg[t] = False
        if neg.sum() == 0:
            continue
        ns = np.sort(s[neg])
        rank_below = np.searchsorted(ns, s[t], side="left")
        aucs.append(float(np.mean(rank_below / neg.sum())))
    return float(np.mean(precs)), float(np.mean(aucs))


# ---- tier-B (statistical) parity: scalable planted problems + a subset evaluator ------------
def planted_clusters(n_users, n_items, nnz, seed, n_clusters=32, p_in=0.8, signed=False):
    """Interactions with learnable structure at any shape, O(nnz) to generate: users and
    items belong to one of `n_clusters` groups; a user draws an item from its own group with
    probability `p_in` (popularity-skewed inside the group), else from the whole catalogue.
    User activity ~ r^1.5 and item popularity ~ r^2 as in SURVEY 8(d)."""
    rng = np.random.default_rng(seed)
    ucl = rng.integers(0, n_clusters, n_users)
    icl = rng.integers(0, n_clusters, n_items)
    order = np.argsort(icl, kind="stable")
    starts = np.searchsorted(icl[order], np.arange(n_clusters + 1))
    m = int(nnz * 1.5) + 64
    u = np.floor(n_users * rng.random(m) ** 1.5).astype(np.int64)
    c = ucl[u]
    size = (starts[c + 1] - starts[c]).astype(np.int64)
    inside = (rng.random(m) < p_in) & (size > 0)
    pos = np.floor(size * rng.random(m) ** 2.0).astype(np.int64)
    i_in = order[np.minimum(starts[c] + pos, n_items - 1)]
    i_out = np.floor(n_items * rng.random(m) ** 2.0).astype(np.int64)
    i = np.where(inside, i_in, i_out)
    key = np.unique(u * n_items + i)
    rng.shuffle(key)
    key = key[:nnz]
    rows = (key // n_items).astype(np.int32)
    cols = (key % n_items).astype(np.int32)
    if signed:
        data = np.where(rng.random(len(key)) < 0.5, 1.0, -1.0).astype(np.float32)
    else:
        data = np.ones(len(key), dtype=np.float32)
    return sp.coo_matrix((data, (rows, cols)), shape=(n_users, n_items), dtype=np.float32)


def eval_subset(arrays, train, test, users, k=10):
    """Held-out precision@k and AUC (train positives excluded, identity features) of a weight
    set on the given users, in float64 numpy.  Users without test items are skipped."""
    users = np.asarray(users)
    ue = arrays["user_embeddings"][users].astype(np.float64)
    scores = ue @ arrays["item_embeddings"].astype(np.float64).T
    scores += arrays["user_biases"][users].astype(np.float64)[:, None]
    scores += arrays["item_biases"].astype(np.float64)[None, :]
    tr, te = train.tocsr(), test.tocsr()
    n_items = scores.shape[1]
    precs, aucs = [], []
    for j, u in enumerate(users):
        t = te.indices[te.indptr[u]:te.indptr[u + 1]]
        if len(t) == 0:
            continue
        s = scores[j]
        tpos = tr.indices[tr.indptr[u]:tr.indptr[u + 1]]
        ts = s[t].copy()
        s[tpos] = -np.inf                                  # never recommended
        top = np.argpartition(-s, k)[:k]
        precs.append(len(np.intersect1d(top, t)) / k)
        nn = n_items - len(np.union1d(tpos, t))            # negatives: neither train nor test
        if nn <= 0:
            continue
        s[tpos] = np.inf                                   # ... and never counted as "below"
        below = (s[None, :] < ts[:, None]).sum(axis=1) - (ts[None, :] < ts[:, None]).sum(axis=1)
        aucs.append(float(np.mean(below / nn)))
    return float(np.mean(precs)), float(np.mean(aucs))


def data_digest(inter):
    """Short fingerprint of a COO matrix (generator drift between hosts would change it)."""
    import hashlib
    h = hashlib.sha256()
    for a in (inter.row, inter.col, inter.data):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


# The tier-B problem set (SURVEY 8(c) tier B).  Shared by the generator of
# tests/golden/tierb_bands.json (run where the real reference is built) and the GPU tests.
TIERB = {
    # C1 shape, 10 epochs, d=16 (SURVEY's own tier-B setting)
    "c1_warp": dict(shape=(943, 1682), nnz=100_000, clusters=8, loss="warp", d=16, epochs=10),
    "c1_bpr": dict(shape=(943, 1682), nnz=100_000, clusters=8, loss="bpr", d=16, epochs=10),
    "c1_logistic": dict(shape=(943, 1682), nnz=100_000, clusters=8, loss="logistic", d=16, epochs=10,
                        signed=True),
    "c1_kos": dict(shape=(943, 1682), nnz=100_000, clusters=8, loss="warp-kos", d=16, epochs=10),
    # C2 shape (138 493 x 26 744), 2.5 M interactions: a full wave of slots in flight on a B200
    "c2_warp": dict(shape=(138_493, 26_744), nnz=2_500_000, clusters=64, loss="warp", d=64, epochs=3),
    # (BPR / logistic take small steps: at lr 0.05 three epochs of ~14 interactions per user leave the
    #  model at chance, so these two run 5 epochs at lr 0.2, where the reference reaches AUC ~0.88-0.90)
    "c2_bpr": dict(shape=(138_493, 26_744), nnz=2_500_000, clusters=64, loss="bpr", d=64, epochs=5, lr=0.2),
    "c2_logistic": dict(shape=(138_493, 26_744), nnz=2_500_000, clusters=64, loss="logistic", d=32,
                        epochs=5, lr=0.2, signed=True),
    "c2_kos": dict(shape=(138_493, 26_744), nnz=2_500_000, clusters=64, loss="warp-kos", d=64, epochs=2),
}
TIERB_SEEDS = (0, 1, 2, 3, 4)


_TIERB_CACHE = {}


def tierb_problem(name):
    """(fit matrix, positives-only train for exclusion, test, eval users) of one tier-B case."""
    key = (TIERB[name]["shape"], TIERB[name]["nnz"], bool(TIERB[name].get("signed")))
    if key not in _TIERB_CACHE:
        _TIERB_CACHE[key] = _tierb_problem(name)
    return _TIERB_CACHE[key]


def _tierb_problem(name):
    cfg = TIERB[name]
    n_users, n_items = cfg["shape"]
    base = (cfg["shape"], cfg["nnz"], cfg["clusters"])
    if base not in _TIERB_CACHE:
        full = planted_clusters(n_users, n_items, cfg["nnz"], seed=11, n_clusters=cfg["clusters"])
        _TIERB_CACHE[base] = split(full, 7)
    train, test = _TIERB_CACHE[base]
    if cfg.get("signed"):
        # logistic needs both classes: add as many uniformly drawn explicit negatives as positives
        rng = np.random.default_rng(3)
        nr = rng.integers(0, n_users, train.nnz).astype(np.int32)
        nc = rng.integers(0, n_items, train.nnz).astype(np.int32)
        fit = sp.coo_matrix((np.concatenate([train.data, -np.ones(train.nnz, np.float32)]),
                             (np.concatenate([train.row, nr]), np.concatenate([train.col, nc]))),
                            shape=train.shape)
    else:
        fit = train
    te = test.tocsr()
    has = np.flatnonzero(np.diff(te.indptr) > 0)
    rng = np.random.default_rng(5)
    users = np.sort(rng.choice(has, size=min(1500, len(has)), replace=False))
    return fit, train, test, users


def tierb_fit(api, name, seed, num_threads):
    """Train one tier-B case through `api`'s native-module surface; returns the weight arrays."""
    cfg = TIERB[name]
    fit, _, _, _ = tierb_problem(name)
    hp = Hyper(d=cfg["d"], lr=cfg.get("lr", 0.05))
    rs = np.random.RandomState(seed)
    arr = init_arrays(rs, fit.shape[1], fit.shape[0], cfg["d"])
    for _ in range(cfg["epochs"]):
        run_epoch(api, cfg["loss"], fit, arr, hp, rs, num_threads=num_threads)
    return arr
