"""GPU: the fused evaluation epilogue, the hand-written per-row rank sort
behind calculate_auc_from_rank (T:1352) and the top-k recommend entry,
each against the oracle / the reference's own numpy post-processing."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

pytestmark = pytest.mark.gpu


def _model(n_users, n_items, d, seed, loss="warp"):
    from lightfm_b200 import LightFM
    rs = np.random.RandomState(seed)
    m = LightFM(loss=loss, no_components=d, random_state=seed)
    m._initialize(d, n_items, n_users)
    m.item_embeddings[:] = rs.normal(size=(n_items, d)).astype(np.float32)
    m.user_embeddings[:] = rs.normal(size=(n_users, d)).astype(np.float32)
    m.item_biases[:] = rs.normal(size=n_items).astype(np.float32)
    m.user_biases[:] = rs.normal(size=n_users).astype(np.float32)
    return m


def _split_matrices(n_users, n_items, seed, heavy=()):
    """train / test CSR without intersections; `heavy` = (user, n_test) rows with many test items."""
    rng = np.random.default_rng(seed)
    rows, cols, part = [], [], []
    for u in range(n_users):
        n = int(rng.integers(0, 40))
        for user, nt in heavy:
            if u == user:
                n = nt + 20
        if n == 0:
            continue
        it = rng.choice(n_items, size=min(n, n_items), replace=False)
        rows += [u] * len(it)
        cols += it.tolist()
        is_test = rng.random(len(it)) < 0.4
        for user, nt in heavy:
            if u == user:
                is_test = np.arange(len(it)) < nt
        part += is_test.tolist()
    rows, cols, part = np.array(rows, np.int32), np.array(cols, np.int32), np.array(part, bool)
    mk = lambda m: sp.csr_matrix((np.ones(int(m.sum()), np.float32), (rows[m], cols[m])), shape=(n_users, n_items))
    return mk(~part), mk(part)


def _host_metrics(model, test, train, k):
    """The reference's evaluation.py reductions (E:71-87, 150-166, 231-254, 312-327) in numpy on
    the rank matrix of predict_rank."""
    from lightfm_b200._lightfm_fast import CSRMatrix, calculate_auc_from_rank
    ranks = model.predict_rank(test, train_interactions=train, num_threads=2)
    hit = ranks.copy()
    hit.data = np.less(hit.data, k, hit.data)
    precision = np.squeeze(np.array(hit.sum(axis=1))) / k
    recall = np.squeeze(np.array(hit.sum(axis=1))) / np.squeeze(test.getnnz(axis=1))
    rr = ranks.copy()
    rr.data = 1.0 / (rr.data + 1.0)
    rr = np.squeeze(np.array(rr.max(axis=1).todense()))
    auc = np.zeros(ranks.shape[0], dtype=np.float32)
    ntp = np.squeeze(np.array(train.getnnz(axis=1)).astype(np.int32))
    calculate_auc_from_rank(CSRMatrix(ranks), np.ascontiguousarray(ntp), ranks.data, auc, 2)
    return precision, recall, rr, auc


@pytest.mark.parametrize("n_items,d", [(1003, 10), (6000, 32)])
def test_fused_metrics_equal_host_reductions(n_items, d):
    from lightfm_b200 import evaluation as ev
    n_users = 300
    model = _model(n_users, n_items, d, 1)
    heavy = ((7, 4500), (11, 200)) if n_items >= 6000 else ((7, 120),)
    train, test = _split_matrices(n_users, n_items, 2, heavy=heavy)
    for k in (1, 10):
        precision, recall, rr, auc = _host_metrics(model, test, train, k)
        with np.errstate(invalid="ignore", divide="ignore"):
            assert np.array_equal(ev.precision_at_k(model, test, train, k=k, preserve_rows=True, num_threads=2), precision)
            got = ev.recall_at_k(model, test, train, k=k, preserve_rows=True, num_threads=2)
            assert np.array_equal(got, recall, equal_nan=True)
        assert np.array_equal(ev.reciprocal_rank(model, test, train, preserve_rows=True, num_threads=2), rr)
        assert np.array_equal(ev.auc_score(model, test, train, preserve_rows=True, num_threads=2), auc)
    keep = test.getnnz(axis=1) > 0
    assert np.array_equal(ev.auc_score(model, test, train, num_threads=2), auc[keep])
    assert ev.precision_at_k(model, test, train, k=5).shape == (int(keep.sum()),)


def test_row_sort_and_auc_match_oracle_for_all_row_lengths():
    """T:1352: rows of 0, 1, 2..32 (warp path), 33..4096 (shared-memory bitonic) and > 4096 values."""
    cu, orc = H.cuda_native(), H.oracle_native()
    rng = np.random.default_rng(3)
    lens = [0, 1, 2, 5, 31, 32, 33, 64, 100, 1000, 4096, 4097, 6000, 0, 3]
    n_cols = 7000
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    indices = np.concatenate([np.sort(rng.choice(n_cols, size=n, replace=False)) for n in lens]).astype(np.int32)
    data = rng.integers(0, 50, size=indptr[-1]).astype(np.float32)   # many ties
    ntp = rng.integers(0, 30, size=len(lens)).astype(np.int32)
    outs = []
    for api in (orc, cu):
        m = sp.csr_matrix((data.copy(), indices, indptr), shape=(len(lens), n_cols))
        auc = np.zeros(len(lens), np.float32)
        api.calculate_auc_from_rank(api.CSRMatrix(m), ntp, m.data, auc, 1)
        outs.append((m.data.copy(), auc))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("n_items,d,k", [(1003, 10, 10), (5000, 64, 100), (37, 16, 64), (5000, 32, 16), (1003, 10, 1),
                                          (9, 8, 16), (20011, 16, 7)])
def test_recommend_equals_argsort_of_predict(n_items, d, k):
    n_users = 120
    model = _model(n_users, n_items, d, 4)
    model.item_embeddings[5] = model.item_embeddings[3]      # exact ties: lower item id first
    model.item_biases[5] = model.item_biases[3]
    model.item_embeddings[n_items - 1] = model.item_embeddings[0]
    model.item_biases[n_items - 1] = model.item_biases[0]
    train, _ = _split_matrices(n_users, n_items, 6)
    users = np.array([0, 3, 3, 50, 119, 17], dtype=np.int32)
    for tr in (None, train):
        items, scores = model.recommend(users, k=k, train_interactions=tr)
        assert items.shape == (len(users), k) and scores.dtype == np.float32
        for r, u in enumerate(users):
            s = model.predict(int(u), np.arange(n_items, dtype=np.int32))
            order = np.argsort(-s, kind="stable")
            if tr is not None:
                seen = set(tr[int(u)].indices.tolist())
                order = np.array([i for i in order if i not in seen], dtype=np.int64)
            want = order[:k]
            assert np.array_equal(items[r, :len(want)], want), (u, items[r, :8], want[:8])
            assert np.array_equal(scores[r, :len(want)], s[want])
            assert np.all(items[r, len(want):] == -1)


def test_predict_ranks_register_tiled_kernel_odd_shapes():
    """n_items not a multiple of 4, d not a multiple of 4, duplicated test entries, users without
    train rows: bit-equal ranks to the oracle."""
    cu, orc = H.cuda_native(), H.oracle_native()
    n_users, n_items, d = 77, 1003, 10
    rs = np.random.RandomState(2)
    arr = H.init_arrays(rs, n_items, n_users, d)
    arr["item_embeddings"][:] = rs.normal(size=(n_items, d)).astype(np.float32)
    arr["user_embeddings"][:] = rs.normal(size=(n_users, d)).astype(np.float32)
    arr["item_biases"][:] = rs.normal(size=n_items).astype(np.float32)
    arr["item_embeddings"][10] = arr["item_embeddings"][20]   # ties with a test item
    arr["item_biases"][10] = arr["item_biases"][20]
    train, test = _split_matrices(n_users, n_items, 8, heavy=((5, 90),))
    test = test.tolil()
    test[3, 10] = 1.0
    test[3, 20] = 1.0
    test = test.tocsr().astype(np.float32)
    hp = H.Hyper(d=d)
    ident_i = sp.identity(n_items, dtype=np.float32, format="csr")
    ident_u = sp.identity(n_users, dtype=np.float32, format="csr")
    outs = []
    for api in (orc, cu):
        ranks = np.zeros(test.nnz, np.float32)
        api.predict_ranks(api.CSRMatrix(ident_i), api.CSRMatrix(ident_u), api.CSRMatrix(test),
                          api.CSRMatrix(train), ranks, H.holder(api, arr, hp), 1)
        outs.append(ranks)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("groups", (1, 2, 3))
def test_predict_ranks_dense_test_rows_both_tilings(groups):
    """The reference's tests/test_api.py::test_predict_ranks scenario (every item of every user is a
    test interaction: 100 > 64 test entries per user, i.e. two chunks per tile), repeated over seeds
    and a larger catalogue, for both CTA layouts of the rank kernel: bit-equal to the oracle."""
    cu, orc = H.cuda_native(), H.oracle_native()
    old = cu.module.set_rank_groups(groups)
    try:
        for seed, (n_users, n_items, d) in enumerate([(10, 100, 10)] * 6 + [(53, 1500, 10), (40, 3000, 32)]):
            rs = np.random.RandomState(seed)
            arr = H.init_arrays(rs, n_items, n_users, d)
            arr["item_embeddings"][:] = rs.normal(size=(n_items, d)).astype(np.float32)
            arr["user_embeddings"][:] = rs.normal(size=(n_users, d)).astype(np.float32)
            arr["item_biases"][:] = rs.normal(size=n_items).astype(np.float32)
            test = sp.csr_matrix(np.ones((n_users, min(n_items, 150)), np.float32))
            test.resize((n_users, n_items))
            test = sp.csr_matrix(test)
            train = sp.rand(n_users, n_items, density=0.02, format="csr", random_state=42 + seed).astype(np.float32)
            train = train - train.multiply(test)
            train.eliminate_zeros()
            train = sp.csr_matrix(train, dtype=np.float32)
            train.sort_indices()
            hp = H.Hyper(d=d)
            ii = sp.identity(n_items, dtype=np.float32, format="csr")
            iu = sp.identity(n_users, dtype=np.float32, format="csr")
            outs = []
            for api in (orc, cu):
                ranks = np.zeros(test.nnz, np.float32)
                api.predict_ranks(api.CSRMatrix(ii), api.CSRMatrix(iu), api.CSRMatrix(test), api.CSRMatrix(train),
                                  ranks, H.holder(api, arr, hp), 2)
                outs.append(ranks)
            assert np.array_equal(outs[0], outs[1]), (groups, seed, int((outs[0] != outs[1]).sum()))
    finally:
        cu.module.set_rank_groups(old)
