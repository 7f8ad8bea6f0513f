"""GPU: predict / predict_rank / AUC kernels against the oracle, plus the exact structural
properties the reference pins in tests/test_api.py:217-282 and tests/test_evaluation.py."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

pytestmark = pytest.mark.gpu


def _fitted(loss="warp", d=10, users=60, items=50, nnz=900, seed=3, feats=False):
    from lightfm_b200 import LightFM
    inter = H.synthetic_interactions(users, items, nnz, seed)
    itf = H.tag_features(items, 12, 3, 9) if feats else None
    usf = H.tag_features(users, 8, 2, 10) if feats else None
    model = LightFM(loss=loss, no_components=d, random_state=1)
    model.fit(inter, item_features=itf, user_features=usf, epochs=2, num_threads=1)
    return model, inter, itf, usf


@pytest.mark.parametrize("feats", (False, True))
@pytest.mark.parametrize("d", (10, 64))
def test_predict_bit_equal_to_oracle(feats, d):
    model, inter, itf, usf = _fitted(d=d, feats=feats)
    orc = H.oracle_native()
    rng = np.random.default_rng(0)
    u = rng.integers(0, inter.shape[0], 500).astype(np.int32)
    i = rng.integers(0, inter.shape[1], 500).astype(np.int32)
    got = model.predict(u, i, item_features=itf, user_features=usf)
    arrays = {k: getattr(model, k) for k in H.MODEL_ARRAYS}
    hp = H.Hyper(d=d)
    want = np.empty(500, dtype=np.float32)
    ci = orc.CSRMatrix(itf if feats else sp.identity(inter.shape[1], dtype=np.float32, format="csr"))
    cu = orc.CSRMatrix(usf if feats else sp.identity(inter.shape[0], dtype=np.float32, format="csr"))
    orc.predict_lightfm(ci, cu, u, i, want, H.holder(orc, arrays, hp), 1)
    assert np.array_equal(got, want)


def test_predict_equals_representation_dot():  # reference tests/test_movielens.py:320-351
    model, inter, itf, usf = _fitted(feats=True)
    ub, ue = model.get_user_representations(usf)
    ib, ie = model.get_item_representations(itf)
    u = np.repeat(np.arange(inter.shape[0]), inter.shape[1]).astype(np.int32)
    i = np.tile(np.arange(inter.shape[1]), inter.shape[0]).astype(np.int32)
    pred = model.predict(u, i, item_features=itf, user_features=usf)
    want = (ue @ ie.T + ub[:, None] + ib[None, :]).ravel()
    assert np.allclose(pred, want, atol=1e-6)


def test_predict_scalar_user_and_lists():  # tests/test_api.py:77-92
    model, inter, _, _ = _fitted()
    items = np.arange(10, dtype=np.int32)
    a = model.predict(3, items)
    b = model.predict(np.repeat(3, 10), items)
    c = model.predict([3] * 10, list(range(10)))
    assert np.array_equal(a, b) and np.array_equal(a, c)
    with pytest.raises(ValueError):
        model.predict(np.arange(3), np.arange(4))
    with pytest.raises(ValueError):
        model.predict(np.array([-1, 0]), np.array([0, 1]))
    with pytest.raises(ValueError):  # int64 overflow wraps negative (tests/test_api.py:354-371)
        model.predict(np.array([2 ** 31, 0]), np.array([0, 1]))


def test_predict_rank_structure():  # tests/test_api.py:217-282
    from lightfm_b200 import LightFM
    users, items = 30, 25
    rng = np.random.RandomState(0)
    train = sp.coo_matrix((rng.rand(users, items) > 0.7).astype(np.float32))
    model = LightFM(loss="bpr", no_components=8, random_state=2).fit(train, epochs=2)
    dense = sp.csr_matrix(np.ones((users, items), dtype=np.float32))
    ranks = model.predict_rank(dense, check_intersections=False).todense()
    want = np.tile(np.arange(items), (users, 1))
    assert np.array_equal(np.sort(np.asarray(ranks), axis=1), want)
    # train exclusion: the largest possible rank is n_items - 1 - nnz_train(user)
    test = sp.csr_matrix((np.asarray(train.todense()) == 0).astype(np.float32))
    ranks = model.predict_rank(test, train_interactions=train.tocsr(), check_intersections=False)
    mx = np.asarray(ranks.max(axis=1).todense()).ravel()
    lim = items - 1 - np.asarray(train.tocsr().getnnz(axis=1)).ravel()
    assert np.all(mx <= lim)
    # pessimistic ties: all-equal scores put every item at rank n_items - 1
    model.item_embeddings[:] = 0
    model.user_embeddings[:] = 0
    model.item_biases[:] = 0
    model.user_biases[:] = 0
    ranks = model.predict_rank(dense, check_intersections=False)
    assert np.all(ranks.data == items - 1)
    with pytest.raises(ValueError):
        model.predict_rank(train.tocsr(), train_interactions=train.tocsr())
    with pytest.raises(Exception):
        model.predict_rank(sp.csr_matrix((users + 5, items), dtype=np.float32),
                           user_features=sp.identity(users, dtype=np.float32, format="csr"))


@pytest.mark.parametrize("feats", (False, True))
def test_predict_rank_and_auc_bit_equal_to_oracle(feats):
    model, inter, itf, usf = _fitted(users=80, items=70, nnz=1500, feats=feats, d=16)
    orc = H.oracle_native()
    train = inter.tocsr().astype(np.float32)
    train.sort_indices()
    cand = H.synthetic_interactions(80, 70, 600, 77).tocsr().astype(bool)
    test = (cand > train.astype(bool)).astype(np.float32).tocsr()
    test.sort_indices()
    got = model.predict_rank(test, train_interactions=train, item_features=itf, user_features=usf)
    arrays = {k: getattr(model, k) for k in H.MODEL_ARRAYS}
    hp = H.Hyper(d=16)
    ci = orc.CSRMatrix(itf if feats else sp.identity(70, dtype=np.float32, format="csr"))
    cu = orc.CSRMatrix(usf if feats else sp.identity(80, dtype=np.float32, format="csr"))
    want = np.zeros_like(test.data)
    orc.predict_ranks(ci, cu, orc.CSRMatrix(test), orc.CSRMatrix(train), want,
                      H.holder(orc, arrays, hp), 1)
    assert np.array_equal(got.data, want)

    from lightfm_b200.evaluation import auc_score, precision_at_k, recall_at_k, reciprocal_rank
    auc = auc_score(model, test, train_interactions=train, item_features=itf, user_features=usf,
                    preserve_rows=True)
    ntp = np.asarray(train.getnnz(axis=1)).astype(np.int32)
    want_auc = np.zeros(80, dtype=np.float32)
    rk = sp.csr_matrix((want.copy(), test.indices, test.indptr), shape=test.shape)
    orc.calculate_auc_from_rank(orc.CSRMatrix(rk), ntp, rk.data, want_auc, 1)
    assert np.array_equal(auc, want_auc)

    # metrics against brute force built on predict (reference tests/test_evaluation.py:34-161)
    k = 5
    prec = precision_at_k(model, test, train_interactions=train, k=k, item_features=itf,
                          user_features=usf, preserve_rows=True)
    rec = recall_at_k(model, test, train_interactions=train, k=k, item_features=itf,
                      user_features=usf, preserve_rows=True)
    rr = reciprocal_rank(model, test, train_interactions=train, item_features=itf,
                         user_features=usf, preserve_rows=True)
    for user in range(80):
        scores = model.predict(user, np.arange(70, dtype=np.int32), item_features=itf,
                               user_features=usf).astype(np.float64)
        tr = set(train[user].indices)
        te = set(test[user].indices)
        if not te:
            assert prec[user] == 0 and rr[user] == 0
            continue
        # rank of t = #{i not in train, i != t, s_i >= s_t}
        rank_of = {t: sum(1 for i in range(70) if i not in tr and i != t and scores[i] >= scores[t])
                   for t in te}
        hits = sum(1 for t in te if rank_of[t] < k)
        assert abs(prec[user] - hits / k) < 1e-6
        assert abs(rec[user] - hits / len(te)) < 1e-6
        assert abs(rr[user] - 1.0 / (min(rank_of.values()) + 1)) < 1e-6


def test_in_positives_known_answers():  # reference tests/test_fast_functions.py:9-17
    cu = H.cuda_native()
    mat = cu.CSRMatrix(sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=np.float32)))
    assert not cu.test_in_positives(0, 0, mat)
    assert cu.test_in_positives(0, 1, mat)
    assert cu.test_in_positives(1, 0, mat)
    assert not cu.test_in_positives(1, 1, mat)
    # long rows exercise the 32-ary narrowing
    rng = np.random.default_rng(0)
    cols = np.sort(rng.choice(100000, 5000, replace=False)).astype(np.int32)
    big = sp.csr_matrix((np.ones(5000, np.float32), cols, np.array([0, 5000], np.int32)),
                        shape=(1, 100000))
    c = cu.CSRMatrix(big)
    present = set(cols.tolist())
    for col in list(cols[:40]) + list(cols[-40:]) + rng.integers(0, 100000, 80).tolist():
        assert cu.test_in_positives(0, int(col), c) == (int(col) in present)


def test_compact_predict_is_bit_identical_to_full_predict():
    """predict() on a small batch gathers only the touched rows (LightFM._predict_compact); the
    scores must equal the full-model path bit for bit."""
    from lightfm_b200 import LightFM, _lightfm_fast as fast
    inter = H.synthetic_interactions(3000, 2000, 40000, 9)
    model = LightFM(loss="bpr", no_components=24, random_state=3).fit(inter, epochs=1, num_threads=4)
    rng = np.random.default_rng(0)
    u = rng.integers(0, 3000, 200).astype(np.int32)
    i = rng.integers(0, 2000, 200).astype(np.int32)
    assert model._predict_compact(u, i, 1) is not None
    got = model.predict(u, i)
    full = np.empty(200, np.float32)
    fast.predict_lightfm(fast.CSRMatrix(sp.identity(2000, dtype=np.float32, format="csr")),
                         fast.CSRMatrix(sp.identity(3000, dtype=np.float32, format="csr")),
                         u, i, full, model._get_lightfm_data(), 1)
    assert np.array_equal(got, full)
    assert np.array_equal(model.predict(7, np.arange(50, dtype=np.int32)),
                          model.predict(np.repeat(7, 50).astype(np.int32), np.arange(50, dtype=np.int32)))
