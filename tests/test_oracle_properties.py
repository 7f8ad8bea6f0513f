"""CPU: behavioural properties of the reference's algorithm, checked on the oracle (which is
bit-pinned to the real reference by test_oracle_golden.py).  They restate invariants from the
reference's tests/test_movielens.py and tests/test_evaluation.py that do not need MovieLens."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H


def _fit(loss, inter, d=8, epochs=2, seed=0, sample_weight=None, **hpkw):
    orc = H.oracle_native()
    hp = H.Hyper(d=d, **hpkw)
    rs = np.random.RandomState(seed)
    arr = H.init_arrays(rs, inter.shape[1], inter.shape[0], d, hp.schedule)
    for _ in range(epochs):
        H.run_epoch(orc, loss, inter, arr, hp, rs, sample_weight=sample_weight)
    return arr, hp


@pytest.mark.parametrize("loss", ("warp", "bpr", "logistic"))
def test_zero_weights_leave_accumulators_exactly_one(loss):  # tests/test_movielens.py:437-460
    inter = H.synthetic_interactions(60, 50, 800, 4)
    arr, _ = _fit(loss, inter, sample_weight=np.zeros(inter.nnz, np.float32))
    for k in ("item_embedding_gradients", "item_bias_gradients", "user_embedding_gradients", "user_bias_gradients"):
        assert np.all(arr[k] == 1.0), k


def test_max_sampled_zero_is_a_noop():  # tests/test_movielens.py:247-263
    inter = H.synthetic_interactions(60, 50, 800, 4)
    before = H.init_arrays(np.random.RandomState(0), 50, 60, 8)
    arr, _ = _fit("warp", inter, max_sampled=0, epochs=1)
    assert np.array_equal(arr["item_embeddings"], before["item_embeddings"])
    assert np.array_equal(arr["user_embeddings"], before["user_embeddings"])


@pytest.mark.parametrize("schedule", ("adagrad", "adadelta"))
def test_schedule_state(schedule):  # tests/test_movielens.py:602-652
    inter = H.synthetic_interactions(60, 50, 800, 4)
    arr, _ = _fit("warp", inter, schedule=schedule)
    if schedule == "adagrad":
        assert np.all(arr["item_embedding_gradients"] >= 1) and np.any(arr["item_embedding_gradients"] > 1)
        assert not arr["item_embedding_momentum"].any() and not arr["user_bias_momentum"].any()
    else:
        assert np.all(arr["item_embedding_gradients"] >= 0) and np.any(arr["item_embedding_gradients"] > 0)
        assert np.all(arr["item_embedding_momentum"] >= 0) and np.any(arr["item_embedding_momentum"] > 0)


def test_same_seed_same_weights_different_seed_different_weights():  # tests/test_movielens.py:655-666
    inter = H.synthetic_interactions(60, 50, 800, 4)
    a, _ = _fit("warp", inter, seed=1)
    b, _ = _fit("warp", inter, seed=1)
    c, _ = _fit("warp", inter, seed=2)
    assert np.array_equal(a["item_embeddings"], b["item_embeddings"])
    assert not np.array_equal(a["item_embeddings"], c["item_embeddings"])


def test_predict_rank_structure_and_metrics_against_brute_force():
    """Ranks from predict_ranks equal a brute-force count built on predict_lightfm, train positives
    excluded, ties pessimistic (tests/test_api.py:217-282, tests/test_evaluation.py:34-161); AUC from
    calculate_auc_from_rank equals the pairwise definition."""
    orc = H.oracle_native()
    full = H.planted_interactions(40, 30, 8, seed=3)
    train, test = H.split(full, 1)
    arr, hp = _fit("warp", train, d=8, epochs=3)
    ident_i = orc.CSRMatrix(sp.identity(30, dtype=np.float32, format="csr"))
    ident_u = orc.CSRMatrix(sp.identity(40, dtype=np.float32, format="csr"))
    tc, trc = test.tocsr().astype(np.float32), train.tocsr().astype(np.float32)
    tc.sort_indices(); trc.sort_indices()
    ranks = np.zeros_like(tc.data)
    h = H.holder(orc, arr, hp)
    orc.predict_ranks(ident_i, ident_u, orc.CSRMatrix(tc), orc.CSRMatrix(trc), ranks, h, 1)
    u = np.repeat(np.arange(40), 30).astype(np.int32)
    i = np.tile(np.arange(30), 40).astype(np.int32)
    scores = np.empty(1200, np.float32)
    orc.predict_lightfm(ident_i, ident_u, u, i, scores, h, 1)
    scores = scores.reshape(40, 30)
    auc = np.zeros(40, np.float32)
    rk = sp.csr_matrix((ranks.copy(), tc.indices, tc.indptr), shape=tc.shape)
    orc.calculate_auc_from_rank(orc.CSRMatrix(rk), np.asarray(trc.getnnz(axis=1)).astype(np.int32), rk.data, auc, 1)
    for user in range(40):
        te, tr = tc[user].indices, set(trc[user].indices)
        for t in te:
            want = sum(1 for it in range(30) if it not in tr and it != t and scores[user, it] >= scores[user, t])
            got = ranks[tc.indptr[user] + list(te).index(t)]
            assert got == want
        if len(te) == 0:
            assert auc[user] == 0.5
            continue
        neg = [it for it in range(30) if it not in tr and it not in set(te)]
        if not neg:
            continue
        # pairwise AUC (no score ties among these random floats)
        wins = sum(scores[user, t] > scores[user, n] for t in te for n in neg)
        assert abs(auc[user] - wins / (len(te) * len(neg))) < 1e-5


def test_all_equal_scores_give_pessimistic_ranks():  # tests/test_api.py:258-266
    orc = H.oracle_native()
    arr = H.init_arrays(np.random.RandomState(0), 20, 10, 4)
    for k in ("item_embeddings", "user_embeddings"):
        arr[k][:] = 0
    dense = sp.csr_matrix(np.ones((10, 20), np.float32))
    empty = sp.csr_matrix((10, 20), dtype=np.float32)
    ranks = np.zeros_like(dense.data)
    ident_i = orc.CSRMatrix(sp.identity(20, dtype=np.float32, format="csr"))
    ident_u = orc.CSRMatrix(sp.identity(10, dtype=np.float32, format="csr"))
    orc.predict_ranks(ident_i, ident_u, orc.CSRMatrix(dense), orc.CSRMatrix(empty), ranks,
                      H.holder(orc, arr, H.Hyper(d=4)), 1)
    assert np.all(ranks == 19)


def test_synthetic_generator_is_deterministic_and_deduplicated():
    from lightfm_b200 import synthetic
    a = synthetic.interactions(300, 200, 5000, seed=3)
    b = synthetic.interactions(300, 200, 5000, seed=3)
    assert np.array_equal(a.row, b.row) and np.array_equal(a.col, b.col)
    assert a.nnz == 5000 and a.row.dtype == np.int32 and a.data.dtype == np.float32
    keys = a.row.astype(np.int64) * 200 + a.col
    assert len(np.unique(keys)) == 5000
    f = synthetic.tag_features(200, 30, 4, seed=1)
    assert f.shape == (200, 230) and np.allclose(np.asarray(f.sum(axis=1)).ravel(), 1.0, atol=1e-6)
