"""GPU: public-API behaviour of LightFM end to end (restates the offline-capable cases of the
reference's tests/test_api.py and the state/RNG invariants of tests/test_movielens.py)."""
import pickle

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from lightfm_b200 import LightFM

pytestmark = pytest.mark.gpu


def test_empty_matrix():  # tests/test_api.py:10-17
    LightFM().fit(sp.coo_matrix((20, 30), dtype=np.int32))


@pytest.mark.parametrize("fmt", ("coo", "csr", "csc", "lil", "dok"))
@pytest.mark.parametrize("dtype", (np.int32, np.int64, np.float32, np.float64))
def test_matrix_types(fmt, dtype):  # tests/test_api.py:20-54
    rng = np.random.RandomState(0)
    dense = (rng.rand(12, 9) > 0.6).astype(dtype)
    train = getattr(sp, fmt + "_matrix")(dense)
    uf = getattr(sp, fmt + "_matrix")(np.eye(12, dtype=dtype))
    itf = getattr(sp, fmt + "_matrix")(np.eye(9, dtype=dtype))
    w = sp.coo_matrix(train, dtype=dtype)
    model = LightFM(no_components=4)
    model.fit(train, user_features=uf, item_features=itf, sample_weight=w if dense.any() else None)
    model.predict(np.arange(5, dtype=np.int32), np.arange(5, dtype=np.int32),
                  user_features=uf, item_features=itf)
    model.predict_rank(train, user_features=uf, item_features=itf, check_intersections=False)


@pytest.mark.parametrize("loss", ("warp", "bpr", "warp-kos"))
def test_coo_with_duplicate_entries(loss):  # tests/test_api.py:57-74
    rng = np.random.RandomState(0)
    rows = rng.randint(0, 40, 300).astype(np.int32)
    cols = rng.randint(0, 30, 300).astype(np.int32)
    rows = np.concatenate([rows, rows[:100]])  # duplicate COO entries (lightfm issue #117)
    cols = np.concatenate([cols, cols[:100]])
    mat = sp.coo_matrix((np.ones(400, np.float32), (rows, cols)), shape=(40, 30))
    for nt in (1, 2):
        LightFM(loss=loss, no_components=4).fit(mat, epochs=2, num_threads=nt)


def test_return_self_and_feature_shape_errors():  # tests/test_api.py:121-168
    train = sp.coo_matrix((np.ones(3, np.float32), ([0, 1, 2], [0, 1, 2])), shape=(3, 3))
    model = LightFM(no_components=4)
    assert model.fit_partial(train) is model
    assert model.fit(train) is model
    with pytest.raises(ValueError):
        model.fit_partial(train, item_features=sp.csr_matrix(np.ones((3, 5), np.float32)))
    with pytest.raises(ValueError):
        model.predict(np.arange(3), np.arange(3), user_features=sp.csr_matrix(np.ones((3, 7), np.float32)))


def test_overflow_divergence_raises():  # tests/test_api.py:285-294
    """Divergence is detected after every epoch and raises ValueError (L:447-464, L:664)."""
    rng = np.random.RandomState(0)
    train = sp.coo_matrix((rng.rand(40, 30) > 0.5).astype(np.float32) * 1e9)
    feats = sp.csr_matrix(rng.rand(30, 10).astype(np.float32) * 1e9)
    # replay mode is deterministic: this configuration overflows within five epochs
    with pytest.raises(ValueError):
        LightFM(loss="logistic", learning_rate=1e6, no_components=4, random_state=0).fit(
            train, item_features=feats, epochs=5, num_threads=1)
    # both modes: a non-finite parameter is caught by the per-epoch check (on the device in the
    # resident path) whatever the training dynamics are
    ident = H.synthetic_interactions(80, 60, 1500, 2)
    for nt in (1, 4):
        for loss in ("warp", "logistic"):
            model = LightFM(loss=loss, no_components=16, random_state=0).fit(ident, epochs=1, num_threads=nt)
            model.item_embeddings[3, 2] = np.inf
            with pytest.raises(ValueError):
                model.fit_partial(ident, epochs=2, num_threads=nt)


def test_warp_few_items_stays_finite():  # tests/test_api.py:374-382
    train = sp.coo_matrix((np.ones(4, np.float32), ([0, 1, 2, 3], [0, 1, 0, 1])), shape=(4, 2))
    for nt in (1, 4):
        model = LightFM(loss="warp", max_sampled=10, no_components=4).fit(train, epochs=3, num_threads=nt)
        assert np.isfinite(model.item_embeddings).all() and np.isfinite(model.user_embeddings).all()


def test_state_reset_resume_and_pickle():  # tests/test_movielens.py:387-412,463-472
    train = H.synthetic_interactions(80, 60, 1500, 2)
    model = LightFM(loss="warp", no_components=8, random_state=3)
    model.fit(train, epochs=2)
    again = pickle.loads(pickle.dumps(model))
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(getattr(model, k), getattr(again, k))
    # resume: 1 + 1 epochs via fit_partial == 2 epochs in one call (same RNG stream)
    a = LightFM(loss="warp", no_components=8, random_state=3)
    a.fit_partial(train, epochs=1)
    a.fit_partial(train, epochs=1)
    assert np.array_equal(a.item_embeddings, model.item_embeddings)
    # fit() discards the previous state
    model.fit(train, epochs=0)
    assert np.all(model.item_bias_gradients == 1)


def test_random_state_advances_each_epoch():  # tests/test_movielens.py:669-682
    train = H.synthetic_interactions(80, 60, 1500, 2)
    for nt in (1, 4):
        model = LightFM(loss="warp", no_components=8, random_state=3)
        states = []
        for _ in range(3):
            model.fit_partial(train, epochs=1, num_threads=nt)
            states.append(model.random_state.get_state()[1].copy())
        assert not np.array_equal(states[0], states[1])
        assert not np.array_equal(states[1], states[2])


def test_user_mutation_between_calls_is_respected():
    train = H.synthetic_interactions(80, 60, 1500, 2)
    model = LightFM(loss="warp", no_components=8, random_state=3).fit(train, epochs=1)
    model.item_biases *= 0.0
    model.item_biases += 5.0
    p = model.predict(0, np.arange(60, dtype=np.int32))
    emb = model.user_embeddings[0] @ model.item_embeddings.T + model.user_biases[0] + 5.0
    assert np.allclose(p, emb, atol=1e-5)


def test_sharded_trainer_single_rank_matches_plain_fit_statistically():
    """ShardedTrainer with one rank (no process group) is the plain resident path: it must learn
    the planted structure as well as LightFM.fit does, and gather() must fill the model arrays."""
    from lightfm_b200 import sharding
    full = H.planted_interactions(400, 300, 30, seed=5)
    train, test = H.split(full, 7)
    model = LightFM(loss="warp", no_components=32, random_state=1)
    trainer = sharding.ShardedTrainer(model, train, axis="item")
    assert trainer.world == 1 and trainer.local_interactions == train.nnz
    trainer.fit_epochs(8)
    trainer.gather()
    trainer.close()
    arrays = {k: getattr(model, k) for k in H.MODEL_ARRAYS}
    p, auc = H.eval_arrays(arrays, 32, train, test)
    plain = LightFM(loss="warp", no_components=32, random_state=1).fit(train, epochs=8, num_threads=8)
    p2, auc2 = H.eval_arrays({k: getattr(plain, k) for k in H.MODEL_ARRAYS}, 32, train, test)
    assert auc > 0.65 and abs(auc - auc2) < 0.03, (auc, auc2)
    assert np.all(model.item_embedding_gradients >= 1) and np.any(model.user_embedding_gradients > 1)


def test_csr_free_resident_fit_learns_and_matches_csr_path():
    """WARP / BPR with identity features on the bitmap fast path never build the positives CSR on
    the host (the library builds a membership bitmap from the COO arrays).  The fit must learn the
    planted structure as well as the CSR-backed path (bitmap disabled -> sorted-row search)."""
    from lightfm_b200 import _lightfm_fast as fast
    full = H.planted_interactions(400, 300, 30, seed=5)
    train, test = H.split(full, 7)
    res = {}
    for name, limit in (("bitmap", 1 << 30), ("csr", 0)):
        fast.set_bitmap_limit(limit)
        try:
            for loss, d in (("warp", 64), ("bpr", 16)):
                m = LightFM(loss=loss, no_components=d, random_state=1).fit(train, epochs=8, num_threads=8)
                res[(name, loss)] = H.eval_arrays({k: getattr(m, k) for k in H.MODEL_ARRAYS}, d, train, test)[1]
        finally:
            fast.set_bitmap_limit(1 << 30)
    for loss in ("warp", "bpr"):
        assert res[("bitmap", loss)] > 0.75 and abs(res[("bitmap", loss)] - res[("csr", loss)]) < 0.03, res


def test_plan_without_csr_is_refused_off_the_fast_path():
    from lightfm_b200 import _lightfm_fast as fast
    inter = H.synthetic_interactions(60, 50, 500, 1)
    arrays = H.init_arrays(np.random.RandomState(0), 50, 60, 10)   # d = 10: not a fast-path size
    holder = H.holder(fast, arrays, H.Hyper(d=10))
    ident = lambda n: fast.CSRMatrix(sp.identity(n, dtype=np.float32, format="csr"))
    with pytest.raises(ValueError):
        fast.ResidentPlan("warp", ident(50), ident(60), None, inter.row, inter.col, inter.data, inter.data,
                          holder, 0.0, 0.0)
