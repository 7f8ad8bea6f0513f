"""CPU: host-side behaviour of LightFM that must hold before any kernel runs
(restates the validation cases of the reference's tests/test_api.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from lightfm_b200 import LightFM
from lightfm_b200.cross_validation import random_train_test_split


def test_constructor_assertions():  # reference tests/test_api.py:171-183
    for kw in (dict(no_components=-1), dict(user_alpha=-1.0), dict(item_alpha=-1.0)):
        with pytest.raises(AssertionError):
            LightFM(**kw)
    with pytest.raises(ValueError):
        LightFM(max_sampled=-1)
    with pytest.raises(ValueError):
        LightFM(max_sampled=0)
    with pytest.raises(AssertionError):
        LightFM(loss="hinge")
    with pytest.raises(AssertionError):
        LightFM(learning_schedule="sgd")


def test_get_set_params_roundtrip():  # tests/test_api.py:297-306
    model = LightFM(loss="warp", no_components=7)
    params = model.get_params()
    assert params["no_components"] == 7 and params["loss"] == "warp"
    clone = LightFM(**params)
    assert clone.get_params() == params
    model.set_params(no_components=3)
    assert model.no_components == 3
    with pytest.raises(ValueError):
        model.set_params(bogus=1)


def test_not_fitted_errors():  # tests/test_api.py:309-323
    model = LightFM()
    with pytest.raises(ValueError):
        model.predict(np.arange(3), np.arange(3))
    with pytest.raises(ValueError):
        model.predict_rank(sp.identity(3, format="csr"))
    with pytest.raises(ValueError):
        model.get_user_representations()
    with pytest.raises(ValueError):
        model.get_item_representations()


def test_sample_weight_validation():  # tests/test_api.py:186-214
    model = LightFM(loss="warp-kos")
    train = sp.coo_matrix((np.ones(3, np.float32), ([0, 1, 2], [0, 1, 2])), shape=(3, 3))
    with pytest.raises(NotImplementedError):
        model.fit(train, sample_weight=train.copy())
    model = LightFM(loss="warp")
    with pytest.raises(ValueError):
        model.fit(train, sample_weight=train.tocsr())
    with pytest.raises(ValueError):
        model.fit(train, sample_weight=sp.coo_matrix((4, 4), dtype=np.float32))
    other = sp.coo_matrix((np.ones(3, np.float32), ([2, 1, 0], [0, 1, 2])), shape=(3, 3))
    with pytest.raises(ValueError):
        model.fit(train, sample_weight=other)


def test_nan_inputs_rejected_before_compute():  # tests/test_api.py:326-351
    train = sp.coo_matrix((np.array([1.0, np.nan, 1.0], np.float32), ([0, 1, 2], [0, 1, 2])),
                          shape=(3, 3))
    with pytest.raises(ValueError):
        LightFM().fit(train)
    good = sp.coo_matrix((np.ones(3, np.float32), ([0, 1, 2], [0, 1, 2])), shape=(3, 3))
    feats = sp.csr_matrix(np.array([[np.inf, 0], [0, 1], [1, 0]], dtype=np.float32))
    with pytest.raises(ValueError):
        LightFM().fit(good, item_features=feats)


def test_feature_row_count_errors():  # tests/test_api.py:121-157
    train = sp.coo_matrix((np.ones(3, np.float32), ([0, 1, 2], [0, 1, 2])), shape=(3, 3))
    with pytest.raises(Exception):
        LightFM().fit(train, user_features=sp.identity(2, format="csr", dtype=np.float32))
    with pytest.raises(Exception):
        LightFM().fit(train, item_features=sp.identity(2, format="csr", dtype=np.float32))
    with pytest.raises(ValueError):
        LightFM().fit(train, num_threads=0)


def test_random_train_test_split_is_a_partition():
    rng = np.random.RandomState(0)
    m = sp.random(50, 40, density=0.2, format="coo", random_state=rng, dtype=np.float32)
    train, test = random_train_test_split(m, test_percentage=0.25, random_state=np.random.RandomState(1))
    assert train.shape == test.shape == m.shape
    assert train.nnz + test.nnz == m.nnz
    assert abs(test.nnz / m.nnz - 0.25) < 0.02
    assert train.tocsr().multiply(test.tocsr()).nnz == 0
    with pytest.raises(ValueError):
        random_train_test_split(np.zeros((3, 3)))


# ---- round 2: resident-plan cache keys, tier-B fixtures -------------------------------------------
def test_resident_cache_matches_only_identical_buffers_and_hyperparameters():
    import scipy.sparse as sp
    from lightfm_b200.lightfm import LightFM, _ResidentCache
    inter = H.synthetic_interactions(60, 40, 500, 1)
    model = LightFM(loss="warp", no_components=16)
    cache = _ResidentCache(model, inter, None, None, None)
    cache.plan = object()          # stands in for a live plan (no GPU needed for the key logic)
    assert cache.matches(model, inter, None, None, None)
    # same values, different buffers: not a hit (identity of the arrays is the key)
    clone = sp.coo_matrix((inter.data.copy(), (inter.row.copy(), inter.col.copy())), shape=inter.shape)
    assert not cache.matches(model, clone, None, None, None)
    # a feature matrix or weights appearing: not a hit
    assert not cache.matches(model, inter, sp.identity(60, format="csr", dtype=np.float32), None, None)
    # hyper-parameter change: not a hit
    model.learning_rate = 0.1
    assert not cache.matches(model, inter, None, None, None)
    model.learning_rate = 0.05
    assert cache.matches(model, inter, None, None, None)
    # in-place edit of the interaction buffer is noticed by the strided sample
    inter.col[0] = (inter.col[0] + 1) % 40
    assert not cache.matches(model, inter, None, None, None)
    cache.plan = None


def test_model_pickles_without_device_handles():
    import pickle
    from lightfm_b200.lightfm import LightFM
    model = LightFM(loss="bpr", no_components=8, random_state=3)
    model._initialize(8, 5, 7)
    model.__dict__["_resident_cache"] = object()   # whatever a fit left behind must not be pickled
    clone = pickle.loads(pickle.dumps(model))
    assert clone.__dict__.get("_resident_cache") is None
    assert np.array_equal(clone.item_embeddings, model.item_embeddings)
    model.__dict__["_resident_cache"] = None


def test_tierb_band_fixture_is_consistent_with_the_generator():
    import json
    import os
    path = os.path.join(H.GOLDEN_DIR, "tierb_bands.json")
    bands = json.load(open(path))
    assert set(bands) == set(H.TIERB)
    for name, rec in bands.items():
        cfg = H.TIERB[name]
        assert rec["config"]["epochs"] == cfg["epochs"] and rec["config"].get("lr") == cfg.get("lr"), name
        for side in ("oracle_1thread", "reference_8threads"):
            assert len(rec["runs"][side]) == len(H.TIERB_SEEDS)
        lo, hi = rec["band"]["auc"]
        assert 0.5 < lo <= hi < 1.0
    # the committed digest is the digest of what the generator produces here (C1 shape: fast)
    fit, _, _, _ = H.tierb_problem("c1_warp")
    assert H.data_digest(fit) == bands["c1_warp"]["digest"]
