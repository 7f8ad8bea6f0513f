"""CPU: the C oracle (oracle/lfm_oracle.c) against the golden vectors generated from the
real reference, and -- where oracle/_ref is present -- against the reference itself."""
import numpy as np
import pytest

import helpers as H


@pytest.mark.parametrize("case", H.golden_cases())
def test_oracle_matches_golden_bit_for_bit(case):
    out, g = H.run_golden(H.oracle_native(), case)
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(out["final_" + k], g["final_" + k]), k
    assert np.array_equal(out["pred"], g["pred"])
    assert np.array_equal(out["ranks"], g["ranks"])
    assert np.array_equal(out["ranks_sorted"], g["ranks_sorted"])
    assert np.array_equal(out["auc"], g["auc"])


def test_in_positives_known_answers():
    # restates reference tests/test_fast_functions.py:9-17
    import scipy.sparse as sp
    orc = H.oracle_native()
    mat = sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=np.float32))
    c = orc.CSRMatrix(mat)
    assert not orc.test_in_positives(0, 0, c)
    assert orc.test_in_positives(0, 1, c)
    assert orc.test_in_positives(1, 0, c)
    assert not orc.test_in_positives(1, 1, c)


def _ref_or_skip():
    import oracle
    if not oracle.reference_available("strict"):
        pytest.skip("oracle/_ref not built (needs /root/reference): golden vectors cover this")
    return H.reference_native("strict")


LOSSES = ("logistic", "warp", "bpr", "warp-kos")


@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("schedule", ("adagrad", "adadelta"))
@pytest.mark.parametrize("feats,alpha", [(False, 0.0), (True, 1e-3)])
def test_oracle_bit_equal_to_reference_build(loss, schedule, feats, alpha):
    ref = _ref_or_skip()
    orc = H.oracle_native()
    inter = H.synthetic_interactions(120, 90, 2500, 1, signed=(loss == "logistic"))
    itf = H.tag_features(90, 20, 4, 5) if feats else None
    usf = H.tag_features(120, 15, 3, 6) if feats else None
    d = 16
    hp = H.Hyper(d=d, schedule=schedule, item_alpha=alpha, user_alpha=alpha)
    nif = itf.shape[1] if feats else 90
    nuf = usf.shape[1] if feats else 120
    outs = []
    for api in (ref, orc):
        rs = np.random.RandomState(42)
        arr = H.init_arrays(rs, nif, nuf, d, schedule)
        for _ in range(2):
            H.run_epoch(api, loss, inter, arr, hp, rs, itf, usf)
        outs.append(arr)
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_oracle_regularize_trigger_matches_reference():
    # large alpha drives item_scale past 1e6 mid-epoch (T:901-904, locked_regularize)
    ref = _ref_or_skip()
    orc = H.oracle_native()
    inter = H.synthetic_interactions(80, 60, 3000, 3)
    hp = H.Hyper(d=8, item_alpha=2.0, user_alpha=2.0, lr=0.5)
    outs = []
    for api in (ref, orc):
        rs = np.random.RandomState(3)
        arr = H.init_arrays(rs, 60, 80, 8)
        H.run_epoch(api, "warp", inter, arr, hp, rs)
        outs.append(arr)
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k


# ---- edge cases of the reference's semantics (SURVEY 8(a) "semantics traps"), oracle vs reference ----
def _both(loss, inter, d=8, epochs=2, sw=None, **hpkw):
    ref = _ref_or_skip()
    orc = H.oracle_native()
    outs = []
    for api in (ref, orc):
        hp = H.Hyper(d=d, **hpkw)
        rs = np.random.RandomState(9)
        arr = H.init_arrays(rs, inter.shape[1], inter.shape[0], d, hp.schedule)
        for _ in range(epochs):
            H.run_epoch(api, loss, inter, arr, hp, rs, sample_weight=sw)
        outs.append(arr)
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k
    return outs[1]


@pytest.mark.parametrize("loss", ("warp", "bpr"))
def test_nonpositive_rows_are_skipped_but_still_reject_negatives(loss):
    # Y <= 0 entries are not trained on (T:831, T:1116) yet stay in the positives lookup
    inter = H.synthetic_interactions(50, 40, 600, 2, signed=True)
    _both(loss, inter)


@pytest.mark.parametrize("loss", ("warp", "bpr", "logistic"))
def test_sample_weights(loss):
    inter = H.synthetic_interactions(50, 40, 600, 2, signed=(loss == "logistic"))
    sw = (0.25 + np.random.default_rng(1).random(inter.nnz) * 3).astype(np.float32)
    _both(loss, inter, sw=sw)


@pytest.mark.parametrize("loss", ("warp", "bpr", "warp-kos"))
def test_duplicate_coo_entries(loss):  # lightfm issue #117 (tests/test_api.py:57-74)
    import scipy.sparse as sp
    base = H.synthetic_interactions(40, 30, 300, 3)
    rows = np.concatenate([base.row, base.row[:80]])
    cols = np.concatenate([base.col, base.col[:80]])
    inter = sp.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=base.shape)
    _both(loss, inter)


def test_kos_k_larger_than_user_history_and_n_larger_than_history():
    inter = H.synthetic_interactions(60, 50, 150, 5)   # most users have 1-4 positives
    _both("warp-kos", inter, k=7, n=12)


def test_max_sampled_one_and_many():
    inter = H.synthetic_interactions(50, 40, 600, 2)
    _both("warp", inter, max_sampled=1)
    _both("warp", inter, max_sampled=37)


def test_tiny_catalogue_log_of_zero_in_kos_propagates_identically():
    # T:1039 has no max(1, .) guard: floor((n_items-1)/sampled) == 0 -> log(0) = -inf -> NaNs
    inter = H.synthetic_interactions(20, 3, 40, 1)
    out = _both("warp-kos", inter, epochs=1, k=2, n=3)
    assert not np.isfinite(out["item_embeddings"]).all()


def test_warp_two_items_stays_finite():  # tests/test_api.py:374-382
    inter = H.synthetic_interactions(20, 2, 25, 1)
    out = _both("warp", inter, epochs=3)
    assert np.isfinite(out["item_embeddings"]).all() and np.isfinite(out["user_embeddings"]).all()


def test_odd_component_counts_and_adadelta_eps_zero():
    inter = H.synthetic_interactions(50, 40, 600, 2)
    _both("warp", inter, d=1)
    _both("bpr", inter, d=33, schedule="adadelta", eps=0.0)
