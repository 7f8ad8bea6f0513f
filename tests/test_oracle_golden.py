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
