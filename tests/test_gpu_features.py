"""GPU: the feature path of the throughput kernels (shared feature rows: BASELINE config 3's
shape at small scale).  Statistical parity with the oracle for
  * the hot-row variant (per-CTA shared-memory aggregation of the most-touched tag rows),
  * the same kernels with the aggregation switched off (direct global reductions),
  * features + L2 regularisation (generic kernel, log-domain lazy scale)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
SEEDS = (0, 1, 2)


def _problem():
    full = H.planted_clusters(3000, 1500, 150_000, seed=4, n_clusters=12)
    train, test = H.split(full, 5)
    itf = H.tag_features(1500, 200, 8, seed=3)   # [I | 200 Zipf tags], rows L1-normalised
    users = np.arange(0, 3000, 3)
    return train, test, itf, users


def _fit(api, train, itf, d, seed, nt, epochs=5, **hpkw):
    hp = H.Hyper(d=d, **hpkw)
    rs = np.random.RandomState(seed)
    arr = H.init_arrays(rs, itf.shape[1], train.shape[0], d)
    for _ in range(epochs):
        H.run_epoch(api, "warp", train, arr, hp, rs, item_features=itf, num_threads=nt)
    return arr


def _eval(arr, itf, train, test, users):
    # project the feature embeddings to item representations, then the identity-feature evaluator
    rep = {"item_embeddings": np.asarray(itf @ arr["item_embeddings"]),
           "item_biases": np.asarray(itf @ arr["item_biases"]).ravel(),
           "user_embeddings": arr["user_embeddings"], "user_biases": arr["user_biases"]}
    return H.eval_subset(rep, train, test, users)


def test_hot_row_aggregation_is_used_and_matches_oracle_statistically():
    cu, orc = H.cuda_native(), H.oracle_native()
    train, test, itf, users = _problem()
    res = {}
    for label, api, nt, hot in (("oracle", orc, 1, None), ("hot", cu, 8, True), ("direct", cu, 8, False)):
        if hot is not None:
            cu.module.set_hot_rows(hot)
        try:
            res[label] = np.array([_eval(_fit(api, train, itf, 32, s, nt), itf, train, test, users)
                                   for s in SEEDS]).mean(axis=0)
        finally:
            cu.module.set_hot_rows(True)
    print("feature path p@10 / auc:", {k: v.round(4).tolist() for k, v in res.items()})
    assert res["oracle"][1] > 0.7
    for label in ("hot", "direct"):
        assert abs(res[label][1] - res["oracle"][1]) < 0.015, res
        assert abs(res[label][0] - res["oracle"][0]) <= 0.08 * res["oracle"][0] + 0.005, res


def test_hot_rows_conserve_updates():
    """Every update reaches the tables exactly once whether it went through shared memory or
    not: with lr = 0 the weights stay put and the accumulators grow by the same total either way
    (same seed => same draws: the accumulator sums agree to fp32 summation noise)."""
    cu = H.cuda_native()
    train, _, itf, _ = _problem()
    sums = []
    for hot in (True, False):
        cu.module.set_hot_rows(hot)
        try:
            arr = _fit(cu, train, itf, 32, 3, 8, epochs=1, lr=0.0)
        finally:
            cu.module.set_hot_rows(True)
        init = H.init_arrays(np.random.RandomState(3), itf.shape[1], train.shape[0], 32)
        assert np.array_equal(arr["item_embeddings"], init["item_embeddings"])
        sums.append((arr["item_embedding_gradients"].astype(np.float64).sum(),
                     arr["item_bias_gradients"].astype(np.float64).sum(),
                     arr["user_embedding_gradients"].astype(np.float64).sum()))
    for a, b in zip(*sums):
        assert abs(a - b) <= 1e-4 * abs(b), sums


def test_features_with_l2_match_oracle_statistically():
    cu, orc = H.cuda_native(), H.oracle_native()
    train, test, itf, users = _problem()
    res = {}
    for label, api, nt in (("oracle", orc, 1), ("gpu", cu, 8)):
        res[label] = np.array([_eval(_fit(api, train, itf, 32, s, nt, item_alpha=1e-5, user_alpha=1e-5),
                                     itf, train, test, users) for s in SEEDS]).mean(axis=0)
    print("features + L2 p@10 / auc:", {k: v.round(4).tolist() for k, v in res.items()})
    assert abs(res["gpu"][1] - res["oracle"][1]) < 0.015, res
    assert abs(res["gpu"][0] - res["oracle"][0]) <= 0.08 * res["oracle"][0] + 0.005, res
