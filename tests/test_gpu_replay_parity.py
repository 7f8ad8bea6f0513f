"""GPU: replay mode (num_threads=1) against the golden vectors and the CPU oracle.

Bar (BASELINE.json north_star): weights within 1e-5 relative => held-out precision@k within
1e-4 relative.  Replay mode is designed to be bit-identical; the tests assert bit equality
where libm does not enter (adagrad + warp/kos: the log terms are host-precomputed) and
<= 1e-6 relative elsewhere (double exp() in the sigmoid may differ in the last ulp)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star tolerance for replay-mode weights


def _compare(out, ref, keys, exact):
    for k in keys:
        a, b = out[k], ref[k]
        if exact:
            assert np.array_equal(a, b), "%s differs (max rel %.3g)" % (k, H.max_rel_diff(a, b))
        else:
            assert H.max_rel_diff(a, b) <= REL_TOL, "%s max rel %.3g" % (k, H.max_rel_diff(a, b))


@pytest.mark.parametrize("case", H.golden_cases())
def test_cuda_replay_matches_golden(case):
    cu = H.cuda_native()
    out, g = H.run_golden(cu, case, num_threads=1)
    exact = str(g["loss"]) in ("warp", "warp-kos")
    _compare(out, g, ["final_" + k for k in H.MODEL_ARRAYS], exact)
    # scoring kernels are compiled without FMA contraction: bit-exact predictions and ranks
    assert np.array_equal(out["pred"], g["pred"])
    assert np.array_equal(out["ranks"], g["ranks"])
    assert np.array_equal(out["ranks_sorted"], g["ranks_sorted"])
    assert np.array_equal(out["auc"], g["auc"])


@pytest.mark.parametrize("loss", ("logistic", "warp", "bpr", "warp-kos"))
@pytest.mark.parametrize("schedule", ("adagrad", "adadelta"))
@pytest.mark.parametrize("feats,alpha,d", [(False, 0.0, 16), (True, 1e-3, 24), (False, 1e-4, 10)])
def test_cuda_replay_matches_oracle(loss, schedule, feats, alpha, d):
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(150, 110, 3000, 11, signed=(loss == "logistic"))
    itf = H.tag_features(110, 20, 4, 5) if feats else None
    usf = H.tag_features(150, 15, 3, 6) if feats else None
    hp = H.Hyper(d=d, schedule=schedule, item_alpha=alpha, user_alpha=alpha)
    nif = itf.shape[1] if feats else 110
    nuf = usf.shape[1] if feats else 150
    outs = []
    for api in (orc, cu):
        rs = np.random.RandomState(5)
        arr = H.init_arrays(rs, nif, nuf, d, schedule)
        for _ in range(2):
            H.run_epoch(api, loss, inter, arr, hp, rs, itf, usf, num_threads=1)
        outs.append(arr)
    exact = loss in ("warp", "warp-kos")
    _compare(outs[1], outs[0], H.MODEL_ARRAYS, exact)


def test_config1_shape_bpr_replay_precision_parity():
    """BASELINE config 1 (ML-100k shape, BPR, d=16, 1 thread): held-out precision@10 of the
    CUDA replay fit equals the oracle's within 1e-4 relative."""
    import scipy.sparse as sp
    cu, orc = H.cuda_native(), H.oracle_native()
    full = H.synthetic_interactions(943, 1682, 100000, 1)
    rs_split = np.random.RandomState(7)
    order = np.arange(full.nnz)
    rs_split.shuffle(order)
    cut = int(0.8 * full.nnz)
    tr, te = order[:cut], order[cut:]
    train = sp.coo_matrix((full.data[tr], (full.row[tr], full.col[tr])), shape=full.shape)
    test = sp.coo_matrix((full.data[te], (full.row[te], full.col[te])), shape=full.shape).tocsr()
    test.sort_indices()
    hp = H.Hyper(d=16)
    res = []
    for api in (orc, cu):
        rs = np.random.RandomState(1)
        arr = H.init_arrays(rs, 1682, 943, 16)
        H.run_epoch(api, "bpr", train, arr, hp, rs, num_threads=1)
        ident_i = sp.identity(1682, dtype=np.float32, format="csr")
        ident_u = sp.identity(943, dtype=np.float32, format="csr")
        ranks = np.zeros_like(test.data)
        trc = train.tocsr().astype(np.float32)
        trc.sort_indices()
        api.predict_ranks(api.CSRMatrix(ident_i), api.CSRMatrix(ident_u), api.CSRMatrix(test),
                          api.CSRMatrix(trc), ranks, H.holder(api, arr, hp), 1)
        hits = sp.csr_matrix((ranks < 10, test.indices, test.indptr), shape=test.shape)
        p = np.asarray(hits.sum(axis=1)).ravel() / 10.0
        res.append((arr, p[test.getnnz(axis=1) > 0].mean()))
    for k in H.MODEL_ARRAYS:
        assert H.max_rel_diff(res[1][0][k], res[0][0][k]) <= REL_TOL, k
    assert abs(res[1][1] - res[0][1]) <= 1e-4 * max(abs(res[0][1]), 1e-12)


def test_replay_is_bit_reproducible():  # reference tests/test_movielens.py:655-666
    from lightfm_b200 import LightFM
    inter = H.synthetic_interactions(200, 150, 4000, 3)
    a = LightFM(loss="warp", no_components=12, random_state=10).fit(inter, epochs=2, num_threads=1)
    b = LightFM(loss="warp", no_components=12, random_state=10).fit(inter, epochs=2, num_threads=1)
    assert np.array_equal(a.item_embeddings, b.item_embeddings)
    assert np.array_equal(a.user_embeddings, b.user_embeddings)


def test_regularize_trigger_mid_epoch_matches_oracle():
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(80, 60, 3000, 3)
    hp = H.Hyper(d=8, item_alpha=2.0, user_alpha=2.0, lr=0.5)
    outs = []
    for api in (orc, cu):
        rs = np.random.RandomState(3)
        arr = H.init_arrays(rs, 60, 80, 8)
        H.run_epoch(api, "warp", inter, arr, hp, rs, num_threads=1)
        outs.append(arr)
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k


# ---- SURVEY 8(c) tier A1 at the BASELINE widths: slices of C2 / C3 / C5 ----------------------
# d = 64 and d = 128 take the KMAX register-prefetch path of warp_update (k = 2 and 4 components
# per lane); C3's [I | tags] item features at d = 128 take the general CSR path.

def _slice_fit(api, loss, inter, d, itf=None, sw=None, epochs=1, **hpkw):
    hp = H.Hyper(d=d, **hpkw)
    rs = np.random.RandomState(21)
    nif = itf.shape[1] if itf is not None else inter.shape[1]
    arr = H.init_arrays(rs, nif, inter.shape[0], d)
    for _ in range(epochs):
        H.run_epoch(api, loss, inter, arr, hp, rs, item_features=itf, sample_weight=sw, num_threads=1)
    return arr


@pytest.mark.parametrize("d", (64, 128))
def test_c2_slice_warp_replay_bit_equal(d):
    """C2 slice: 60 k interactions of the 138 493 x 26 744 problem's generator at 1/10 scale."""
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(13_849, 2_674, 60_000, 2)
    a, b = _slice_fit(orc, "warp", inter, d), _slice_fit(cu, "warp", inter, d)
    _compare(b, a, H.MODEL_ARRAYS, exact=True)


def test_c2_slice_kos_replay_bit_equal():
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(13_849, 2_674, 40_000, 4)
    a, b = _slice_fit(orc, "warp-kos", inter, 64), _slice_fit(cu, "warp-kos", inter, 64)
    _compare(b, a, H.MODEL_ARRAYS, exact=True)


def test_c3_slice_tag_features_d128_replay_bit_equal():
    """C3 slice: item features = [I | 1000 tags] (8 Zipf tags per item, rows L1-normalised), d=128."""
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(6_000, 2_674, 30_000, 3)
    itf = H.tag_features(2_674, 1000, 8, 3)
    a, b = _slice_fit(orc, "warp", inter, 128, itf=itf), _slice_fit(cu, "warp", inter, 128, itf=itf)
    _compare(b, a, H.MODEL_ARRAYS, exact=True)


def test_c5_slice_weighted_logistic_d32_replay():
    """C5 slice: explicit +-1 feedback with sample_weight ~ U(0.5, 1.5), logistic, d=32."""
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(20_000, 2_000, 80_000, 5, signed=True)
    sw = (0.5 + np.random.default_rng(5).random(inter.nnz)).astype(np.float32)
    a = _slice_fit(orc, "logistic", inter, 32, sw=sw)
    b = _slice_fit(cu, "logistic", inter, 32, sw=sw)
    _compare(b, a, H.MODEL_ARRAYS, exact=False)   # device exp() in the sigmoid: <= 1e-5 relative
