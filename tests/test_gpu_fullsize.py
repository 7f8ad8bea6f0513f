"""GPU: BASELINE.json's full sizes (config 2: 138 493 x 26 744, 20 M interactions) through
size-independent properties -- the oracle cannot run these sizes in test time."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

pytestmark = pytest.mark.gpu

N_USERS, N_ITEMS, NNZ = 138_493, 26_744, 20_000_000


@pytest.fixture(scope="module")
def c2():
    import torch
    import bench as B
    rows, cols = B.gen_interactions(N_USERS, N_ITEMS, NNZ, seed=2, device="cuda")
    torch.cuda.empty_cache()
    return rows, cols


def _state(d, zero_embeddings=False):
    rs = np.random.RandomState(0)
    out = []
    for n in (N_ITEMS, N_USERS):
        emb = np.zeros((n, d), np.float32) if zero_embeddings else ((rs.rand(n, d) - 0.5) / d).astype(np.float32)
        out += [emb, np.ones_like(emb), np.zeros_like(emb), np.zeros(n, np.float32),
                np.ones(n, np.float32), np.zeros(n, np.float32)]
    return out


def test_every_interaction_visited_exactly_once_device_permutation(c2):
    """Logistic, lr = 0, zero weights: prediction is 0.5 for every pair, so one epoch adds exactly
    (0.5 w)^2 to the bias accumulators of the pair's user and item.  The accumulators then equal a
    bincount of the inputs iff the device-side permutation visited every interaction once."""
    from lightfm_b200 import _lightfm_fast as fast
    rows, cols = c2
    rng = np.random.default_rng(0)
    w = (0.5 + rng.random(NNZ)).astype(np.float32)
    y = np.ones(NNZ, np.float32)
    st = _state(32, zero_embeddings=True)
    holder = fast.FastLightFM(*st, 32, 0, 0.0, 0.95, 1e-6, 10)
    itf = sp.identity(N_ITEMS, dtype=np.float32, format="csr")
    usf = sp.identity(N_USERS, dtype=np.float32, format="csr")
    plan = fast.ResidentPlan("logistic", fast.CSRMatrix(itf), fast.CSRMatrix(usf), None, rows, cols, y, w,
                             holder, 0.0, 0.0)
    c = plan.epoch(seed=123, num_threads=8)
    plan.download()
    plan.close()
    assert c["positives"] == NNZ and c["updates"] == NNZ
    g2 = (0.5 * w.astype(np.float64)) ** 2
    want_u = 1.0 + np.bincount(rows, weights=g2, minlength=N_USERS)
    want_i = 1.0 + np.bincount(cols, weights=g2, minlength=N_ITEMS)
    assert np.allclose(st[10], want_u, rtol=2e-4), np.abs(st[10] - want_u).max()   # user_bias_gradients
    assert np.allclose(st[4], want_i, rtol=2e-3), np.abs(st[4] / want_i - 1).max()  # item (hot rows: fp32 sums)
    assert np.all(st[0] == 0) and np.all(st[6] == 0)  # lr = 0: embeddings untouched


def test_warp_epoch_conservation_laws(c2):
    """Each WARP update adds loss^2 to one user bias accumulator and to two item bias accumulators,
    and |delta| of the same size to both item biases with opposite sign: size-independent checks
    that no reduction was lost or duplicated among ~18 M concurrent updates."""
    from lightfm_b200 import _lightfm_fast as fast
    rows, cols = c2
    y = np.ones(NNZ, np.float32)
    pos = sp.csr_matrix((y, (rows, cols)), shape=(N_USERS, N_ITEMS))
    pos.sort_indices()
    st = _state(64)
    holder = fast.FastLightFM(*st, 64, 0, 0.05, 0.95, 1e-6, 10)
    itf = sp.identity(N_ITEMS, dtype=np.float32, format="csr")
    usf = sp.identity(N_USERS, dtype=np.float32, format="csr")
    plan = fast.ResidentPlan("warp", fast.CSRMatrix(itf), fast.CSRMatrix(usf), fast.CSRMatrix(pos), rows, cols,
                             y, y, holder, 0.0, 0.0)
    c = plan.epoch(seed=7, num_threads=8)
    assert plan.all_finite()
    plan.download()
    plan.close()
    assert c["positives"] == NNZ
    assert c["updates"] <= NNZ and c["negatives_drawn"] >= c["updates"]
    assert c["negatives_drawn"] <= 10 * NNZ
    gu = np.sum(st[10].astype(np.float64) - 1.0)
    gi = np.sum(st[4].astype(np.float64) - 1.0)
    assert gu > 0 and abs(gi / (2.0 * gu) - 1.0) < 1e-3, (gu, gi)
    for k in (1, 4, 7, 10):  # accumulators only grow
        assert st[k].min() >= 1.0
    for k in (2, 5, 8, 11):  # momentum untouched under adagrad
        assert not st[k].any()


def test_predict_matches_representation_dot_at_full_size(c2):
    from lightfm_b200 import _lightfm_fast as fast
    st = _state(64)
    rng = np.random.default_rng(1)
    st[3][:] = rng.normal(size=N_ITEMS).astype(np.float32)
    st[9][:] = rng.normal(size=N_USERS).astype(np.float32)
    holder = fast.FastLightFM(*st, 64, 0, 0.05, 0.95, 1e-6, 10)
    u = rng.integers(0, N_USERS, 1_000_000).astype(np.int32)
    i = rng.integers(0, N_ITEMS, 1_000_000).astype(np.int32)
    out = np.empty(1_000_000, np.float32)
    itf = sp.identity(N_ITEMS, dtype=np.float32, format="csr")
    usf = sp.identity(N_USERS, dtype=np.float32, format="csr")
    fast.predict_lightfm(fast.CSRMatrix(itf), fast.CSRMatrix(usf), u, i, out, holder, 4)
    want = np.einsum("ij,ij->i", st[6][u].astype(np.float64), st[0][i].astype(np.float64)) + st[9][u] + st[3][i]
    assert np.allclose(out, want, atol=1e-5)


def test_host_entry_point_full_size_roundtrip(c2):
    """The stateless boundary call at full size: in-place contract (arrays mutated, finite) and
    H2D / D2H accounting equal to the sizes of what was passed."""
    from lightfm_b200 import _lightfm_fast as fast
    rows, cols = c2
    y = np.ones(NNZ, np.float32)
    pos = sp.csr_matrix((y, (rows, cols)), shape=(N_USERS, N_ITEMS))
    pos.sort_indices()
    st = _state(64)
    before = st[0].copy()
    holder = fast.FastLightFM(*st, 64, 0, 0.05, 0.95, 1e-6, 10)
    itf = sp.identity(N_ITEMS, dtype=np.float32, format="csr")
    usf = sp.identity(N_USERS, dtype=np.float32, format="csr")
    shuffle = np.arange(NNZ, dtype=np.int32)
    rs = np.random.RandomState(0)
    rs.shuffle(shuffle)
    fast.fit_warp(fast.CSRMatrix(itf), fast.CSRMatrix(usf), fast.CSRMatrix(pos), rows, cols, y, y, shuffle,
                  holder, 0.05, 0.0, 0.0, 8, rs)
    c = fast.last_counters["fit"]
    assert c["mode"] == 2 and c["positives"] == NNZ
    assert not np.array_equal(before, st[0]) and np.isfinite(st[0]).all() and np.isfinite(st[6]).all()
    state_bytes = 4 * ((N_ITEMS + N_USERS) * 65 * 2)
    assert c["d2h_bytes"] == state_bytes
    assert c["h2d_bytes"] >= state_bytes + 4 * NNZ * 4 + 4 * NNZ


def test_membership_bitmap_and_sorted_row_search_agree(c2):
    """Resident plans answer in_positives from an exact users x items bitmap when it fits; with the
    limit at 0 the kernels search the sorted CSR row instead.  Same data, same seeds: the rejection
    / sampling / update statistics of one epoch must coincide (hogwild => statistically)."""
    from lightfm_b200 import _lightfm_fast as fast
    rows, cols = c2
    y = np.ones(NNZ, np.float32)
    pos = sp.csr_matrix((y, (rows, cols)), shape=(N_USERS, N_ITEMS))
    pos.sort_indices()
    itf = sp.identity(N_ITEMS, dtype=np.float32, format="csr")
    usf = sp.identity(N_USERS, dtype=np.float32, format="csr")
    out = []
    for limit in (1 << 30, 0):
        fast.set_bitmap_limit(limit)
        try:
            st = _state(64)
            holder = fast.FastLightFM(*st, 64, 0, 0.05, 0.95, 1e-6, 10)
            plan = fast.ResidentPlan("warp", fast.CSRMatrix(itf), fast.CSRMatrix(usf), fast.CSRMatrix(pos),
                                     rows, cols, y, y, holder, 0.0, 0.0)
            c = plan.epoch(seed=11, num_threads=8)
            plan.close()
        finally:
            fast.set_bitmap_limit(1 << 30)
        out.append(c)
    a, b = out
    assert a["positives"] == b["positives"] == NNZ
    assert a["rejected"] > 0 and abs(a["rejected"] - b["rejected"]) < 0.03 * b["rejected"], (a["rejected"], b["rejected"])
    # same sampler, but the two kernels run at different speeds, so the concurrently evolving
    # model differs a little within the epoch: allow 3 %
    assert abs(a["negatives_drawn"] - b["negatives_drawn"]) < 0.03 * b["negatives_drawn"]
    assert abs(a["updates"] - b["updates"]) < 0.03 * b["updates"]
