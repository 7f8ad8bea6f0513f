"""GPU: replay mode on the semantic edge cases for which the oracle is pinned to the real reference
on the CPU (tests/test_oracle_golden.py, second half).

All twelve passed on the B200 in the round-1 driver run (then still marked xfail: 12 XPASS in
GPUTEST_r01.json); they are plain tests now."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

pytestmark = pytest.mark.gpu


def _both(loss, inter, d=8, epochs=2, sw=None, exact=None, **hpkw):
    cu, orc = H.cuda_native(), H.oracle_native()
    outs = []
    for api in (orc, cu):
        hp = H.Hyper(d=d, **hpkw)
        rs = np.random.RandomState(9)
        arr = H.init_arrays(rs, inter.shape[1], inter.shape[0], d, hp.schedule)
        for _ in range(epochs):
            H.run_epoch(api, loss, inter, arr, hp, rs, sample_weight=sw, num_threads=1)
        outs.append(arr)
    exact = loss in ("warp", "warp-kos") if exact is None else exact
    for k in H.MODEL_ARRAYS:
        if exact:
            assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k
        else:
            assert H.max_rel_diff(outs[1][k], outs[0][k]) <= 1e-5, k


@pytest.mark.parametrize("loss", ("warp", "bpr"))
def test_nonpositive_rows(loss):
    _both(loss, H.synthetic_interactions(50, 40, 600, 2, signed=True))


@pytest.mark.parametrize("loss", ("warp", "bpr", "logistic"))
def test_sample_weights(loss):
    inter = H.synthetic_interactions(50, 40, 600, 2, signed=(loss == "logistic"))
    sw = (0.25 + np.random.default_rng(1).random(inter.nnz) * 3).astype(np.float32)
    _both(loss, inter, sw=sw)


@pytest.mark.parametrize("loss", ("warp", "bpr", "warp-kos"))
def test_duplicate_coo_entries(loss):
    base = H.synthetic_interactions(40, 30, 300, 3)
    rows = np.concatenate([base.row, base.row[:80]])
    cols = np.concatenate([base.col, base.col[:80]])
    _both(loss, sp.coo_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=base.shape))


def test_kos_k_and_n_larger_than_history():
    _both("warp-kos", H.synthetic_interactions(60, 50, 150, 5), k=7, n=12)


def test_max_sampled_one_and_many():
    inter = H.synthetic_interactions(50, 40, 600, 2)
    _both("warp", inter, max_sampled=1)
    _both("warp", inter, max_sampled=37)


def test_tiny_catalogue_nan_propagation_in_kos():
    _both("warp-kos", H.synthetic_interactions(20, 3, 40, 1), epochs=1, k=2, n=3)


def test_odd_component_counts_and_adadelta_eps_zero():
    inter = H.synthetic_interactions(50, 40, 600, 2)
    _both("warp", inter, d=1)
    _both("bpr", inter, d=33, schedule="adadelta", eps=0.0)
