"""GPU: the dependency-graph replay path (lfm_replay_dataflow.cuh) for BPR / logistic at
num_threads=1 must produce the SAME BITS as the sequential replay kernel (which is what the oracle /
golden tests pin against the reference), on every shape that stresses its scheduler: heavy
negative rejection, chunk boundaries, skipped (Y <= 0) interactions, sample weights, adadelta,
popular rows with long dependency chains, several epochs."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

pytestmark = pytest.mark.gpu


def _fit(loss, inter, hp, epochs, dataflow, sample_weight=None, seed=5):
    cu = H.cuda_native()
    prev = cu.module.set_replay_dataflow(dataflow)
    try:
        rs = np.random.RandomState(seed)
        arr = H.init_arrays(rs, inter.shape[1], inter.shape[0], hp.d, hp.schedule)
        counters = []
        for _ in range(epochs):
            H.run_epoch(cu, loss, inter, arr, hp, rs, sample_weight=sample_weight, num_threads=1)
            counters.append(dict(cu.module.last_counters["fit"]))
        return arr, counters
    finally:
        cu.module.set_replay_dataflow(prev)


def _same(loss, inter, hp, epochs=2, sample_weight=None):
    a, ca = _fit(loss, inter, hp, epochs, True, sample_weight)
    b, cb = _fit(loss, inter, hp, epochs, False, sample_weight)
    for k in H.MODEL_ARRAYS:
        assert np.array_equal(a[k], b[k]), "%s differs (max rel %.3g)" % (k, H.max_rel_diff(a[k], b[k]))
    for x, y in zip(ca, cb):
        for key in ("positives", "negatives_drawn", "updates", "rejected"):
            assert x[key] == y[key], (key, x[key], y[key])
    assert ca[0]["kernel_launches"] >= 1
    return a


@pytest.mark.parametrize("loss", ("bpr", "logistic"))
@pytest.mark.parametrize("schedule", ("adagrad", "adadelta"))
@pytest.mark.parametrize("d", (16, 33, 64, 128))
def test_dataflow_equals_sequential_replay(loss, schedule, d):
    inter = H.synthetic_interactions(300, 200, 6000, 3, signed=(loss == "logistic"))
    _same(loss, inter, H.Hyper(d=d, schedule=schedule))


@pytest.mark.parametrize("loss", ("bpr", "logistic"))
def test_dataflow_matches_oracle(loss):
    """... and, like the sequential kernel, the oracle to <= 1e-6 relative (device exp in the sigmoid)."""
    orc = H.oracle_native()
    inter = H.synthetic_interactions(300, 200, 6000, 4, signed=(loss == "logistic"))
    hp = H.Hyper(d=16)
    got, _ = _fit(loss, inter, hp, 2, True)
    rs = np.random.RandomState(5)
    want = H.init_arrays(rs, 200, 300, 16)
    for _ in range(2):
        H.run_epoch(orc, loss, inter, want, hp, rs, num_threads=1)
    for k in H.MODEL_ARRAYS:
        assert H.max_rel_diff(got[k], want[k]) <= 1e-5, k


@pytest.mark.parametrize("n", (1, 2, 31, 32, 33, 64, 65, 1000))
def test_chunk_boundaries(n):
    inter = H.synthetic_interactions(80, 70, n, 7)
    assert inter.nnz == n
    _same("bpr", inter, H.Hyper(d=16), epochs=3)
    _same("logistic", inter, H.Hyper(d=16), epochs=3)


def test_heavy_rejection_and_long_chains():
    """Dense users (most draws are positives and get rejected) and a handful of rows that every
    interaction touches: the schedule re-aligns after nearly every draw and the graph is a chain."""
    rng = np.random.default_rng(0)
    dense = (rng.random((20, 24)) < 0.8).astype(np.float32)
    dense[:, 0] = 1.0
    inter = sp.coo_matrix(dense)
    a = _same("bpr", inter, H.Hyper(d=32), epochs=3)
    assert all(np.isfinite(v).all() for v in a.values())
    _same("logistic", inter, H.Hyper(d=32), epochs=3)


def test_give_up_case_negative_equals_positive():
    """Every item is a positive of every user: each BPR draw is rejected until the loop gives up after
    no_examples draws and keeps the last one -- possibly the positive item itself (T:1123-1127)."""
    inter = sp.coo_matrix(np.ones((2, 3), np.float32))
    _same("bpr", inter, H.Hyper(d=16), epochs=3)
    inter = sp.coo_matrix(np.ones((3, 1), np.float32))
    _same("bpr", inter, H.Hyper(d=16), epochs=2)


def test_skipped_interactions_and_sample_weights():
    """BPR skips Y <= 0 without drawing (T:1112-1113); weights scale the loss."""
    inter = H.synthetic_interactions(200, 150, 5000, 9, signed=True)
    sw = np.random.default_rng(1).random(inter.nnz).astype(np.float32) * 2
    a, c = _fit("bpr", inter, H.Hyper(d=16), 1, True, sample_weight=sw)
    assert c[0]["positives"] == int((inter.data > 0).sum())
    _same("bpr", inter, H.Hyper(d=16), sample_weight=sw)
    _same("logistic", inter, H.Hyper(d=16), sample_weight=sw)


def test_midscale_c1_shape():
    """C1's shape (943 x 1682, 100 k, BPR d=16): the case BASELINE.json quotes for one thread."""
    inter = H.synthetic_interactions(943, 1682, 100_000, 1)
    _same("bpr", inter, H.Hyper(d=16), epochs=1)
