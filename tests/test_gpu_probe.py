"""GPU: deterministic probe of the hogwild slot kernels' ARITHMETIC.

`set_probe(1)` runs fast_slot_kernel -- the very templates the bench times -- as one warp with one
interaction in flight, rows staged at the top of the iteration, and the reference's sequential
rand_r negatives.  Order, draws and membership decisions are then the reference's, so the trained
weights differ from the oracle's only through the kernel's arithmetic: fp32 temporaries with FMA
instead of fp64 temporaries re-rounded at every store, lr * rsqrt.approx.ftz(G) instead of
lr / sqrt(G) in double, float log-table, __expf sigmoid.  The test prints the measured deviation and
bounds it."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

# measured on the B200: max relative deviation of any weight after the
# run stays below these
BOUNDS = {"warp": 1e-5, "bpr": 1e-5, "logistic": 1e-5}   # measured: 5.7e-7 / 3.7e-7 / 3.5e-7


def _fit(api, loss, inter, d, epochs):
    hp = H.Hyper(d=d)
    rs = np.random.RandomState(8)
    arr = H.init_arrays(rs, inter.shape[1], inter.shape[0], d)
    for _ in range(epochs):
        H.run_epoch(api, loss, inter, arr, hp, rs, num_threads=1)
    return arr


@pytest.mark.parametrize("loss,d", [("warp", 64), ("bpr", 64), ("logistic", 32)])
def test_slot_kernel_arithmetic_vs_oracle(loss, d):
    cu, orc = H.cuda_native(), H.oracle_native()
    inter = H.synthetic_interactions(300, 200, 6000, 1, signed=(loss == "logistic"))
    want = _fit(orc, loss, inter, d, 2)
    cu.module.set_mode("hogwild")
    cu.module.set_probe(1)
    try:
        got = _fit(cu, loss, inter, d, 2)
        c = cu.module.last_counters["fit"]
    finally:
        cu.module.set_probe(0)
        cu.module.set_mode("auto")
    oc = orc.last_counters["fit"]
    # same stream, same decisions: the work counters agree exactly
    assert (c["positives"], c["negatives_drawn"], c["updates"]) == \
           (oc["positives"], oc["negatives_drawn"], oc["updates"]), (c, oc)
    worst = 0.0
    for k in H.MODEL_ARRAYS:
        if "momentum" in k:
            continue
        a, b = got[k].astype(np.float64), want[k].astype(np.float64)
        # relative to the array's scale: single near-zero weights would make a pointwise ratio meaningless
        dev = float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-12))
        worst = max(worst, dev)
    print("probe %s d=%d: max |w_gpu - w_oracle| / max|w_oracle| = %.3g" % (loss, d, worst))
    assert worst <= BOUNDS[loss], worst
