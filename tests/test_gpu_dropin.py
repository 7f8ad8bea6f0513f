"""GPU: the drop-in claim itself.  The UNMODIFIED reference `lightfm.py` / `evaluation.py`
(as built into oracle/_ref by oracle/build_ref.py) are placed in a scratch package whose
`_lightfm_fast.py` is our ctypes shim; that package must reproduce the real reference
(oracle/_ref/strict, its own Cython extension) at num_threads=1 -- bit for bit where libm does
not enter (WARP, k-OS), <= 1e-5 relative otherwise."""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

REF_SRC = os.path.join(H.ROOT, "oracle", "_ref", "csrc")


@pytest.fixture(scope="module")
def dropin(tmp_path_factory):
    import oracle
    if not oracle.reference_available("strict") or not os.path.exists(os.path.join(REF_SRC, "lightfm.py")):
        pytest.skip("oracle/_ref not present (built by oracle/build_ref.py where /root/reference exists)")
    root = tmp_path_factory.mktemp("dropin")
    pkg = root / "lightfm_dropin"
    pkg.mkdir()
    for f in ("lightfm.py", "evaluation.py"):          # the reference's files, byte for byte
        shutil.copy(os.path.join(REF_SRC, f), pkg / f)
    (pkg / "__init__.py").write_text("")
    (pkg / "_lightfm_fast.py").write_text(              # the one file a maintainer replaces
        "from lightfm_b200._lightfm_fast import *  # noqa\n"
        "from lightfm_b200 import _lightfm_fast as _m\n"
        "globals()['__test_in_positives'] = getattr(_m, '__test_in_positives')\n")
    sys.path.insert(0, str(root))
    try:
        mod = importlib.import_module("lightfm_dropin.lightfm")
        ev = importlib.import_module("lightfm_dropin.evaluation")
    finally:
        sys.path.remove(str(root))
    ref = oracle.load_reference("strict")
    ref_ev = importlib.import_module("lightfm.evaluation")
    return mod.LightFM, ev, ref.LightFM, ref_ev


@pytest.mark.parametrize("loss", ("warp", "warp-kos", "bpr", "logistic"))
def test_reference_class_over_our_shim_equals_reference(dropin, loss):
    Ours, ev, Ref, ref_ev = dropin
    data = H.synthetic_interactions(150, 110, 3000, 11, signed=(loss == "logistic"))
    train, test = H.split(H.synthetic_interactions(150, 110, 3000, 11), 5)
    fit_on = data if loss == "logistic" else train
    a = Ref(loss=loss, no_components=16, random_state=3).fit(fit_on, epochs=2, num_threads=1)
    b = Ours(loss=loss, no_components=16, random_state=3).fit(fit_on, epochs=2, num_threads=1)
    exact = loss in ("warp", "warp-kos")
    for k in H.MODEL_ARRAYS:
        x, y = getattr(a, k), getattr(b, k)
        if exact:
            assert np.array_equal(x, y), k
        else:
            assert H.max_rel_diff(y, x) <= 1e-5, k
    # both RandomStates were consumed identically (shuffle + randint per epoch)
    assert np.array_equal(a.random_state.get_state()[1], b.random_state.get_state()[1])
    if exact:
        u = np.arange(150, dtype=np.int32).repeat(3)
        i = np.tile(np.arange(3, dtype=np.int32), 150)
        assert np.array_equal(a.predict(u, i), b.predict(u, i))
        ra = a.predict_rank(test.tocsr(), train_interactions=train.tocsr())
        rb = b.predict_rank(test.tocsr(), train_interactions=train.tocsr())
        assert np.array_equal(ra.data, rb.data)
        assert np.array_equal(ref_ev.auc_score(a, test.tocsr(), train_interactions=train.tocsr()),
                              ev.auc_score(b, test.tocsr(), train_interactions=train.tocsr()))
        assert np.array_equal(ref_ev.precision_at_k(a, test.tocsr(), train_interactions=train.tocsr(), k=5),
                              ev.precision_at_k(b, test.tocsr(), train_interactions=train.tocsr(), k=5))


def test_in_positives_hook_through_dropin(dropin):
    import scipy.sparse as sp
    mod = importlib.import_module("lightfm_dropin._lightfm_fast")
    mat = mod.CSRMatrix(sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=np.float32)))
    hook = getattr(mod, "__test_in_positives")
    assert [hook(0, 0, mat), hook(0, 1, mat), hook(1, 0, mat), hook(1, 1, mat)] == [False, True, True, False]
