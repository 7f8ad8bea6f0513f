"""CPU: libfm_cuda.so loads, exports every symbol include/lfm_cuda.h declares, and fails
loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H


def test_exports_every_declared_symbol():
    from lightfm_b200 import _abi, _lightfm_fast
    lib = ctypes.CDLL(_lightfm_fast.LIBRARY_PATH)
    header = open(os.path.join(H.ROOT, "include", "lfm_cuda.h")).read()
    # function declarations only: "<type> lfm_name(" at the start of a declaration line
    declared = set(re.findall(r"^(?:const char \*|int )(lfm_[a-z_]+)\(", header, flags=re.M))
    assert declared, "header parse failed"
    assert declared == set(_abi.declared_symbols()), declared ^ set(_abi.declared_symbols())
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_native_module_surface_matches_reference_names():
    from lightfm_b200 import _lightfm_fast as fast
    for name in ("CSRMatrix", "FastLightFM", "fit_logistic", "fit_warp", "fit_warp_kos", "fit_bpr",
                 "predict_lightfm", "predict_ranks", "calculate_auc_from_rank",
                 "__test_in_positives"):
        assert hasattr(fast, name), name


def test_no_cpu_fallback_without_device():
    from lightfm_b200 import _lightfm_fast as fast
    if fast.device_count() > 0:
        pytest.skip("a CUDA device is present")
    mat = fast.CSRMatrix(sp.identity(3, dtype=np.float32, format="csr"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        getattr(fast, "__test_in_positives")(0, 0, mat)
    from lightfm_b200 import LightFM
    model = LightFM(no_components=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.fit(sp.identity(5, dtype=np.float32, format="coo"))


def test_dtype_and_contiguity_checks_precede_compute():
    from lightfm_b200 import _lightfm_fast as fast
    bad = sp.identity(3, dtype=np.float64, format="csr")
    with pytest.raises(ValueError):
        fast.CSRMatrix(bad)
    arrays = H.init_arrays(np.random.RandomState(0), 3, 3, 4)
    arrays["item_embeddings"] = arrays["item_embeddings"].astype(np.float64)
    with pytest.raises(ValueError):
        H.holder(fast, arrays, H.Hyper(d=4))
    arrays = H.init_arrays(np.random.RandomState(0), 3, 3, 4)
    arrays["user_biases"].flags.writeable = False
    with pytest.raises(ValueError):
        H.holder(fast, arrays, H.Hyper(d=4))


def test_product_does_not_import_oracle():
    """The shipped path must never import, link, dlopen or execute anything under oracle/."""
    pat = re.compile(r"import\s+oracle|from\s+oracle|liblfm_oracle|lfm_oracle\.(c|so)|oracle/_ref|build_ref")
    pkg = os.path.join(H.ROOT, "lightfm_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                hits = [ln for ln in open(os.path.join(root, f)) if pat.search(ln)
                        and not ln.lstrip().startswith(("//", "#", "*"))]
                assert not hits, "%s references the oracle: %s" % (f, hits)
