"""ctypes wrapper over oracle/liblfm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Exposes the same ten names as the reference's native module, backed by the CPU
restatement in lfm_oracle.c, so a test can drive oracle and CUDA path with the
same arguments.  Also `load_reference(variant)` imports the REAL reference built
by build_ref.py into oracle/_ref (the stronger checker where available).
"""
import ctypes
import importlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from lightfm_b200 import _abi  # noqa: E402  (declarations only; no CUDA involved)

CSRMatrix = _abi.CSRMatrix
FastLightFM = _abi.FastLightFM


def build(force=False):
    so = os.path.join(HERE, "liblfm_oracle.so")
    src = os.path.join(HERE, "lfm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "liblfm_oracle.so"])
    return so


def _check(status):
    if status != 0:
        raise RuntimeError("oracle returned %d" % status)


_lib = _abi.bind(ctypes.CDLL(build()), prefix="oracle_", with_lib_state=False)
_api = _abi.make_api(_lib, "oracle_", _check)
fit_logistic = _api["fit_logistic"]
fit_warp = _api["fit_warp"]
fit_bpr = _api["fit_bpr"]
fit_warp_kos = _api["fit_warp_kos"]
predict_lightfm = _api["predict_lightfm"]
predict_ranks = _api["predict_ranks"]
calculate_auc_from_rank = _api["calculate_auc_from_rank"]
test_in_positives = _api["__test_in_positives"]
last_counters = _api["last_counters"]


def reference_available(variant="strict"):
    d = os.path.join(HERE, "_ref", variant, "lightfm")
    return os.path.isdir(d) and any(f.endswith(".so") for f in os.listdir(d))


def load_reference(variant="strict"):
    """Import the real reference package (`lightfm`) from oracle/_ref/<variant>."""
    root = os.path.join(HERE, "_ref", variant)
    if not reference_available(variant):
        raise ImportError("reference not built: run `python oracle/build_ref.py` "
                          "where /root/reference exists")
    if "lightfm" in sys.modules:
        mod = sys.modules["lightfm"]
        if not os.path.abspath(mod.__file__).startswith(os.path.abspath(root)):
            raise ImportError("another `lightfm` (%s) is already imported" % mod.__file__)
        return mod
    sys.path.insert(0, root)
    try:
        return importlib.import_module("lightfm")
    finally:
        sys.path.remove(root)
