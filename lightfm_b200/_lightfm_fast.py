"""Drop-in replacement for the reference's ``lightfm/_lightfm_fast.py`` import shim.

The reference module re-exports the names of its Cython extension
(``/root/reference/lightfm/_lightfm_fast.py:1-15``).  This one exports the same
ten names -- ``CSRMatrix, FastLightFM, fit_logistic, fit_warp, fit_warp_kos,
fit_bpr, predict_lightfm, predict_ranks, calculate_auc_from_rank,
__test_in_positives`` -- bound with ctypes to ``libfm_cuda.so`` (C ABI in
``include/lfm_cuda.h``), whose kernels are hand-written CUDA for sm_100a.

There is no CPU fallback: if the shared library is missing the import fails, and
if no CUDA device is usable every compute call raises ``RuntimeError``.

Execution mode (``num_threads`` has no natural meaning on a GPU):
  ``num_threads == 1``  deterministic replay of the reference's single-thread order
  ``num_threads  > 1``  hogwild throughput mode (thousands of interactions in flight, a slot of
                        4-32 lanes each, lock-free updates)
override with ``LIGHTFM_CUDA_MODE=replay|hogwild|auto`` or :func:`set_mode`.
"""
import ctypes
import os

import numpy as np

from . import _abi
from ._abi import CSRMatrix, FastLightFM  # noqa: F401  (re-exported)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("LIGHTFM_CUDA_LIB", os.path.join(_HERE, "csrc", "libfm_cuda.so"))

if not os.path.exists(_LIB_PATH):
    raise ImportError(
        "libfm_cuda.so not found at %s. Build it with `python -m lightfm_b200._build` "
        "(needs nvcc). lightfm_b200 has no CPU fallback." % _LIB_PATH)

_lib = _abi.bind(ctypes.CDLL(_LIB_PATH), prefix="lfm_")
LIBRARY_PATH = _LIB_PATH


def _check(status):
    if status == 0:
        return
    msg = (_lib.lfm_last_error() or b"").decode("utf-8", "replace")
    if status == -1:
        raise ValueError(msg)
    if status == -3:
        raise MemoryError(msg)
    raise RuntimeError("libfm_cuda: %s (status %d)" % (msg, status))


_MODES = {"auto": 0, "replay": 1, "hogwild": 2}


def set_mode(mode):
    """'auto' (num_threads==1 -> replay, >1 -> hogwild), 'replay' or 'hogwild'."""
    _check(_lib.lfm_set_mode(_MODES[mode]))


def get_mode():
    return {v: k for k, v in _MODES.items()}[_lib.lfm_get_mode()]


def device_count():
    return _lib.lfm_device_count()


def set_device(device):
    """Select the CUDA device of this process (before the first compute call; one process per
    GPU).  Raises if the library was already initialised on another device."""
    _check(_lib.lfm_set_device(int(device)))


def resolves_to_hogwild(num_threads):
    """True when a fit call with this num_threads runs the throughput kernels."""
    mode = _lib.lfm_get_mode()
    return mode == 2 or (mode == 0 and num_threads > 1)


def set_fast_path(enabled):
    """Testing hook: disable the specialised hogwild kernels (generic ones run instead)."""
    return _lib.lfm_set_fast_path(int(bool(enabled)))


def set_probe(enabled):
    """Testing hook: run the slot kernels as one warp with one interaction in flight and the
    reference's rand_r negatives (tests/test_gpu_probe.py: arithmetic-only comparison with the oracle)."""
    return _lib.lfm_set_probe(int(bool(enabled)))


def set_atomic_accumulators(enabled):
    """Hogwild slot kernels: scale every Adagrad step by the accumulator value returned by an atomic
    add (default on) instead of a value read earlier (round-1 behaviour)."""
    return _lib.lfm_set_atomic_accumulators(int(bool(enabled)))


def set_rank_groups(groups):
    """predict_ranks tiling: 1 or 3 user tiles per CTA (see lfm_set_rank_groups)."""
    return _lib.lfm_set_rank_groups(int(groups))


def set_replay_fast(enabled):
    """Replay mode: use the prefetching WARP kernel where it applies (default on; both bit-equal)."""
    return _lib.lfm_set_replay_fast(int(bool(enabled)))


def set_replay_dataflow(enabled):
    """Replay mode, BPR / logistic with identity features: walk the epoch as a dependency graph on
    many warps (default on) instead of sequentially; the same bits either way."""
    return _lib.lfm_set_replay_dataflow(int(bool(enabled)))


def last_replay_dataflow():
    """(scheduler warp ms, whole kernel ms, tasks) of the last dataflow replay epoch."""
    import ctypes as C
    a, b, n = C.c_double(0), C.c_double(0), C.c_int32(0)
    _lib.lfm_last_replay_dataflow(C.byref(a), C.byref(b), C.byref(n))
    return a.value, b.value, n.value


def set_hot_rows(enabled):
    """Feature path: per-CTA shared-memory aggregation of hot feature rows (default on)."""
    return _lib.lfm_set_hot_rows(int(bool(enabled)))


def set_inflight_divisor(divisor):
    """Hogwild launches keep at most max(64, n / divisor) interactions in flight (default 128)."""
    return _lib.lfm_set_inflight_divisor(int(divisor))


def set_tuning(variant):
    """WARP fast-path kernel variant: 0 = warp per interaction (v1); 4/5 = slot per interaction,
    one float4 per lane, 3/4 CTAs per SM; 6/7/8 = two float4 per lane, 2/3/4 CTAs per SM."""
    return _lib.lfm_set_tuning(int(variant))


_TUNING_NAMES = {0: "fast_rank_kernel<WARP,LPR=d/4>", 4: "fast_slot_kernel<WARP,d,1,3>", 5: "fast_slot_kernel<WARP,d,1,4>",
                 6: "fast_slot_kernel<WARP,d,2,2>", 7: "fast_slot_kernel<WARP,d,2,3>",
                 8: "fast_slot_kernel<WARP,d,2,4>", 9: "fast_slot_kernel<WARP,d,2,3,SPEC>",
                 10: "fast_slot_kernel<WARP,d,2,2,SPEC>"}


def warp_kernel_name(d=64):
    """Name of the WARP fast-path kernel the current tuning launches (for reports)."""
    t = set_tuning(-1)  # out-of-range: returns the current value without changing it
    return _TUNING_NAMES.get(t, "?").replace(",d,", ",%d," % d).replace("LPR=d/4", "LPR=%d" % (d // 4))


_bitmap_limit = [1 << 30]


def bitmap_limit():
    return _bitmap_limit[0]


def set_bitmap_limit(nbytes):
    """Resident plans build an exact users x items membership bitmap of the positives when it
    fits in `nbytes` (default 1 GiB; 0 disables it and every kernel uses the sorted-row search)."""
    _bitmap_limit[0] = max(0, int(nbytes))
    return _lib.lfm_set_bitmap_limit(int(nbytes))


def release_cache():
    _check(_lib.lfm_release_cache())


class PinnedArrays(object):
    """Page-lock a set of long-lived numpy arrays in place (cudaHostRegister) and keep them alive
    until ``release()``: memory must never be freed while it is registered."""

    def __init__(self, arrays):
        self.arrays = []
        for a in arrays:
            if a is None or a.nbytes == 0:
                continue
            if _lib.lfm_pin_host(ctypes.c_void_p(a.ctypes.data), a.nbytes) == 0:
                self.arrays.append(a)

    def holds(self, arrays):
        have = {id(a) for a in self.arrays}
        return all(a is None or a.nbytes == 0 or id(a) in have for a in arrays)

    def release(self):
        for a in self.arrays:
            _lib.lfm_unpin_host(ctypes.c_void_p(a.ctypes.data))
        self.arrays = []

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


if os.environ.get("LIGHTFM_CUDA_INFLIGHT_DIVISOR"):
    set_inflight_divisor(int(os.environ["LIGHTFM_CUDA_INFLIGHT_DIVISOR"]))
if os.environ.get("LIGHTFM_CUDA_BITMAP_LIMIT"):
    set_bitmap_limit(int(os.environ["LIGHTFM_CUDA_BITMAP_LIMIT"]))
if os.environ.get("LIGHTFM_CUDA_TUNING"):
    set_tuning(int(os.environ["LIGHTFM_CUDA_TUNING"]))
if os.environ.get("LIGHTFM_CUDA_MODE"):
    set_mode(os.environ["LIGHTFM_CUDA_MODE"].lower())

_api = _abi.make_api(_lib, "lfm_", _check)
fit_logistic = _api["fit_logistic"]
fit_warp = _api["fit_warp"]
fit_bpr = _api["fit_bpr"]
fit_warp_kos = _api["fit_warp_kos"]
predict_lightfm = _api["predict_lightfm"]
predict_ranks = _api["predict_ranks"]
calculate_auc_from_rank = _api["calculate_auc_from_rank"]
globals()["__test_in_positives"] = _api["__test_in_positives"]
# Work counters of the most recent fit_* call (negatives drawn, updates, device ms, bytes).
last_counters = _api["last_counters"]

__all__ = ["CSRMatrix", "FastLightFM", "fit_logistic", "fit_warp", "fit_bpr", "fit_warp_kos",
           "predict_lightfm", "predict_ranks", "calculate_auc_from_rank"]


def evaluate_ranks(item_features, user_features, test_interactions, train_interactions, lightfm, k,
                   want_hits=True, want_best=True, want_auc=True, num_threads=1):
    """predict_ranks + the per-user reductions of lightfm/evaluation.py on the device
    (``lfm_evaluate_ranks``).  Returns (hits int32[U] | None, best_rank float32[U] | None,
    auc float32[U] | None); the per-interaction ranks never reach the host."""
    rows = test_interactions.rows
    hits = np.zeros(rows, np.int32) if want_hits else None
    best = np.zeros(rows, np.float32) if want_best else None
    auc = np.zeros(rows, np.float32) if want_auc else None
    _check(_lib.lfm_evaluate_ranks(
        item_features.ptr, user_features.ptr, test_interactions.ptr, train_interactions.ptr, lightfm.ptr,
        int(k), _abi.i32p(hits) if want_hits else ctypes.cast(None, _abi.c_i32p),
        _abi.f32p(best) if want_best else ctypes.cast(None, _abi.c_f32p),
        _abi.f32p(auc) if want_auc else ctypes.cast(None, _abi.c_f32p), int(num_threads)))
    return hits, best, auc


def last_scoring_ms():
    """Device milliseconds of the kernels of the last predict_ranks / evaluate_ranks / recommend call."""
    ms = ctypes.c_double()
    _check(_lib.lfm_last_scoring_ms(ctypes.byref(ms)))
    return ms.value


def recommend(item_features, user_features, exclude, user_ids, n_items, k, lightfm):
    """Top-k items per user (``lfm_recommend``): (items int32[n, k], scores float32[n, k])."""
    _abi._require(user_ids, np.int32, 1, "user_ids")
    n = len(user_ids)
    items = np.full((n, int(k)), -1, np.int32)
    scores = np.full((n, int(k)), np.nan, np.float32)
    _check(_lib.lfm_recommend(item_features.ptr, user_features.ptr,
                              exclude.ptr if exclude is not None else None, _abi.i32p(user_ids), n,
                              int(n_items), int(k), lightfm.ptr, _abi.i32p(items), _abi.f32p(scores)))
    return items, scores


_LOSS_CODES = {"logistic": 0, "warp": 1, "bpr": 2, "warp-kos": 3}


class ResidentPlan(object):
    """One training problem kept in HBM across epochs (``lfm_plan_*`` in lfm_cuda.h).

    Uploads the interactions, feature matrices, positives lookup and the model once;
    ``epoch()`` then runs entirely on the device; ``download()`` writes the model back
    into the numpy arrays of the ``FastLightFM`` it was created from.
    """

    def __init__(self, loss, item_features, user_features, interactions, user_ids, item_ids, Y,
                 sample_weight, lightfm, item_alpha, user_alpha, k=5, n=10):
        self._handle = ctypes.c_void_p()
        self._args = (loss, item_features, user_features, interactions, user_ids, item_ids, Y,
                      sample_weight, lightfm, item_alpha, user_alpha, k, n)
        self._upload()

    def refresh(self):
        """Upload every input and the model again into the plan's existing device buffers."""
        self._upload()

    def _upload(self):
        (loss, item_features, user_features, interactions, user_ids, item_ids, Y, sample_weight,
         lightfm, item_alpha, user_alpha, k, n) = self._args
        self._lightfm = lightfm
        # same typed-buffer checks as the per-epoch entry points (int32 ids, float32 values)
        _abi._require(user_ids, np.int32, 1, "user_ids")
        if item_ids is not None:
            _abi._require(item_ids, np.int32, 1, "item_ids")
        if Y is not None:
            _abi._require(Y, np.float32, 1, "Y")
        if sample_weight is not None:
            _abi._require(sample_weight, np.float32, 1, "sample_weight")
        for name, arr in (("item_ids", item_ids), ("Y", Y), ("sample_weight", sample_weight)):
            if arr is not None and len(arr) != len(user_ids):
                raise ValueError("%s has %d entries, expected %d" % (name, len(arr), len(user_ids)))
        n_ex = len(user_ids)
        null_i = ctypes.cast(None, _abi.c_i32p)
        null_f = ctypes.cast(None, _abi.c_f32p)
        _check(_lib.lfm_plan_create(
            ctypes.byref(self._handle), _LOSS_CODES[loss], item_features.ptr, user_features.ptr,
            interactions.ptr if interactions is not None else None,
            _abi.i32p(user_ids), _abi.i32p(item_ids) if item_ids is not None else null_i,
            _abi.f32p(Y) if Y is not None else null_f,
            _abi.f32p(sample_weight) if sample_weight is not None else null_f,
            n_ex, lightfm.ptr, float(item_alpha), float(user_alpha), int(k), int(n)))

    def epoch(self, seed, num_threads=2, shuffle_indices=None, next_seed=None):
        """One epoch.  ``next_seed``: the seed the following ``epoch`` call will use (device-generated
        order only) -- its tuples are packed beside this epoch's SGD kernel."""
        if next_seed is not None and shuffle_indices is None:
            cnt = _abi.LfmCounters()
            _check(_lib.lfm_plan_epoch_next(self._handle, int(seed) & 0xFFFFFFFF, int(next_seed) & 0xFFFFFFFF,
                                            int(num_threads), ctypes.byref(cnt)))
            return cnt.as_dict()
        if shuffle_indices is not None:
            _abi._require(shuffle_indices, np.int32, 1, "shuffle_indices")
            if len(shuffle_indices) != len(self._args[4]):
                raise ValueError("shuffle_indices has the wrong length")
        cnt = _abi.LfmCounters()
        sh = _abi.i32p(shuffle_indices) if shuffle_indices is not None else ctypes.cast(None, _abi.c_i32p)
        _check(_lib.lfm_plan_epoch(self._handle, sh, int(seed) & 0xFFFFFFFF, int(num_threads),
                                   ctypes.byref(cnt)))
        return cnt.as_dict()

    def epoch_range(self, seed, begin, count=-1, num_threads=2):
        """One pass over interactions [begin, begin + count) of the uploaded list, random order."""
        cnt = _abi.LfmCounters()
        _check(_lib.lfm_plan_epoch_range(self._handle, int(seed) & 0xFFFFFFFF, int(num_threads), int(begin),
                                         int(count), ctypes.byref(cnt)))
        return cnt.as_dict()

    def delta_begin(self, side, row_begin=0, row_count=-1):
        """Snapshot rows of one side's w, g, b, bg before a local epoch (side 0 item, 1 user)."""
        ms = ctypes.c_double()
        _check(_lib.lfm_plan_delta_begin(self._handle, int(side), int(row_begin), int(row_count), ctypes.byref(ms)))
        return ms.value

    def delta_make(self, side):
        """(device pointer, float count, sweep ms) of the local delta since ``delta_begin``: all-reduce it."""
        ptr, cnt, ms = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_double()
        _check(_lib.lfm_plan_delta_make(self._handle, int(side), ctypes.byref(ptr), ctypes.byref(cnt),
                                        ctypes.byref(ms)))
        return ptr.value, cnt.value, ms.value

    def delta_apply(self, side):
        """Add the other ranks' share of the all-reduced delta to the resident table (returns sweep ms)."""
        ms = ctypes.c_double()
        _check(_lib.lfm_plan_delta_apply(self._handle, int(side), ctypes.byref(ms)))
        return ms.value

    def download(self):
        _check(_lib.lfm_plan_download(self._handle, self._lightfm.ptr))

    def upload_model(self, lightfm, wait=True):
        """Refresh the resident state from the arrays of `lightfm` (same shapes as the plan's);
        later ``download()`` calls write into these arrays.  ``wait=False``: return while the copies
        are in flight (the next ``epoch`` packs its tuples beside them); the arrays must not be
        touched before that call returns."""
        if wait:
            _check(_lib.lfm_plan_upload_model(self._handle, lightfm.ptr))
        else:
            _check(_lib.lfm_plan_upload_model_async(self._handle, lightfm.ptr))
        self._lightfm = lightfm

    def all_finite(self):
        ok = ctypes.c_int32(0)
        _check(_lib.lfm_plan_check_finite(self._handle, ctypes.byref(ok)))
        return bool(ok.value)

    def set_global_items(self, n_items_global):
        _check(_lib.lfm_plan_set_global_items(self._handle, int(n_items_global)))

    def table(self, which):
        """(device pointer, element count) of resident state array `which` (see lfm_plan_table)."""
        ptr, cnt = ctypes.c_void_p(), ctypes.c_int64()
        _check(_lib.lfm_plan_table(self._handle, int(which), ctypes.byref(ptr), ctypes.byref(cnt)))
        return ptr.value, cnt.value

    def close(self):
        if self._handle:
            _lib.lfm_plan_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
