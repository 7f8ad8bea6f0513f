"""ctypes mirror of ``include/lfm_cuda.h`` (structs + function prototypes).

Pure declarations: no compute, no fallback.  ``bind(lib)`` attaches argtypes /
restypes for every symbol the header declares to a loaded shared object and
raises if one is missing, so an out-of-date ``libfm_cuda.so`` fails at import
time rather than at the first call.
"""
import ctypes as C

import numpy as np

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_f32p = C.POINTER(C.c_float)


class LfmCsr(C.Structure):
    _fields_ = [
        ("indptr", c_i32p),
        ("indices", c_i32p),
        ("data", c_f32p),
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("nnz", C.c_int64),
    ]


class LfmModel(C.Structure):
    _fields_ = [
        ("item_features", c_f32p),
        ("item_feature_gradients", c_f32p),
        ("item_feature_momentum", c_f32p),
        ("item_biases", c_f32p),
        ("item_bias_gradients", c_f32p),
        ("item_bias_momentum", c_f32p),
        ("user_features", c_f32p),
        ("user_feature_gradients", c_f32p),
        ("user_feature_momentum", c_f32p),
        ("user_biases", c_f32p),
        ("user_bias_gradients", c_f32p),
        ("user_bias_momentum", c_f32p),
        ("n_item_features", C.c_int32),
        ("n_user_features", C.c_int32),
        ("no_components", C.c_int32),
        ("adadelta", C.c_int32),
        ("learning_rate", C.c_float),
        ("rho", C.c_float),
        ("eps", C.c_float),
        ("max_sampled", C.c_int32),
    ]


class LfmCounters(C.Structure):
    _fields_ = [
        ("positives", C.c_int64),
        ("negatives_drawn", C.c_int64),
        ("updates", C.c_int64),
        ("rejected", C.c_int64),
        ("kernel_ms", C.c_double),
        ("train_kernel_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("kernel_launches", C.c_int32),
        ("mode", C.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


CsrP = C.POINTER(LfmCsr)
ModelP = C.POINTER(LfmModel)
CountersP = C.POINTER(LfmCounters)

# name -> (restype, argtypes); `prefix` lets the CPU oracle (oracle_*) share the table.
_FIT_COMMON_TAIL = [ModelP, C.c_double, C.c_double, C.c_int32]
PROTOTYPES = {
    "fit_logistic": (C.c_int, [CsrP, CsrP, c_i32p, c_i32p, c_f32p, c_f32p, c_i32p, C.c_int64]
                     + _FIT_COMMON_TAIL + [CountersP]),
    "fit_warp": (C.c_int, [CsrP, CsrP, CsrP, c_i32p, c_i32p, c_f32p, c_f32p, c_i32p, C.c_int64]
                 + _FIT_COMMON_TAIL + [c_u32p, C.c_int32, CountersP]),
    "fit_warp_kos": (C.c_int, [CsrP, CsrP, CsrP, c_i32p, c_i32p, C.c_int64, ModelP, C.c_double,
                               C.c_double, C.c_int32, C.c_int32, C.c_int32, c_u32p, C.c_int32,
                               CountersP]),
    "fit_bpr": (C.c_int, [CsrP, CsrP, CsrP, c_i32p, c_i32p, c_f32p, c_f32p, c_i32p, C.c_int64]
                + _FIT_COMMON_TAIL + [c_u32p, C.c_int32, CountersP]),
    "predict_lightfm": (C.c_int, [CsrP, CsrP, c_i32p, c_i32p, c_f32p, C.c_int64, ModelP,
                                  C.c_int32]),
    "predict_ranks": (C.c_int, [CsrP, CsrP, CsrP, CsrP, c_f32p, ModelP, C.c_int32]),
    "calculate_auc_from_rank": (C.c_int, [CsrP, c_i32p, c_f32p, c_f32p, C.c_int32]),
    "test_in_positives": (C.c_int, [C.c_int32, C.c_int32, CsrP]),
}

LIB_ONLY = {
    "lfm_last_error": (C.c_char_p, []),
    "lfm_version": (C.c_char_p, []),
    "lfm_device_count": (C.c_int, []),
    "lfm_set_device": (C.c_int, [C.c_int]),
    "lfm_set_mode": (C.c_int, [C.c_int]),
    "lfm_get_mode": (C.c_int, []),
    "lfm_set_bitmap_limit": (C.c_int, [C.c_int64]),
    "lfm_release_cache": (C.c_int, []),
    "lfm_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, CsrP, CsrP, CsrP, c_i32p, c_i32p,
                                  c_f32p, c_f32p, C.c_int64, ModelP, C.c_double, C.c_double,
                                  C.c_int32, C.c_int32]),
    "lfm_plan_epoch": (C.c_int, [C.c_void_p, c_i32p, C.c_uint32, C.c_int32, CountersP]),
    "lfm_plan_epoch_next": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, CountersP]),
    "lfm_plan_epoch_range": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, C.c_int64, C.c_int64, CountersP]),
    "lfm_plan_delta_begin": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_double)]),
    "lfm_plan_delta_make": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_double)]),
    "lfm_plan_delta_apply": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "lfm_plan_download": (C.c_int, [C.c_void_p, ModelP]),
    "lfm_plan_upload_model": (C.c_int, [C.c_void_p, ModelP]),
    "lfm_plan_upload_model_async": (C.c_int, [C.c_void_p, ModelP]),
    "lfm_evaluate_ranks": (C.c_int, [CsrP, CsrP, CsrP, CsrP, ModelP, C.c_int32, c_i32p, c_f32p, c_f32p,
                                     C.c_int32]),
    "lfm_recommend": (C.c_int, [CsrP, CsrP, CsrP, c_i32p, C.c_int64, C.c_int32, C.c_int32, ModelP,
                                c_i32p, c_f32p]),
    "lfm_last_scoring_ms": (C.c_int, [C.POINTER(C.c_double)]),
    "lfm_pin_host": (C.c_int, [C.c_void_p, C.c_int64]),
    "lfm_unpin_host": (C.c_int, [C.c_void_p]),
    "lfm_set_tuning": (C.c_int, [C.c_int]),
    "lfm_set_fast_path": (C.c_int, [C.c_int]),
    "lfm_set_inflight_divisor": (C.c_int, [C.c_int]),
    "lfm_set_probe": (C.c_int, [C.c_int]),
    "lfm_set_rank_groups": (C.c_int, [C.c_int]),
    "lfm_set_atomic_accumulators": (C.c_int, [C.c_int]),
    "lfm_set_replay_fast": (C.c_int, [C.c_int]),
    "lfm_set_replay_dataflow": (C.c_int, [C.c_int]),
    "lfm_last_replay_dataflow": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "lfm_set_hot_rows": (C.c_int, [C.c_int]),
    "lfm_plan_table": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "lfm_plan_check_finite": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "lfm_plan_set_global_items": (C.c_int, [C.c_void_p, C.c_int32]),
    "lfm_plan_destroy": (C.c_int, [C.c_void_p]),
}


def bind(lib, prefix="lfm_", with_lib_state=True):
    """Attach prototypes; raise AttributeError naming any missing symbol."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
    if with_lib_state:
        for name, (res, args) in LIB_ONLY.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
    return lib


def declared_symbols():
    """Every symbol include/lfm_cuda.h declares (used by the CPU load test)."""
    return ["lfm_" + n for n in PROTOTYPES] + list(LIB_ONLY)


# ---- numpy -> pointer helpers -------------------------------------------------

def _require(arr, dtype, ndim, name, writable=False):
    """Typed-memoryview-style coercion checks (T: ``flt[::1]`` / ``int[::1]``):
    wrong dtype / ndim / non-contiguous / read-only raise before any compute."""
    if not isinstance(arr, np.ndarray):
        raise TypeError("%s: expected a numpy array, got %s" % (name, type(arr).__name__))
    if arr.dtype != dtype:
        raise ValueError("Buffer dtype mismatch for %s, expected '%s' but got '%s'"
                         % (name, np.dtype(dtype).name, arr.dtype.name))
    if arr.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions for %s (expected %d, got %d)"
                         % (name, ndim, arr.ndim))
    if not arr.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous (%s)" % name)
    if writable and not arr.flags.writeable:
        raise ValueError("buffer source array is read-only (%s)" % name)
    return arr


def f32p(arr):
    return arr.ctypes.data_as(c_f32p)


def i32p(arr):
    return arr.ctypes.data_as(c_i32p)


def u32p(arr):
    return arr.ctypes.data_as(c_u32p)


# ---- the Python-visible surface of the native module ---------------------------
# Mirrors the 10 names the reference's Cython extension exports:
# CSRMatrix, FastLightFM, fit_logistic, fit_warp, fit_warp_kos, fit_bpr,
# predict_lightfm, predict_ranks, calculate_auc_from_rank, __test_in_positives.

class CSRMatrix(object):
    """Borrowed view of a scipy CSR matrix (reference: T:145-182)."""

    def __init__(self, csr_matrix):
        self.indices = _require(csr_matrix.indices, np.int32, 1, "indices")
        self.indptr = _require(csr_matrix.indptr, np.int32, 1, "indptr")
        self.data = _require(csr_matrix.data, np.float32, 1, "data")
        self.rows, self.cols = csr_matrix.shape
        self.nnz = len(self.data)
        self._c = LfmCsr(i32p(self.indptr), i32p(self.indices), f32p(self.data),
                         int(self.rows), int(self.cols), int(self.nnz))

    @property
    def ptr(self):
        return C.byref(self._c)


_MODEL_ARRAYS = ["item_features", "item_feature_gradients", "item_feature_momentum",
                 "item_biases", "item_bias_gradients", "item_bias_momentum",
                 "user_features", "user_feature_gradients", "user_feature_momentum",
                 "user_biases", "user_bias_gradients", "user_bias_momentum"]


class FastLightFM(object):
    """Borrowed view of the 12 model arrays + hyper-parameters (reference: T:185-259)."""

    def __init__(self, item_features, item_feature_gradients, item_feature_momentum,
                 item_biases, item_bias_gradients, item_bias_momentum,
                 user_features, user_feature_gradients, user_feature_momentum,
                 user_biases, user_bias_gradients, user_bias_momentum,
                 no_components, adadelta, learning_rate, rho, epsilon, max_sampled):
        args = [item_features, item_feature_gradients, item_feature_momentum,
                item_biases, item_bias_gradients, item_bias_momentum,
                user_features, user_feature_gradients, user_feature_momentum,
                user_biases, user_bias_gradients, user_bias_momentum]
        for name, arr in zip(_MODEL_ARRAYS, args):
            ndim = 2 if name.endswith(("features", "feature_gradients", "feature_momentum")) else 1
            setattr(self, name, _require(arr, np.float32, ndim, name, writable=True))
        self.no_components = int(no_components)
        self.adadelta = int(adadelta)
        self.learning_rate = float(learning_rate)
        self.rho = float(rho)
        self.eps = float(epsilon)
        self.max_sampled = int(max_sampled)
        d = self.no_components
        for name in ("item", "user"):
            emb = getattr(self, name + "_features")
            if emb.shape[1] != d:
                raise ValueError("%s_features has %d columns, expected no_components=%d"
                                 % (name, emb.shape[1], d))
            n = emb.shape[0]
            for suffix, shape in (("_feature_gradients", (n, d)), ("_feature_momentum", (n, d)),
                                  ("_biases", (n,)), ("_bias_gradients", (n,)),
                                  ("_bias_momentum", (n,))):
                if getattr(self, name + suffix).shape != shape:
                    raise ValueError("%s%s has shape %s, expected %s"
                                     % (name, suffix, getattr(self, name + suffix).shape, shape))
        self._c = LfmModel(*[f32p(getattr(self, n)) for n in _MODEL_ARRAYS],
                           int(self.item_features.shape[0]), int(self.user_features.shape[0]),
                           d, self.adadelta, self.learning_rate, self.rho, self.eps,
                           self.max_sampled)

    @property
    def ptr(self):
        return C.byref(self._c)


def make_api(lib, prefix, check):
    """Build the eight native-function wrappers over `lib` (symbols `prefix`+name).

    `check(status)` turns a non-zero status into a Python exception.
    """
    last_counters = {}

    def _seeds(random_state, num_threads):
        # Same draw as the reference (T:812-814): advances the caller's RandomState
        # by exactly one randint(size=num_threads) call.
        return random_state.randint(0, np.iinfo(np.int32).max,
                                    size=num_threads).astype(np.uint32)

    def _ids(arr, name):
        return _require(arr, np.int32, 1, name)

    def _flt(arr, name, writable=False):
        return _require(arr, np.float32, 1, name, writable)

    def fit_logistic(item_features, user_features, user_ids, item_ids, Y, sample_weight,
                     shuffle_indices, lightfm, learning_rate, item_alpha, user_alpha,
                     num_threads):
        cnt = LfmCounters()
        check(getattr(lib, prefix + "fit_logistic")(
            item_features.ptr, user_features.ptr, i32p(_ids(user_ids, "user_ids")),
            i32p(_ids(item_ids, "item_ids")), f32p(_flt(Y, "Y")),
            f32p(_flt(sample_weight, "sample_weight")),
            i32p(_ids(shuffle_indices, "shuffle_indices")), len(Y), lightfm.ptr,
            float(item_alpha), float(user_alpha), int(num_threads), C.byref(cnt)))
        last_counters["fit"] = cnt.as_dict()

    def _pairwise(symbol):
        def fit(item_features, user_features, interactions, user_ids, item_ids, Y,
                sample_weight, shuffle_indices, lightfm, learning_rate, item_alpha,
                user_alpha, num_threads, random_state):
            seeds = _seeds(random_state, num_threads)
            cnt = LfmCounters()
            check(getattr(lib, prefix + symbol)(
                item_features.ptr, user_features.ptr, interactions.ptr,
                i32p(_ids(user_ids, "user_ids")), i32p(_ids(item_ids, "item_ids")),
                f32p(_flt(Y, "Y")), f32p(_flt(sample_weight, "sample_weight")),
                i32p(_ids(shuffle_indices, "shuffle_indices")), len(Y), lightfm.ptr,
                float(item_alpha), float(user_alpha), int(num_threads),
                u32p(seeds), len(seeds), C.byref(cnt)))
            last_counters["fit"] = cnt.as_dict()
        fit.__name__ = symbol
        return fit

    def fit_warp_kos(item_features, user_features, data, user_ids, shuffle_indices, lightfm,
                     learning_rate, item_alpha, user_alpha, k, n, num_threads, random_state):
        seeds = _seeds(random_state, num_threads)
        cnt = LfmCounters()
        check(getattr(lib, prefix + "fit_warp_kos")(
            item_features.ptr, user_features.ptr, data.ptr, i32p(_ids(user_ids, "user_ids")),
            i32p(_ids(shuffle_indices, "shuffle_indices")), len(user_ids), lightfm.ptr,
            float(item_alpha), float(user_alpha), int(k), int(n), int(num_threads),
            u32p(seeds), len(seeds), C.byref(cnt)))
        last_counters["fit"] = cnt.as_dict()

    def predict_lightfm(item_features, user_features, user_ids, item_ids, predictions, lightfm,
                        num_threads):
        check(getattr(lib, prefix + "predict_lightfm")(
            item_features.ptr, user_features.ptr, i32p(_ids(user_ids, "user_ids")),
            i32p(_ids(item_ids, "item_ids")),
            f32p(_flt(predictions, "predictions", writable=True)), len(predictions),
            lightfm.ptr, int(num_threads)))

    def predict_ranks(item_features, user_features, test_interactions, train_interactions,
                      ranks, lightfm, num_threads):
        check(getattr(lib, prefix + "predict_ranks")(
            item_features.ptr, user_features.ptr, test_interactions.ptr,
            train_interactions.ptr, f32p(_flt(ranks, "ranks", writable=True)), lightfm.ptr,
            int(num_threads)))

    def calculate_auc_from_rank(ranks, num_train_positives, rank_data, auc, num_threads):
        check(getattr(lib, prefix + "calculate_auc_from_rank")(
            ranks.ptr, i32p(_ids(num_train_positives, "num_train_positives")),
            f32p(_flt(rank_data, "rank_data", writable=True)),
            f32p(_flt(auc, "auc", writable=True)), int(num_threads)))

    def __test_in_positives(row, col, mat):
        r = getattr(lib, prefix + "test_in_positives")(int(row), int(col), mat.ptr)
        if r < 0:
            check(r)
        return bool(r)

    return {
        "fit_logistic": fit_logistic,
        "fit_warp": _pairwise("fit_warp"),
        "fit_bpr": _pairwise("fit_bpr"),
        "fit_warp_kos": fit_warp_kos,
        "predict_lightfm": predict_lightfm,
        "predict_ranks": predict_ranks,
        "calculate_auc_from_rank": calculate_auc_from_rank,
        "__test_in_positives": __test_in_positives,
        "last_counters": last_counters,
    }
