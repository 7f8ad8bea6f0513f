# coding=utf-8
"""The ``LightFM`` model class over the B200 native module.

Public behaviour (constructor arguments and assertions, ``fit`` / ``fit_partial`` /
``predict`` / ``predict_rank`` / ``get_*_representations`` / ``get_params`` /
``set_params`` signatures, input coercion rules, exception types, the twelve float32
state arrays and the consumption order of ``random_state``) follows the reference
class ``/root/reference/lightfm/lightfm.py`` (cited below as ``L:``) so that code
written against ``lightfm.LightFM`` runs unchanged; the implementation is new and
all arithmetic happens in ``libfm_cuda.so`` through ``_lightfm_fast``.

State layout ("FitModel layout", L:281-312): per side (item / user)
``*_embeddings [n_features, no_components]``, ``*_embedding_gradients``,
``*_embedding_momentum``, ``*_biases [n_features]``, ``*_bias_gradients``,
``*_bias_momentum`` -- C-contiguous float32 numpy arrays owned by this object and
mutated in place by every epoch, so pickling, resuming with ``fit_partial`` and
editing the arrays between calls keep working.
"""
from __future__ import print_function

import warnings

import numpy as np
import scipy.sparse as sp

from . import _lightfm_fast as _native

__all__ = ["LightFM"]

CYTHON_DTYPE = np.float32

_STATE = tuple(
    "%s_%s" % (side, part)
    for side in ("item", "user")
    for part in ("embeddings", "embedding_gradients", "embedding_momentum",
                 "biases", "bias_gradients", "bias_momentum")
)
_PARAMS = ("loss", "learning_schedule", "no_components", "learning_rate", "k", "n", "rho",
           "epsilon", "max_sampled", "item_alpha", "user_alpha", "random_state")


def _abi_model_arrays():
    from ._abi import _MODEL_ARRAYS
    return _MODEL_ARRAYS


def _as_float32(mat):
    return mat if mat.dtype == CYTHON_DTYPE else mat.astype(CYTHON_DTYPE)


def _is_identity(csr):
    """True for the identity matrix in canonical CSR form (what None features become)."""
    n = csr.shape[0]
    return (csr.shape[0] == csr.shape[1] and csr.nnz == n
            and np.array_equal(csr.indptr, np.arange(n + 1, dtype=csr.indptr.dtype))
            and np.array_equal(csr.indices, np.arange(n, dtype=csr.indices.dtype))
            and np.array_equiv(csr.data, 1.0))


def _sample(arr):
    """A strided sample of a 1-D array (at most ~4096 elements) used to notice in-place edits."""
    if arr is None:
        return None
    return np.array(arr[::max(1, len(arr) // 4096)])


class _ResidentCache(object):
    """What one throughput-mode ``fit_partial`` call leaves behind for the next one: the resident
    plan (interactions, feature matrices and the membership bitmap / CSR stay in HBM) and the
    page-locked registration of the model's state arrays.  A later call reuses it when it is given
    the same input buffers (identity of the COO / feature / weight arrays, their sizes, and a
    strided sample of their contents) and the same hyper-parameters; anything else rebuilds it."""

    def __init__(self, model, interactions, user_features, item_features, sample_weight):
        self.hyper = model._hyper_key()
        self.shape = interactions.shape
        self.bufs = [interactions.row, interactions.col, interactions.data]
        self.objs = (user_features, item_features, sample_weight)
        for m in self.objs:
            if m is not None:
                self.bufs.append(m.data)
        self.samples = [_sample(b) for b in self.bufs]
        self.plan = None
        self.pins = None
        self.features = None            # (user_features, item_features) CSR float32 as uploaded
        self.sample_weight_data = None

    def matches(self, model, interactions, user_features, item_features, sample_weight):
        if self.plan is None or self.hyper != model._hyper_key() or self.shape != interactions.shape:
            return False
        given = (user_features, item_features, sample_weight)
        if any(a is not b for a, b in zip(given, self.objs)):
            return False
        bufs = [interactions.row, interactions.col, interactions.data] + \
               [m.data for m in given if m is not None]
        if len(bufs) != len(self.bufs) or any(a is not b for a, b in zip(bufs, self.bufs)):
            return False
        return all(np.array_equal(_sample(b), smp) for b, smp in zip(bufs, self.samples))

    def close(self):
        if self.plan is not None:
            self.plan.close()
            self.plan = None
        if self.pins is not None:
            self.pins.release()
            self.pins = None


class LightFM(object):
    """Hybrid latent representation recommender (Kula, 2015) trained on a B200.

    Parameters mirror the reference (L:24-203): ``no_components``, ``k``, ``n``,
    ``learning_schedule`` ('adagrad' | 'adadelta'), ``loss`` ('logistic' | 'bpr' |
    'warp' | 'warp-kos'), ``learning_rate``, ``rho``, ``epsilon``, ``item_alpha``,
    ``user_alpha``, ``max_sampled``, ``random_state``.

    ``num_threads`` in ``fit`` / ``predict`` selects the GPU execution mode instead
    of a CPU thread count: ``1`` replays the reference's single-thread order
    deterministically (bit-reproducible for a fixed seed), ``> 1`` runs the
    lock-free throughput kernels (see ``_lightfm_fast``).
    """

    def __init__(self, no_components=10, k=5, n=10, learning_schedule="adagrad",
                 loss="logistic", learning_rate=0.05, rho=0.95, epsilon=1e-6,
                 item_alpha=0.0, user_alpha=0.0, max_sampled=10, random_state=None):
        # L:205-216
        assert item_alpha >= 0.0
        assert user_alpha >= 0.0
        assert no_components > 0
        assert k > 0
        assert n > 0
        assert 0 < rho < 1
        assert epsilon >= 0
        assert learning_schedule in ("adagrad", "adadelta")
        assert loss in ("logistic", "warp", "bpr", "warp-kos")
        if max_sampled < 1:
            raise ValueError("max_sampled must be a positive integer")

        self.loss = loss
        self.learning_schedule = learning_schedule
        self.no_components = no_components
        self.learning_rate = learning_rate
        self.k = int(k)
        self.n = int(n)
        self.rho = rho
        self.epsilon = epsilon
        self.max_sampled = max_sampled
        self.item_alpha = item_alpha
        self.user_alpha = user_alpha

        if random_state is None:
            self.random_state = np.random.RandomState()
        elif isinstance(random_state, np.random.RandomState):
            self.random_state = random_state
        else:
            self.random_state = np.random.RandomState(random_state)

        self._reset_state()

    # ---- state ----------------------------------------------------------------
    def _reset_state(self):
        for name in _STATE:
            setattr(self, name, None)
        # the resident cache is keyed on the INPUT buffers and survives fit(): the state arrays are
        # uploaded at every call anyway
        self.__dict__.setdefault("_resident_cache", None)

    def release_device(self):
        """Free what throughput-mode training keeps on the GPU between ``fit_partial`` calls
        (the resident plan) and un-pin the state arrays.  Training afterwards re-uploads."""
        cache = self.__dict__.get("_resident_cache")
        if cache is not None:
            cache.close()
        self.__dict__["_resident_cache"] = None

    def __getstate__(self):
        # device handles do not pickle; the twelve numpy arrays are the whole model (as in L:)
        state = dict(self.__dict__)
        state.pop("_resident_cache", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__["_resident_cache"] = None

    def __del__(self):
        try:
            self.release_device()
        except Exception:
            pass

    def _hyper_key(self):
        return (self.loss, self.learning_schedule, self.no_components, float(self.learning_rate),
                self.k, self.n, float(self.rho), float(self.epsilon), int(self.max_sampled),
                float(self.item_alpha), float(self.user_alpha))

    def _check_initialized(self):
        if any(getattr(self, name) is None for name in _STATE):
            raise ValueError("You must fit the model before trying to obtain predictions.")

    def _initialize(self, no_components, no_item_features, no_user_features):
        """Allocate the state arrays (L:281-312).  The item table is drawn before the
        user table; this order is part of the seed contract."""
        ones = self.learning_schedule == "adagrad"
        for side, rows in (("item", no_item_features), ("user", no_user_features)):
            emb = ((self.random_state.rand(rows, no_components) - 0.5) / no_components).astype(np.float32)
            setattr(self, side + "_embeddings", emb)
            setattr(self, side + "_embedding_gradients",
                    np.ones_like(emb) if ones else np.zeros_like(emb))
            setattr(self, side + "_embedding_momentum", np.zeros_like(emb))
            bias = np.zeros(rows, dtype=np.float32)
            setattr(self, side + "_biases", bias)
            setattr(self, side + "_bias_gradients",
                    np.ones_like(bias) if ones else np.zeros_like(bias))
            setattr(self, side + "_bias_momentum", np.zeros_like(bias))

    def _get_lightfm_data(self):
        """Wrap the state for one native call (L:422-445)."""
        return _native.FastLightFM(
            self.item_embeddings, self.item_embedding_gradients, self.item_embedding_momentum,
            self.item_biases, self.item_bias_gradients, self.item_bias_momentum,
            self.user_embeddings, self.user_embedding_gradients, self.user_embedding_momentum,
            self.user_biases, self.user_bias_gradients, self.user_bias_momentum,
            self.no_components, int(self.learning_schedule == "adadelta"),
            self.learning_rate, self.rho, self.epsilon, self.max_sampled)

    # ---- input handling -----------------------------------------------------------
    def _construct_feature_matrices(self, n_users, n_items, user_features, item_features):
        """None -> identity; CSR float32; row / column count checks (L:314-363)."""
        if user_features is None:
            user_features = sp.identity(n_users, dtype=CYTHON_DTYPE, format="csr")
        else:
            user_features = user_features.tocsr()
        if item_features is None:
            item_features = sp.identity(n_items, dtype=CYTHON_DTYPE, format="csr")
        else:
            item_features = item_features.tocsr()

        if n_users > user_features.shape[0]:
            raise Exception("Number of user feature rows does not equal the number of users")
        if n_items > item_features.shape[0]:
            raise Exception("Number of item feature rows does not equal the number of items")

        if self.user_embeddings is not None and \
                not self.user_embeddings.shape[0] >= user_features.shape[1]:
            raise ValueError(
                "The user feature matrix specifies more features than there are estimated "
                "feature embeddings: {} vs {}.".format(self.user_embeddings.shape[0],
                                                       user_features.shape[1]))
        if self.item_embeddings is not None and \
                not self.item_embeddings.shape[0] >= item_features.shape[1]:
            raise ValueError(
                "The item feature matrix specifies more features than there are estimated "
                "feature embeddings: {} vs {}.".format(self.item_embeddings.shape[0],
                                                       item_features.shape[1]))
        return _as_float32(user_features), _as_float32(item_features)

    @staticmethod
    def _positives_lookup(interactions):
        """CSR with sorted column indices, for membership tests (L:365-372)."""
        mat = interactions.tocsr()
        return mat if mat.has_sorted_indices else mat.sorted_indices()

    def _process_sample_weight(self, interactions, sample_weight):
        """L:381-420."""
        if sample_weight is None:
            if np.array_equiv(interactions.data, 1.0):
                return interactions.data  # all ones: share the buffer
            return np.ones_like(interactions.data, dtype=CYTHON_DTYPE)

        if self.loss == "warp-kos":
            raise NotImplementedError("k-OS loss with sample weights not implemented.")
        if not isinstance(sample_weight, sp.coo_matrix):
            raise ValueError("Sample_weight must be a COO matrix.")
        if sample_weight.shape != interactions.shape:
            raise ValueError("Sample weight and interactions matrices must be the same shape")
        if not (np.array_equal(interactions.row, sample_weight.row)
                and np.array_equal(interactions.col, sample_weight.col)):
            raise ValueError("Sample weight and interaction matrix entries must be in the same order")
        if sample_weight.data.dtype != CYTHON_DTYPE:
            return sample_weight.data.astype(CYTHON_DTYPE)
        return sample_weight.data

    @staticmethod
    def _check_input_finite(data):
        if not np.isfinite(np.sum(data)):
            raise ValueError("Not all input values are finite. "
                             "Check the input for NaNs and infinite values.")

    def _check_finite(self):
        """Divergence check after every epoch (L:447-464)."""
        for parameter in (self.item_embeddings, self.item_biases,
                          self.user_embeddings, self.user_biases):
            if not np.isfinite(np.sum(parameter)):
                raise ValueError(
                    "Not all estimated parameters are finite, your model may have diverged. "
                    "Try decreasing the learning rate or normalising feature values and "
                    "sample weights")

    @staticmethod
    def _progress(n, verbose):
        if not verbose:
            return range(n)
        try:
            from tqdm import trange
            return trange(n, desc="Epoch")
        except ImportError:
            def verbose_range():
                for i in range(n):
                    print("Epoch {}".format(i))
                    yield i
            return verbose_range()

    # ---- training --------------------------------------------------------------------
    def fit(self, interactions, user_features=None, item_features=None, sample_weight=None,
            epochs=1, num_threads=1, verbose=False):
        """Fit from scratch (discarding previous state), L:494-558.  Returns self."""
        self._reset_state()
        return self.fit_partial(interactions, user_features=user_features,
                                item_features=item_features, sample_weight=sample_weight,
                                epochs=epochs, num_threads=num_threads, verbose=verbose)

    def fit_partial(self, interactions, user_features=None, item_features=None,
                    sample_weight=None, epochs=1, num_threads=1, verbose=False):
        """Continue training from the current state (L:560-666).  Returns self."""
        interactions = interactions.tocoo()
        if interactions.dtype != CYTHON_DTYPE:
            interactions.data = interactions.data.astype(CYTHON_DTYPE)
        if interactions.row.dtype != np.int32 or interactions.col.dtype != np.int32:
            # scipy switches to int64 indices for very large matrices; the native layer is int32
            # like the reference's (nnz and both dimensions must stay below 2**31)
            if max(interactions.shape) >= 2 ** 31:
                raise ValueError("interaction matrices with a dimension >= 2**31 are not supported")
            interactions = sp.coo_matrix(
                (interactions.data, (interactions.row.astype(np.int32), interactions.col.astype(np.int32))),
                shape=interactions.shape)

        if num_threads < 1:
            raise ValueError("Number of threads must be 1 or larger.")
        resident = epochs > 0 and _native.resolves_to_hogwild(num_threads) and self._resident_ok()
        cache = self.__dict__.get("_resident_cache")
        given = (user_features, item_features, sample_weight)
        reuse = (resident and cache is not None and self.item_embeddings is not None
                 and cache.matches(self, interactions, *given))

        n_users, n_items = interactions.shape
        if reuse:
            # same buffers as the call that built the plan: the O(nnz) host scans below
            # (all-ones / finiteness / feature coercion) were done then and are not repeated
            user_features, item_features = cache.features
            sample_weight_data = cache.sample_weight_data
        else:
            sample_weight_data = self._process_sample_weight(interactions, sample_weight)
            user_features, item_features = self._construct_feature_matrices(
                n_users, n_items, user_features, item_features)
            for input_data in (user_features.data, item_features.data, interactions.data,
                               sample_weight_data):
                self._check_input_finite(input_data)

        if self.item_embeddings is None:
            self._initialize(self.no_components, item_features.shape[1], user_features.shape[1])

        if not item_features.shape[1] == self.item_embeddings.shape[0]:
            raise ValueError("Incorrect number of features in item_features")
        if not user_features.shape[1] == self.user_embeddings.shape[0]:
            raise ValueError("Incorrect number of features in user_features")

        if resident:
            if not reuse:
                self.release_device()
                cache = _ResidentCache(self, interactions, *given)
                cache.features = (user_features, item_features)
                cache.sample_weight_data = sample_weight_data
                self.__dict__["_resident_cache"] = cache
            self._run_epochs_resident(cache, item_features, user_features, interactions,
                                      sample_weight_data, num_threads, epochs, verbose)
            return self

        if epochs > 0 and len(interactions.data) > 200000 and not getattr(LightFM, "_warned_replay", False):
            LightFM._warned_replay = True
            warnings.warn("lightfm_b200: num_threads=1 selects the deterministic replay mode (one "
                          "sequential stream, bit-reproducible, slow); pass num_threads > 1 for the "
                          "GPU throughput kernels.", RuntimeWarning, stacklevel=2)
        for _ in self._progress(epochs, verbose=verbose):
            self._run_epoch(item_features, user_features, interactions, sample_weight_data,
                            num_threads, self.loss)
            self._check_finite()
        return self

    def _resident_ok(self):
        # k-OS with n > 32 falls back to replay inside the library; keep the per-epoch path there
        return self.no_components <= 256 and not (self.loss == "warp-kos" and self.n > 32)

    def _run_epochs_resident(self, cache, item_features, user_features, interactions, sample_weight,
                             num_threads, epochs, verbose):
        """Throughput mode: the problem lives in HBM.  The reference re-builds
        the positives CSR, re-shuffles on the host and re-crosses the boundary with every array
        once per epoch (L:668-759).  Here the interactions, features and membership structure are
        uploaded once per distinct input (``_ResidentCache``), every call uploads the twelve state
        arrays (the numpy arrays stay authoritative between calls: pickling, resuming and editing
        them keep working), runs all epochs on the device and writes the state back once.  Each
        epoch consumes one block of ``random_state.randint`` draws (folded into the key of the
        device-side permutation and of the Philox negative sampler), so the caller's RandomState
        still advances every epoch."""
        state = self._get_lightfm_data()
        arrays = [getattr(state, n) for n in _abi_model_arrays()]
        if self.learning_schedule != "adadelta":
            arrays = [a for n, a in zip(_abi_model_arrays(), arrays) if "momentum" not in n]
        if cache.pins is None or not cache.pins.holds(arrays):
            if cache.pins is not None:
                cache.pins.release()
            cache.pins = _native.PinnedArrays(arrays)
        if cache.plan is None:
            cache.plan = self._make_plan(item_features, user_features, interactions, sample_weight, state)
        else:
            cache.plan.upload_model(state, wait=False)  # the first epoch's pack kernel runs beside the copies
        plan = cache.plan
        finite = True
        try:
            # One seed per epoch, drawn up front (the same draws in the same order as one per
            # iteration: nothing else touches the RandomState in between), so that every epoch can tell
            # the library the next one's seed and have its tuples packed while this one trains.
            # 625 words each: one full Mersenne-Twister block, so get_state()[1] changes every epoch
            # as it does under the reference's per-epoch shuffle (reference tests/test_movielens.py:669-682)
            seeds = []
            for _ in range(epochs):
                words = self.random_state.randint(0, np.iinfo(np.int32).max, size=625)
                seeds.append(int(np.bitwise_xor.reduce(words.astype(np.uint32) * np.uint32(2654435761))))
            for e in self._progress(epochs, verbose=verbose):
                plan.epoch(seeds[e], num_threads=max(2, num_threads),
                           next_seed=seeds[e + 1] if e + 1 < epochs else None)
                finite = plan.all_finite()   # the divergence check of L:447-464, on the device
                if not finite:
                    break
        except Exception:
            self.release_device()
            raise
        plan.download()
        if not finite:
            self._check_finite()

    def _make_plan(self, item_features, user_features, interactions, sample_weight, state):
        pairwise = self.loss in ("warp", "bpr", "warp-kos")
        kos = self.loss == "warp-kos"
        # WARP / BPR on the bitmap fast path: the library builds the membership bitmap on the
        # device straight from the COO arrays, so the COO -> sorted-CSR conversion the reference
        # repeats every epoch (L:684-686, ~1.3 s of host time at 20 M interactions) is not needed
        # at all.  Conditions mirror lfm_plan_create's (sizes are the FEATURE matrices' row counts:
        # negatives are drawn from [0, item_features.shape[0]), L:314-363 allows more rows than items).
        n_urows, n_irows = user_features.shape[0], item_features.shape[0]
        csr_free = (self.loss in ("warp", "bpr") and self.learning_schedule == "adagrad"
                    and self.item_alpha == 0.0 and self.user_alpha == 0.0
                    and self.no_components in (16, 32, 64, 128)
                    and _is_identity(item_features) and _is_identity(user_features)
                    and n_urows * ((n_irows + 31) // 32) * 4 <= _native.bitmap_limit())

        def build(positives):
            return _native.ResidentPlan(
                self.loss, _native.CSRMatrix(item_features), _native.CSRMatrix(user_features), positives,
                interactions.row, None if kos else interactions.col, None if kos else interactions.data,
                None if kos else sample_weight, state, self.item_alpha, self.user_alpha, self.k, self.n)

        if pairwise and csr_free:
            try:
                return build(None)
            except ValueError:
                pass  # the library declined the bitmap-only plan: build the sorted CSR after all
        positives = _native.CSRMatrix(self._positives_lookup(interactions)) if pairwise else None
        return build(positives)

    def _run_epoch(self, item_features, user_features, interactions, sample_weight,
                   num_threads, loss):
        """One pass over the interactions (L:668-759).

        RNG contract kept from the reference: the CSR conversion happens before the
        shuffle, the shuffle consumes ``random_state.shuffle(arange(nnz))``, and the
        pairwise losses then consume ``randint(0, INT32_MAX, size=num_threads)``
        inside the native call.
        """
        pairwise = loss in ("warp", "bpr", "warp-kos")
        if pairwise:
            positives = _native.CSRMatrix(self._positives_lookup(interactions))

        shuffle_indices = np.arange(len(interactions.data), dtype=np.int32)
        self.random_state.shuffle(shuffle_indices)

        state = self._get_lightfm_data()
        item_csr = _native.CSRMatrix(item_features)
        user_csr = _native.CSRMatrix(user_features)

        if loss == "warp":
            _native.fit_warp(item_csr, user_csr, positives, interactions.row, interactions.col,
                             interactions.data, sample_weight, shuffle_indices, state,
                             self.learning_rate, self.item_alpha, self.user_alpha, num_threads,
                             self.random_state)
        elif loss == "bpr":
            _native.fit_bpr(item_csr, user_csr, positives, interactions.row, interactions.col,
                            interactions.data, sample_weight, shuffle_indices, state,
                            self.learning_rate, self.item_alpha, self.user_alpha, num_threads,
                            self.random_state)
        elif loss == "warp-kos":
            _native.fit_warp_kos(item_csr, user_csr, positives, interactions.row, shuffle_indices,
                                 state, self.learning_rate, self.item_alpha, self.user_alpha,
                                 self.k, self.n, num_threads, self.random_state)
        else:
            _native.fit_logistic(item_csr, user_csr, interactions.row, interactions.col,
                                 interactions.data, sample_weight, shuffle_indices, state,
                                 self.learning_rate, self.item_alpha, self.user_alpha, num_threads)

    # ---- scoring -----------------------------------------------------------------------
    def predict(self, user_ids, item_ids, item_features=None, user_features=None, num_threads=1):
        """Scores for (user, item) pairs (L:761-872).  ``user_ids`` may be a single int."""
        self._check_initialized()

        if isinstance(user_ids, int):
            user_ids = np.repeat(np.int32(user_ids), len(item_ids))
        if isinstance(user_ids, (list, tuple)):
            user_ids = np.array(user_ids, dtype=np.int32)
        if isinstance(item_ids, (list, tuple)):
            item_ids = np.array(item_ids, dtype=np.int32)

        if len(user_ids) != len(item_ids):
            raise ValueError(
                f"Expected the number of user IDs ({len(user_ids)}) to equal the number"
                f" of item IDs ({len(item_ids)})")

        if user_ids.dtype != np.int32:
            user_ids = user_ids.astype(np.int32)
        if item_ids.dtype != np.int32:
            item_ids = item_ids.astype(np.int32)

        if num_threads < 1:
            raise ValueError("Number of threads must be 1 or larger.")

        if user_ids.min() < 0 or item_ids.min() < 0:
            raise ValueError("User or item ids cannot be negative. Check your inputs for "
                             "negative numbers or very large numbers that can overflow.")

        n_users = user_ids.max() + 1
        n_items = item_ids.max() + 1
        identity = user_features is None and item_features is None
        user_features, item_features = self._construct_feature_matrices(
            n_users, n_items, user_features, item_features)

        if identity:
            compact = self._predict_compact(user_ids, item_ids, num_threads)
            if compact is not None:
                return compact

        predictions = np.empty(len(user_ids), dtype=np.float32)
        _native.predict_lightfm(_native.CSRMatrix(item_features), _native.CSRMatrix(user_features),
                                np.ascontiguousarray(user_ids), np.ascontiguousarray(item_ids),
                                predictions, self._get_lightfm_data(), num_threads)
        return predictions

    def _predict_compact(self, user_ids, item_ids, num_threads):
        """Small batches with identity features: the native call stages the whole model on the
        device, which dominates when only a few rows are needed (the reference's own tests and
        docs call ``predict`` once per user).  Gather just the rows the batch touches into a
        compact model on the host and score that -- same kernel, same arithmetic, bit-identical
        scores.  Returns None when the batch touches a large part of the tables."""
        uu, ui = np.unique(user_ids, return_inverse=True)
        iu, ii = np.unique(item_ids, return_inverse=True)
        if 4 * (len(uu) + len(iu)) >= self.user_embeddings.shape[0] + self.item_embeddings.shape[0]:
            return None
        d = self.no_components

        def side(emb, bias, ids):
            e = np.ascontiguousarray(emb[ids])
            b = np.ascontiguousarray(bias[ids])
            z, zb = np.zeros_like(e), np.zeros_like(b)
            return [e, z, z.copy(), b, zb, zb.copy()]

        state = _native.FastLightFM(
            *(side(self.item_embeddings, self.item_biases, iu)
              + side(self.user_embeddings, self.user_biases, uu)),
            d, int(self.learning_schedule == "adadelta"), self.learning_rate, self.rho, self.epsilon,
            self.max_sampled)
        predictions = np.empty(len(user_ids), dtype=np.float32)
        _native.predict_lightfm(
            _native.CSRMatrix(sp.identity(len(iu), dtype=CYTHON_DTYPE, format="csr")),
            _native.CSRMatrix(sp.identity(len(uu), dtype=CYTHON_DTYPE, format="csr")),
            np.ascontiguousarray(ui.astype(np.int32)), np.ascontiguousarray(ii.astype(np.int32)),
            predictions, state, num_threads)
        return predictions

    @staticmethod
    def _check_test_train_intersections(test_mat, train_mat):
        if train_mat is not None:
            n_intersections = test_mat.multiply(train_mat).nnz
            if n_intersections:
                raise ValueError(
                    "Test interactions matrix and train interactions matrix share %d "
                    "interactions. This will cause incorrect evaluation, check your data split."
                    % n_intersections)

    def predict_rank(self, test_interactions, train_interactions=None, item_features=None,
                     user_features=None, num_threads=1, check_intersections=True):
        """Rank of every test interaction among all items (L:884-989): 0 is best, train
        positives are excluded from the count, ties count against the test item."""
        self._check_initialized()
        if num_threads < 1:
            raise ValueError("Number of threads must be 1 or larger.")
        if check_intersections:
            self._check_test_train_intersections(test_interactions, train_interactions)

        n_users, n_items = test_interactions.shape
        user_features, item_features = self._construct_feature_matrices(
            n_users, n_items, user_features, item_features)

        if not item_features.shape[1] == self.item_embeddings.shape[0]:
            raise ValueError("Incorrect number of features in item_features")
        if not user_features.shape[1] == self.user_embeddings.shape[0]:
            raise ValueError("Incorrect number of features in user_features")

        test_interactions = _as_float32(test_interactions.tocsr())
        if train_interactions is None:
            train_interactions = sp.csr_matrix((n_users, n_items), dtype=CYTHON_DTYPE)
        else:
            train_interactions = _as_float32(train_interactions.tocsr())

        ranks = sp.csr_matrix(
            (np.zeros_like(test_interactions.data), test_interactions.indices,
             test_interactions.indptr),
            shape=test_interactions.shape)

        _native.predict_ranks(_native.CSRMatrix(item_features), _native.CSRMatrix(user_features),
                              _native.CSRMatrix(test_interactions),
                              _native.CSRMatrix(train_interactions), ranks.data,
                              self._get_lightfm_data(), num_threads)
        return ranks

    def _rank_inputs(self, test_interactions, train_interactions, item_features, user_features,
                     num_threads, check_intersections):
        """Shared argument handling of predict_rank / evaluate_ranks (L:884-989)."""
        self._check_initialized()
        if num_threads < 1:
            raise ValueError("Number of threads must be 1 or larger.")
        if check_intersections:
            self._check_test_train_intersections(test_interactions, train_interactions)
        n_users, n_items = test_interactions.shape
        user_features, item_features = self._construct_feature_matrices(
            n_users, n_items, user_features, item_features)
        if not item_features.shape[1] == self.item_embeddings.shape[0]:
            raise ValueError("Incorrect number of features in item_features")
        if not user_features.shape[1] == self.user_embeddings.shape[0]:
            raise ValueError("Incorrect number of features in user_features")
        test_interactions = _as_float32(test_interactions.tocsr())
        if train_interactions is None:
            train_interactions = sp.csr_matrix((n_users, n_items), dtype=CYTHON_DTYPE)
        else:
            train_interactions = _as_float32(train_interactions.tocsr())
        return test_interactions, train_interactions, item_features, user_features

    def evaluate_ranks(self, test_interactions, train_interactions=None, k=10, item_features=None,
                       user_features=None, num_threads=1, check_intersections=True,
                       hits=True, best_rank=True, auc=True):
        """``predict_rank`` fused with the reductions of ``lightfm.evaluation``:
        returns per-user ``(hits, best_rank, auc)`` -- the number of test items ranked below ``k``,
        the smallest rank (-1 for users without test items) and the AUC exactly as
        ``evaluation.auc_score`` computes it -- without materialising the per-interaction rank
        matrix on the host.  Outputs that are not requested come back as ``None``."""
        test, train, item_features, user_features = self._rank_inputs(
            test_interactions, train_interactions, item_features, user_features, num_threads,
            check_intersections)
        return _native.evaluate_ranks(
            _native.CSRMatrix(item_features), _native.CSRMatrix(user_features), _native.CSRMatrix(test),
            _native.CSRMatrix(train), self._get_lightfm_data(), k, want_hits=hits, want_best=best_rank,
            want_auc=auc, num_threads=num_threads)

    def recommend(self, user_ids, k=10, train_interactions=None, item_features=None, user_features=None,
                  n_items=None):
        """Top-``k`` items for each user.  The reference documents
        ``np.argsort(-model.predict(user_id, np.arange(n_items)))`` (doc/quickstart.rst:125-126);
        this scores the whole catalogue for a batch of users on the device and returns
        ``(items, scores)``, both ``[len(user_ids), k]``: the ``k`` best items in descending score
        order (scores bit-identical to ``predict``; ties by ascending item id), skipping the items
        the user has in ``train_interactions`` when given.  Unused slots hold ``-1`` / ``nan``."""
        self._check_initialized()
        user_ids = np.ascontiguousarray(np.atleast_1d(np.asarray(user_ids)).astype(np.int32))
        if len(user_ids) and user_ids.min() < 0:
            raise ValueError("User ids cannot be negative.")
        if n_items is None:
            if train_interactions is not None:
                n_items = train_interactions.shape[1]
            elif item_features is not None:
                n_items = item_features.shape[0]
            else:
                n_items = self.item_embeddings.shape[0]
        n_users = int(user_ids.max()) + 1 if len(user_ids) else 0
        if train_interactions is not None:
            n_users = max(n_users, train_interactions.shape[0])
        user_features, item_features = self._construct_feature_matrices(
            n_users, n_items, user_features, item_features)
        exclude = None
        if train_interactions is not None:
            exclude = _native.CSRMatrix(_as_float32(train_interactions.tocsr()))
        return _native.recommend(_native.CSRMatrix(item_features), _native.CSRMatrix(user_features),
                                 exclude, user_ids, n_items, k, self._get_lightfm_data())

    # ---- representations / sklearn plumbing -------------------------------------------------
    def get_item_representations(self, features=None):
        """(biases, embeddings) of items, optionally projected through ``features`` (L:991-1018)."""
        self._check_initialized()
        if features is None:
            return self.item_biases, self.item_embeddings
        features = sp.csr_matrix(features, dtype=CYTHON_DTYPE)
        return features * self.item_biases, features * self.item_embeddings

    def get_user_representations(self, features=None):
        """(biases, embeddings) of users, optionally projected through ``features`` (L:1020-1047)."""
        self._check_initialized()
        if features is None:
            return self.user_biases, self.user_embeddings
        features = sp.csr_matrix(features, dtype=CYTHON_DTYPE)
        return features * self.user_biases, features * self.user_embeddings

    def get_params(self, deep=True):
        """Constructor parameters, sklearn style (L:1049-1082)."""
        return {name: getattr(self, name) for name in _PARAMS}

    def set_params(self, **params):
        """Set constructor parameters, sklearn style (L:1084-1107)."""
        valid_params = self.get_params()
        for key, value in params.items():
            if key not in valid_params:
                raise ValueError(
                    "Invalid parameter %s for estimator %s. Check the list of available "
                    "parameters with `estimator.get_params().keys()`."
                    % (key, self.__class__.__name__))
            setattr(self, key, value)
        return self
