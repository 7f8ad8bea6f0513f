/*
 * lfm_cuda.h -- C ABI of libfm_cuda.so, the B200 (sm_100a) replacement for the
 * native half of LightFM's fit_partial / predict hot path.
 *
 * Every entry point below replaces one Python-visible function of the
 * reference's Cython extension (lightfm/_lightfm_fast.pyx.template, "T:").
 * A maintainer binds them with ctypes from lightfm/_lightfm_fast.py
 * (see INTEGRATION.md); signatures use plain pointers and sizes only.
 *
 *   T:145-182   cdef class CSRMatrix         -> lfm_csr   (borrowed view)
 *   T:185-259   cdef class FastLightFM       -> lfm_model (borrowed view of the 12 arrays)
 *   T:694-781   fit_logistic                 -> lfm_fit_logistic
 *   T:784-912   fit_warp                     -> lfm_fit_warp
 *   T:915-1071  fit_warp_kos                 -> lfm_fit_warp_kos
 *   T:1074-1182 fit_bpr                      -> lfm_fit_bpr
 *   T:1185-1229 predict_lightfm              -> lfm_predict_lightfm
 *   T:1232-1323 predict_ranks                -> lfm_predict_ranks
 *   T:1326-1376 calculate_auc_from_rank      -> lfm_calculate_auc_from_rank
 *   T:1380-1385 __test_in_positives          -> lfm_test_in_positives
 *
 * Conventions
 *   - All "host" entry points take HOST pointers owned by the caller (numpy /
 *     scipy buffers).  They copy inputs to the GPU, launch the kernels, copy
 *     the mutated arrays back and return when the host buffers are up to date
 *     -- exactly the in-place contract of the Cython functions.
 *   - Return value: LFM_OK (0) or a negative lfm_status; lfm_last_error() gives
 *     the message of the last failure on the calling thread.  There is NO CPU
 *     fallback: without a usable CUDA device every compute entry point returns
 *     LFM_ERR_CUDA.
 *   - num_threads keeps its slot from the reference signature.  On the GPU it
 *     selects the execution mode (overridable with lfm_set_mode):
 *        num_threads == 1  -> LFM_MODE_REPLAY : one sequential stream, the
 *                             reference's exact order, rand_r stream and
 *                             per-element arithmetic (bit-reproducible);
 *        num_threads  > 1  -> LFM_MODE_HOGWILD: thousands of interactions in
 *                             flight over the whole GPU (a slot of 4-32 lanes
 *                             each), lock-free concurrent updates (the
 *                             reference's OpenMP semantics, scaled up).
 */
#ifndef LFM_CUDA_H
#define LFM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    LFM_OK = 0,
    LFM_ERR_ARG = -1,   /* bad argument (null pointer, negative size, shape mismatch) */
    LFM_ERR_CUDA = -2,  /* CUDA runtime / driver failure, or no device */
    LFM_ERR_OOM = -3,   /* device allocation failed */
    LFM_ERR_STATE = -4  /* call order / handle misuse */
} lfm_status;

typedef enum {
    LFM_MODE_AUTO = 0,    /* decide from num_threads (1 -> replay, >1 -> hogwild) */
    LFM_MODE_REPLAY = 1,
    LFM_MODE_HOGWILD = 2
} lfm_mode;

/* Borrowed view of a scipy CSR matrix (T:145-166). `data` may be NULL where the
 * callee only needs the structure (positives lookup, rank structure). */
typedef struct {
    const int32_t *indptr;  /* [rows + 1] */
    const int32_t *indices; /* [nnz]      */
    const float *data;      /* [nnz]      */
    int32_t rows;
    int32_t cols;
    int64_t nnz;
} lfm_csr;

/* Borrowed view of the model state, the "FitModel layout" (T:185-259, built at
 * lightfm.py:422-445).  All arrays are C-contiguous float32. */
typedef struct {
    float *item_features;          /* [n_item_features, no_components] */
    float *item_feature_gradients; /* Adagrad / Adadelta accumulator   */
    float *item_feature_momentum;  /* Adadelta only                    */
    float *item_biases;            /* [n_item_features]                */
    float *item_bias_gradients;
    float *item_bias_momentum;
    float *user_features;          /* [n_user_features, no_components] */
    float *user_feature_gradients;
    float *user_feature_momentum;
    float *user_biases;            /* [n_user_features]                */
    float *user_bias_gradients;
    float *user_bias_momentum;
    int32_t n_item_features;
    int32_t n_user_features;
    int32_t no_components;
    int32_t adadelta;              /* 0 = adagrad, 1 = adadelta        */
    float learning_rate;
    float rho;
    float eps;
    int32_t max_sampled;
} lfm_model;

/* Per-call work counters (algorithmic bytes are computed from them). */
typedef struct {
    int64_t positives;        /* rows actually trained on (Y > 0 for warp/bpr)        */
    int64_t negatives_drawn;  /* S_total: negative items scored                       */
    int64_t updates;          /* U_total: gradient steps applied                      */
    int64_t rejected;         /* violating draws discarded because in_positives       */
    double kernel_ms;         /* device time of all kernels of the call (pack + train + regularize) */
    double train_kernel_ms;   /* device time of the SGD kernel(s) alone (roofline denominator)      */
    double h2d_ms, d2h_ms;    /* copy time inside the call                            */
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches;  /* kernels of this library launched by the call         */
    int32_t mode;             /* lfm_mode actually used                               */
} lfm_counters;

/* ---- library state ------------------------------------------------------- */
const char *lfm_last_error(void);
const char *lfm_version(void);
/* Number of usable CUDA devices (0 on a CPU-only box; never fails). */
int lfm_device_count(void);
/* Select the device used by the host entry points of this thread (default 0). */
int lfm_set_device(int device);
/* Override the num_threads -> mode mapping (LFM_MODE_AUTO restores it). */
int lfm_set_mode(int mode);
int lfm_get_mode(void);
/* Resident plans answer in_positives (T:270-284) from an exact users x items bitmap when it
 * fits in `bytes` of HBM (default 1 GiB; 0 disables it: sorted-row search everywhere). */
int lfm_set_bitmap_limit(int64_t bytes);
/* Free the cached device staging buffers held by the host entry points. */
int lfm_release_cache(void);
/* Page-lock / unlock caller-owned host memory in place (cudaHostRegister), for buffers that
 * outlive one call (the model's state arrays across fit_partial calls).  The caller must
 * unpin before the memory is freed. */
int lfm_pin_host(void *ptr, int64_t bytes);
int lfm_unpin_host(void *ptr);

/* ---- tuning / test knobs (process-wide; each returns the previous value) ------------ */
/* WARP slot-kernel variant: 0 first-generation warp-per-interaction kernel; 4/5 one float4 per
 * lane at 3/4 CTAs per SM; 6/7/8 two float4 per lane at 2/3/4 CTAs per SM; 9/10 = 7/6 with two
 * candidates per slot per round (default 9). */
int lfm_set_tuning(int variant);
/* 0: route fast-path-eligible problems through the generic kernels (tests). */
int lfm_set_fast_path(int enabled);
/* Hogwild launches keep at most max(64, n / divisor) interactions in flight (default 128). */
int lfm_set_inflight_divisor(int divisor);
/* 0: disable the per-CTA shared-memory aggregation of hot feature rows (feature path; tests / A-B timing). */
int lfm_set_hot_rows(int enabled);
/* 0: always use the general replay kernel (the prefetching WARP replay kernel is the default where
 * it applies: identity features, adagrad, alpha == 0); both are bit-equal to the oracle. */
int lfm_set_replay_fast(int enabled);
/* 0: BPR / logistic replay epochs walk the list sequentially (replay_kernel) instead of through the
 * dependency-graph path (many warps, same bits; identity features, alpha == 0); tests / A-B timing. */
int lfm_set_replay_dataflow(int enabled);
/* Device time of the last dataflow replay epoch: its scheduler warp (schedule_ms) and the whole kernel
 * (execute_ms; the two overlap), and the number of tasks. */
int lfm_last_replay_dataflow(double *schedule_ms, double *execute_ms, int32_t *tasks);
/* Hogwild slot kernels: 1 (default) = the Adagrad accumulator update is an atomic add that returns
 * the old value and the step is scaled by it (every earlier update of the element is seen, whatever
 * is in flight); 0 = read-then-reduce (round-1 behaviour). */
int lfm_set_atomic_accumulators(int enabled);
/* predict_ranks tiling: 1 = one 8-user tile per CTA; 3 = three tiles per CTA in lockstep over the
 * item table (L1 sharing).  Returns the previous value. */
int lfm_set_rank_groups(int groups);
/* 1: run the slot kernels as ONE warp with ONE interaction in flight and the reference's rand_r
 * negatives, so that only their arithmetic differs from the oracle (tests/test_gpu_probe.py). */
int lfm_set_probe(int enabled);

/* ---- host entry points: the drop-in boundary ------------------------------ */

/* T:694-781.  no_examples = len(Y).  `counters` may be NULL. */
int lfm_fit_logistic(const lfm_csr *item_features, const lfm_csr *user_features,
                     const int32_t *user_ids, const int32_t *item_ids,
                     const float *Y, const float *sample_weight,
                     const int32_t *shuffle_indices, int64_t no_examples,
                     lfm_model *model, double item_alpha, double user_alpha,
                     int32_t num_threads, lfm_counters *counters);

/* T:784-912.  `random_states` are the per-thread rand_r seeds the reference
 * draws with random_state.randint(0, INT32_MAX, size=num_threads) (T:812-814);
 * replay mode consumes random_states[0] as the rand_r stream, hogwild mode
 * hashes all of them into its Philox key. */
int lfm_fit_warp(const lfm_csr *item_features, const lfm_csr *user_features,
                 const lfm_csr *interactions,
                 const int32_t *user_ids, const int32_t *item_ids,
                 const float *Y, const float *sample_weight,
                 const int32_t *shuffle_indices, int64_t no_examples,
                 lfm_model *model, double item_alpha, double user_alpha,
                 int32_t num_threads, const uint32_t *random_states,
                 int32_t n_random_states, lfm_counters *counters);

/* T:915-1071.  no_examples = len(user_ids). */
int lfm_fit_warp_kos(const lfm_csr *item_features, const lfm_csr *user_features,
                     const lfm_csr *data, const int32_t *user_ids,
                     const int32_t *shuffle_indices, int64_t no_examples,
                     lfm_model *model, double item_alpha, double user_alpha,
                     int32_t k, int32_t n, int32_t num_threads,
                     const uint32_t *random_states, int32_t n_random_states,
                     lfm_counters *counters);

/* T:1074-1182. */
int lfm_fit_bpr(const lfm_csr *item_features, const lfm_csr *user_features,
                const lfm_csr *interactions,
                const int32_t *user_ids, const int32_t *item_ids,
                const float *Y, const float *sample_weight,
                const int32_t *shuffle_indices, int64_t no_examples,
                lfm_model *model, double item_alpha, double user_alpha,
                int32_t num_threads, const uint32_t *random_states,
                int32_t n_random_states, lfm_counters *counters);

/* T:1185-1229.  Fills predictions[0..no_examples). */
int lfm_predict_lightfm(const lfm_csr *item_features, const lfm_csr *user_features,
                        const int32_t *user_ids, const int32_t *item_ids,
                        float *predictions, int64_t no_examples,
                        const lfm_model *model, int32_t num_threads);

/* T:1232-1323.  Accumulates into ranks[0..test.nnz) (caller pre-zeroes, lightfm.py:968-975). */
int lfm_predict_ranks(const lfm_csr *item_features, const lfm_csr *user_features,
                      const lfm_csr *test_interactions, const lfm_csr *train_interactions,
                      float *ranks, const lfm_model *model, int32_t num_threads);

/* T:1326-1376.  Sorts rank_data in place per row, fills auc[0..ranks.rows). */
int lfm_calculate_auc_from_rank(const lfm_csr *ranks, const int32_t *num_train_positives,
                                float *rank_data, float *auc, int32_t num_threads);

/* evaluation.py:14-327 fused behind predict_ranks: the ranks stay on the
 * device; per user, hits[u] = #{rank < k} (precision@k = hits / k, recall@k = hits / #test),
 * best_rank[u] = smallest rank (-1 without test interactions; reciprocal rank = 1 / (best + 1)),
 * auc[u] as calculate_auc_from_rank (T:1326-1376) with num_train_positives = train row lengths.
 * Each output is [test.rows] and may be NULL. */
int lfm_evaluate_ranks(const lfm_csr *item_features, const lfm_csr *user_features,
                       const lfm_csr *test_interactions, const lfm_csr *train_interactions,
                       const lfm_model *model, int32_t k, int32_t *hits, float *best_rank,
                       float *auc, int32_t num_threads);

/* Top-k recommendation (replaces np.argsort(-model.predict(u, arange(n_items))),
 * doc/quickstart.rst:125-126): for every user in user_ids the k best of items [0, n_items) by
 * predict_lightfm's score (bit-identical), descending, ties by ascending item id; items stored in
 * the user's row of `exclude` (may be NULL) are skipped.  Outputs are [n_users * k]; unused slots
 * hold -1 / NaN.  k <= 1024. */
int lfm_recommend(const lfm_csr *item_features, const lfm_csr *user_features, const lfm_csr *exclude,
                  const int32_t *user_ids, int64_t n_users, int32_t n_items, int32_t k,
                  const lfm_model *model, int32_t *out_items, float *out_scores);

/* Device milliseconds spanned by the kernels of the most recent predict_ranks / evaluate_ranks /
 * recommend call (CUDA events on the library's stream). */
int lfm_last_scoring_ms(double *ms);

/* T:1380-1385 (test hook; runs the device membership search). Returns 0/1, <0 on error. */
int lfm_test_in_positives(int32_t row, int32_t col, const lfm_csr *mat);

/* ---- resident plans (epoch prep of lightfm.py:668-759) ------------
 * The reference re-wraps and re-walks every input once per epoch.  A plan uploads the
 * interactions, feature matrices, positives lookup and model state once, runs epochs
 * on the device and copies the model back on request.  `loss`: 0 logistic, 1 warp,
 * 2 bpr, 3 warp-kos.  `interactions` is the sorted positives CSR (NULL for logistic);
 * item_ids / Y / sample_weight are NULL for warp-kos.                                  */
typedef struct lfm_plan lfm_plan;
/* *out must be NULL (new plan) or an existing plan, which is then refreshed in place: all inputs
 * and the model are uploaded again into the plan's existing device buffers. */
int lfm_plan_create(lfm_plan **out, int32_t loss, const lfm_csr *item_features,
                    const lfm_csr *user_features, const lfm_csr *interactions,
                    const int32_t *user_ids, const int32_t *item_ids, const float *Y,
                    const float *sample_weight, int64_t no_examples, const lfm_model *model,
                    double item_alpha, double user_alpha, int32_t k, int32_t n);
/* One epoch.  shuffle_indices == NULL (hogwild only): a fresh device-generated
 * permutation keyed by `seed` replaces the host shuffle of lightfm.py:689-690. */
int lfm_plan_epoch(lfm_plan *plan, const int32_t *shuffle_indices, uint32_t seed,
                   int32_t num_threads, lfm_counters *counters);
/* lfm_plan_epoch (device-generated order) that is also told the seed of the epoch to follow: the
 * following epoch's interaction tuples are packed on a side stream while this epoch trains, and the
 * next lfm_plan_epoch* call with seed == next_seed starts its SGD kernel at once. */
int lfm_plan_epoch_next(lfm_plan *plan, uint32_t seed, uint32_t next_seed, int32_t num_threads,
                        lfm_counters *counters);
/* One pass over interactions [begin, begin + count) of the uploaded list (count < 0: to the end) in
 * a device-generated random order; hogwild mode only.  Lets a caller cut an epoch into phases. */
int lfm_plan_epoch_range(lfm_plan *plan, uint32_t seed, int32_t num_threads, int64_t begin,
                         int64_t count, lfm_counters *counters);
/* Multi-GPU delta exchange of a replicated table: W <- W0 + sum_g (W_g - W0) for rows
 * [row_begin, row_begin + row_count) (row_count < 0: to the end) of one side's w, g, b, bg
 * (side 0 item, 1 user; adagrad state).  begin: snapshot before the local epoch; make: the local
 * delta as ONE contiguous device buffer (*count floats) for the caller's all-reduce (SUM, in place);
 * apply: add what the other ranks did.  The subtract / add-back are two sweeps of this library's
 * kernels; the collective itself is the caller's (NCCL through torch.distributed).  Each call
 * returns when its sweep is complete; *ms (may be NULL) receives the sweep's device time. */
int lfm_plan_delta_begin(lfm_plan *plan, int32_t side, int64_t row_begin, int64_t row_count, double *ms);
int lfm_plan_delta_make(lfm_plan *plan, int32_t side, void **dev_ptr, int64_t *count, double *ms);
int lfm_plan_delta_apply(lfm_plan *plan, int32_t side, double *ms);
int lfm_plan_download(lfm_plan *plan, lfm_model *model);
/* Refresh the resident state arrays and scalar hyper-parameters from `model` (shapes must equal
 * the plan's); interactions, features and the positives lookup stay as uploaded. */
int lfm_plan_upload_model(lfm_plan *plan, const lfm_model *model);
/* The same without waiting for the copies to finish: the arrays must stay untouched until the next
 * call on this plan that synchronises (lfm_plan_epoch*, lfm_plan_download, lfm_plan_check_finite);
 * the next lfm_plan_epoch packs its interaction tuples while the copies are still in flight. */
int lfm_plan_upload_model_async(lfm_plan *plan, const lfm_model *model);
/* Device address and element count of one resident state array (for the caller's own
 * collectives): which = 0..5 item {w,g,m,b,bg,bm}, 6..11 user {w,g,m,b,bg,bm}. */
int lfm_plan_table(lfm_plan *plan, int32_t which, void **dev_ptr, int64_t *count);
/* Device-side version of the per-epoch divergence check (lightfm.py:447-464): *all_finite = 1
 * when every embedding and bias is finite. */
int lfm_plan_check_finite(lfm_plan *plan, int32_t *all_finite);
/* Item-sharded runs: keep the global catalogue size in the WARP rank estimate (T:881). */
int lfm_plan_set_global_items(lfm_plan *plan, int32_t n_items_global);
int lfm_plan_destroy(lfm_plan *plan);

#ifdef __cplusplus
}
#endif
#endif /* LFM_CUDA_H */
