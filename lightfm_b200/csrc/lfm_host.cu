// lfm_host.cu -- the C ABI of libfm_cuda.so (include/lfm_cuda.h): host-pointer entry
// points that stage inputs into HBM, launch the kernels and write results back in place,
// mirroring the in-place contract of the reference's Cython functions.
//
// No CPU fallback lives here: without a CUDA device every compute entry returns
// LFM_ERR_CUDA with a message.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "lfm_common.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int> g_mode{LFM_MODE_AUTO};
std::atomic<int64_t> g_bitmap_limit_bytes{(int64_t)1 << 30};  // resident plans: membership bitmap if <= 1 GiB
int g_device = 0;
std::mutex g_mu;  // the staging arena is process-global; host entry points serialise on it

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                       \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess)                                                        \
            return fail(e__ == cudaErrorMemoryAllocation ? LFM_ERR_OOM : LFM_ERR_CUDA, \
                        "%s failed: %s", #call, cudaGetErrorString(e__));              \
    } while (0)

// ---- device staging arena: named buffers that only grow -------------------------
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};
typedef std::unordered_map<std::string, Buf> Arena;
Arena g_arena;
Arena* g_cur = &g_arena;  // arena the staging helpers allocate from (global, or a plan's own)
cudaStream_t g_stream = nullptr;
cudaStream_t g_stream2 = nullptr;  // side stream: pack_kernel under the model state's upload (plans)
cudaEvent_t g_ev_side[2] = {nullptr, nullptr};
cudaEvent_t g_ev_prepack = nullptr;  // end of the pack kernel launched ahead for the next epoch (lfm_plan_epoch_next)
cudaEvent_t g_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
bool g_init = false;
bool g_scoring_timed = false;  // g_ev[4..5] bracket the kernels of the last scoring call

int ensure_init() {
    if (g_init) return LFM_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(LFM_ERR_CUDA, "no usable CUDA device (%s); libfm_cuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    CU(cudaSetDevice(g_device));
    CU(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&g_stream2, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&g_ev_side[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g_ev_prepack, cudaEventDisableTiming));
    for (int i = 0; i < 6; i++) CU(cudaEventCreate(&g_ev[i]));
    g_init = true;
    return LFM_OK;
}

int arena_get(const char* name, size_t bytes, void** out) {
    Buf& b = (*g_cur)[name];
    if (bytes == 0) bytes = 16;
    if (b.cap < bytes) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&b.p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            b.p = nullptr;
            return fail(LFM_ERR_OOM, "cudaMalloc(%zu bytes) for '%s' failed: %s", bytes, name,
                        cudaGetErrorString(e));
        }
        b.cap = want;
    }
    *out = b.p;
    return LFM_OK;
}

struct Xfer {  // byte accounting for one call
    int64_t h2d = 0, d2h = 0;
};

template <typename T>
int upload(const char* name, const T* host, size_t count, T** dev, Xfer& x) {
    void* p = nullptr;
    int rc = arena_get(name, count * sizeof(T), &p);
    if (rc != LFM_OK) return rc;
    if (count) {
        if (!host) return fail(LFM_ERR_ARG, "null host pointer for '%s'", name);
        CU(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, g_stream));
        x.h2d += (int64_t)(count * sizeof(T));
    }
    *dev = (T*)p;
    return LFM_OK;
}

template <typename T>
int download(T* host, const T* dev, size_t count, Xfer& x) {
    if (count) {
        CU(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, g_stream));
        x.d2h += (int64_t)(count * sizeof(T));
    }
    return LFM_OK;
}

int check_csr(const lfm_csr* c, const char* what, bool need_data) {
    if (!c) return fail(LFM_ERR_ARG, "%s: null CSR", what);
    if (c->rows < 0 || c->cols < 0 || c->nnz < 0) return fail(LFM_ERR_ARG, "%s: negative size", what);
    if (!c->indptr) return fail(LFM_ERR_ARG, "%s: null indptr", what);
    if (c->nnz > 0 && (!c->indices || (need_data && !c->data)))
        return fail(LFM_ERR_ARG, "%s: null indices/data with nnz > 0", what);
    return LFM_OK;
}

int upload_csr(const char* name, const lfm_csr* c, bool need_data, bool detect_identity, DevCsr* out,
               Xfer& x) {
    int rc = check_csr(c, name, need_data);
    if (rc != LFM_OK) return rc;
    std::string base(name);
    int32_t *indptr = nullptr, *indices = nullptr;
    float* data = nullptr;
    rc = upload((base + ".indptr").c_str(), c->indptr, (size_t)c->rows + 1, &indptr, x);
    if (rc != LFM_OK) return rc;
    rc = upload((base + ".indices").c_str(), c->indices, (size_t)c->nnz, &indices, x);
    if (rc != LFM_OK) return rc;
    if (need_data) {
        rc = upload((base + ".data").c_str(), c->data, (size_t)c->nnz, &data, x);
        if (rc != LFM_OK) return rc;
    }
    out->indptr = indptr;
    out->indices = indices;
    out->data = data;
    out->rows = c->rows;
    out->cols = c->cols;
    out->nnz = c->nnz;
    out->identity = 0;
    if (detect_identity && need_data && c->rows == c->cols && c->nnz == (int64_t)c->rows && c->rows > 0) {
        int32_t* flag = nullptr;
        void* p = nullptr;
        rc = arena_get((base + ".idflag").c_str(), sizeof(int32_t), &p);
        if (rc != LFM_OK) return rc;
        flag = (int32_t*)p;
        int32_t one = 1, h = 0;
        CU(cudaMemcpyAsync(flag, &one, sizeof(one), cudaMemcpyHostToDevice, g_stream));
        CU(lfm_launch_check_identity(*out, flag, g_stream));
        CU(cudaMemcpyAsync(&h, flag, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        out->identity = h;
    }
    return LFM_OK;
}

int check_model(const lfm_model* m) {
    if (!m) return fail(LFM_ERR_ARG, "null model");
    if (m->no_components < 1) return fail(LFM_ERR_ARG, "no_components must be >= 1");
    if (m->n_item_features < 0 || m->n_user_features < 0) return fail(LFM_ERR_ARG, "negative table size");
    const float* req[] = {m->item_features, m->item_feature_gradients, m->item_biases,
                          m->item_bias_gradients, m->user_features, m->user_feature_gradients,
                          m->user_biases, m->user_bias_gradients};
    for (const float* p : req)
        if (!p && (m->n_item_features > 0 && m->n_user_features > 0))
            return fail(LFM_ERR_ARG, "null model array");
    if (m->adadelta && (!m->item_feature_momentum || !m->item_bias_momentum ||
                        !m->user_feature_momentum || !m->user_bias_momentum))
        return fail(LFM_ERR_ARG, "adadelta needs the momentum arrays");
    return LFM_OK;
}

// Upload the model.  `train`: also stage accumulators (and momentum for adadelta).
int upload_model(const lfm_model* m, bool train, DevModel* out, Xfer& x) {
    int rc = check_model(m);
    if (rc != LFM_OK) return rc;
    size_t d = (size_t)m->no_components;
    size_t ni = (size_t)m->n_item_features, nu = (size_t)m->n_user_features;
    memset(out, 0, sizeof(*out));
    out->d = m->no_components;
    out->adadelta = m->adadelta;
    out->lr = m->learning_rate;
    out->rho = m->rho;
    out->eps = m->eps;
    out->max_sampled = m->max_sampled;
    out->item.n = m->n_item_features;
    out->user.n = m->n_user_features;
#define UP(field, src, cnt)                                    \
    rc = upload("model." #field, (const float*)src, cnt, &out->field, x); \
    if (rc != LFM_OK) return rc;
    UP(item.w, m->item_features, ni * d)
    UP(item.b, m->item_biases, ni)
    UP(user.w, m->user_features, nu * d)
    UP(user.b, m->user_biases, nu)
    if (train) {
        UP(item.g, m->item_feature_gradients, ni * d)
        UP(item.bg, m->item_bias_gradients, ni)
        UP(user.g, m->user_feature_gradients, nu * d)
        UP(user.bg, m->user_bias_gradients, nu)
        if (m->adadelta) {
            UP(item.m, m->item_feature_momentum, ni * d)
            UP(item.bm, m->item_bias_momentum, ni)
            UP(user.m, m->user_feature_momentum, nu * d)
            UP(user.bm, m->user_bias_momentum, nu)
        }
    }
#undef UP
    return LFM_OK;
}

int download_model(lfm_model* m, const DevModel& dm, Xfer& x) {
    size_t d = (size_t)m->no_components;
    size_t ni = (size_t)m->n_item_features, nu = (size_t)m->n_user_features;
    int rc;
#define DN(dst, field, cnt)                 \
    rc = download(dst, dm.field, cnt, x);   \
    if (rc != LFM_OK) return rc;
    DN(m->item_features, item.w, ni * d)
    DN(m->item_feature_gradients, item.g, ni * d)
    DN(m->item_biases, item.b, ni)
    DN(m->item_bias_gradients, item.bg, ni)
    DN(m->user_features, user.w, nu * d)
    DN(m->user_feature_gradients, user.g, nu * d)
    DN(m->user_biases, user.b, nu)
    DN(m->user_bias_gradients, user.bg, nu)
    if (m->adadelta) {
        DN(m->item_feature_momentum, item.m, ni * d)
        DN(m->item_bias_momentum, item.bm, ni)
        DN(m->user_feature_momentum, user.m, nu * d)
        DN(m->user_bias_momentum, user.bm, nu)
    }
#undef DN
    return LFM_OK;
}

int resolve_mode(int num_threads, int loss, int d, int nkos) {
    int mode = g_mode;
    if (mode == LFM_MODE_AUTO) mode = (num_threads <= 1) ? LFM_MODE_REPLAY : LFM_MODE_HOGWILD;
    if (mode == LFM_MODE_HOGWILD && !lfm_hogwild_supported(loss, d, nkos)) mode = LFM_MODE_REPLAY;
    return mode;
}

struct FitInputs {
    const lfm_csr *itf, *usf, *pos;
    const int32_t *user_ids, *item_ids;
    const float *y, *w;
    const int32_t* shuffle;
    int64_t n;
    lfm_model* model;
    double item_alpha, user_alpha;
    int k, nkos;
    int num_threads;
    const uint32_t* seeds;
    int n_seeds;
    bool allow_no_positives = false;  // resident plans may answer membership from a COO-built bitmap
};

struct Staged {
    FitArgs a;
    int loss = 0;
    Xfer x;
    std::vector<double> table;
    std::vector<float> table_f;
};

// Hot feature rows (lfm_hogwild.cu): with shared feature rows (tags) the rows touched by a large
// share of the interactions get a slot in the kernels' per-CTA shared-memory accumulators.
// Touch counts come from one pass over the interactions on the device; the selection (at most
// LFM_MAX_HOT rows touched by >= 1/512 of the interactions, most-touched first) is host work on
// n_features floats.
#define LFM_MAX_HOT 160
int select_hot_rows(int loss, FitArgs& a) {
    a.hot_slot_item = a.hot_slot_user = a.hot_rows = nullptr;
    a.n_hot = 0;
    const DevModel& m = a.model;
    if ((a.itf.identity && a.usf.identity) || m.adadelta || a.item_alpha != 0.0 || a.user_alpha != 0.0 ||
        m.d % 4 != 0 || m.d > 128 || a.n < 4096)
        return LFM_OK;
    if (loss == LOSS_KOS && !a.pos.indices) return LFM_OK;
    const size_t ni = (size_t)m.item.n, nu = (size_t)m.user.n;
    void *ci = nullptr, *cu = nullptr;
    int rc = arena_get("hot.cnt_item", sizeof(float) * ni, &ci);
    if (rc != LFM_OK) return rc;
    rc = arena_get("hot.cnt_user", sizeof(float) * nu, &cu);
    if (rc != LFM_OK) return rc;
    const float per_item = a.itf.rows > 0 ? 3.0f * (float)a.n / (float)a.itf.rows : 0.0f;  // ~3 negatives scored per positive
    CU(lfm_launch_feature_counts(a, loss, a.itf.identity ? nullptr : (float*)ci, a.usf.identity ? nullptr : (float*)cu,
                                 per_item, g_stream));
    std::vector<float> hi(a.itf.identity ? 0 : ni), hu(a.usf.identity ? 0 : nu);
    if (!hi.empty()) CU(cudaMemcpyAsync(hi.data(), ci, sizeof(float) * ni, cudaMemcpyDeviceToHost, g_stream));
    if (!hu.empty()) CU(cudaMemcpyAsync(hu.data(), cu, sizeof(float) * nu, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    const float threshold = (float)a.n / 512.0f;
    std::vector<std::pair<float, int32_t>> cand;  // (count, row | user bit)
    for (size_t r = 0; r < hi.size(); r++) if (hi[r] >= threshold) cand.push_back({hi[r], (int32_t)r});
    for (size_t r = 0; r < hu.size(); r++) if (hu[r] >= threshold) cand.push_back({hu[r], (int32_t)(r | 0x80000000u)});
    if (cand.empty()) return LFM_OK;
    // shared memory per slot: 2 rows + one float4 of bias deltas + the lock word; keep ~200 KB
    const size_t per_slot = (size_t)(2 * (m.d / 4) + 1) * 16 + 4;
    size_t max_slots = (200 * 1024) / per_slot;
    if (max_slots > LFM_MAX_HOT) max_slots = LFM_MAX_HOT;
    std::sort(cand.begin(), cand.end(), [](const std::pair<float, int32_t>& x, const std::pair<float, int32_t>& y) {
        return x.first > y.first || (x.first == y.first && x.second < y.second); });
    if (cand.size() > max_slots) cand.resize(max_slots);
    std::vector<int32_t> slot_i(ni, -1), slot_u(nu, -1), rows(cand.size());
    for (size_t s = 0; s < cand.size(); s++) {
        rows[s] = cand[s].second;
        if (cand[s].second < 0) slot_u[(size_t)(cand[s].second & 0x7fffffff)] = (int32_t)s;
        else slot_i[(size_t)cand[s].second] = (int32_t)s;
    }
    Xfer x;
    int32_t *d_si = nullptr, *d_su = nullptr, *d_rows = nullptr;
    rc = upload("hot.slot_item", slot_i.data(), ni, &d_si, x);
    if (rc != LFM_OK) return rc;
    rc = upload("hot.slot_user", slot_u.data(), nu, &d_su, x);
    if (rc != LFM_OK) return rc;
    rc = upload("hot.rows", rows.data(), rows.size(), &d_rows, x);
    if (rc != LFM_OK) return rc;
    CU(cudaStreamSynchronize(g_stream));  // the host vectors die with this frame
    a.hot_slot_item = d_si;
    a.hot_slot_user = d_su;
    a.hot_rows = d_rows;
    a.n_hot = (int32_t)rows.size();
    return LFM_OK;
}

// Validate + upload everything one epoch needs.  `with_shuffle`: stage the host shuffle too.
int stage_fit(int loss, const FitInputs& in, bool with_shuffle, Staged* out) {
    if (in.n < 0) return fail(LFM_ERR_ARG, "negative no_examples");
    if (in.n > 0 && (!in.user_ids || (with_shuffle && !in.shuffle)))
        return fail(LFM_ERR_ARG, "null id / shuffle array");
    if (loss != LOSS_KOS && in.n > 0 && (!in.item_ids || !in.y || !in.w))
        return fail(LFM_ERR_ARG, "null item_ids / Y / sample_weight");
    if (in.n > 0x7fffffffLL) return fail(LFM_ERR_ARG, "no_examples exceeds int32 (reference limit)");
    int rc = check_model(in.model);
    if (rc != LFM_OK) return rc;
    rc = ensure_init();
    if (rc != LFM_OK) return rc;

    Xfer& x = out->x;
    FitArgs& a = out->a;
    memset(&a, 0, sizeof(a));
    out->loss = loss;
    rc = upload_csr("itf", in.itf, true, true, &a.itf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("usf", in.usf, true, true, &a.usf, x);
    if (rc != LFM_OK) return rc;
    if (loss != LOSS_LOGISTIC && in.pos) {
        rc = upload_csr("pos", in.pos, false, false, &a.pos, x);
        if (rc != LFM_OK) return rc;
    } else if (loss != LOSS_LOGISTIC && !in.allow_no_positives) {
        return fail(LFM_ERR_ARG, "pos: null CSR");
    }
    rc = upload_model(in.model, true, &a.model, x);
    if (rc != LFM_OK) return rc;
    if (a.itf.cols > a.model.item.n || a.usf.cols > a.model.user.n)
        return fail(LFM_ERR_ARG, "feature matrix has more columns than the model has embeddings");

    int32_t *d_users = nullptr, *d_items = nullptr, *d_shuffle = nullptr;
    float *d_y = nullptr, *d_w = nullptr;
    rc = upload("fit.user_ids", in.user_ids, (size_t)in.n, &d_users, x);
    if (rc != LFM_OK) return rc;
    if (with_shuffle) {
        rc = upload("fit.shuffle", in.shuffle, (size_t)in.n, &d_shuffle, x);
        if (rc != LFM_OK) return rc;
    }
    if (loss != LOSS_KOS) {
        rc = upload("fit.item_ids", in.item_ids, (size_t)in.n, &d_items, x);
        if (rc != LFM_OK) return rc;
        rc = upload("fit.y", in.y, (size_t)in.n, &d_y, x);
        if (rc != LFM_OK) return rc;
        if (in.w == in.y) {
            d_w = d_y;  // lightfm.py:412-415 aliases sample_weight to interactions.data
        } else {
            rc = upload("fit.w", in.w, (size_t)in.n, &d_w, x);
            if (rc != LFM_OK) return rc;
        }
    }
    if (loss != LOSS_KOS && in.n > 0) {
        // all-ones Y / weights (the common implicit-feedback case): let pack_kernel skip reading them
        void* f = nullptr;
        rc = arena_get("fit.unitflag", sizeof(int32_t), &f);
        if (rc != LFM_OK) return rc;
        int32_t one = 1, hflag = 0;
        CU(cudaMemcpyAsync(f, &one, sizeof(one), cudaMemcpyHostToDevice, g_stream));
        CU(lfm_launch_check_unit(d_y, d_w, in.n, (int32_t*)f, g_stream));
        CU(cudaMemcpyAsync(&hflag, f, sizeof(hflag), cudaMemcpyDeviceToHost, g_stream));
        CU(cudaStreamSynchronize(g_stream));
        a.unit_weights = hflag;
    }
    a.user_ids = d_users;
    a.item_ids = d_items;
    a.y = d_y;
    a.sample_weight = d_w;
    a.shuffle = d_shuffle;
    a.n = in.n;
    a.n_all = in.n;
    a.item_alpha = in.item_alpha;
    a.user_alpha = in.user_alpha;
    a.k = in.k;
    a.nkos = in.nkos;

    // log terms of the WARP loss (T:881 / T:1039) precomputed with the host libm so that
    // replay mode is bit-identical to the reference: loss_table[s] for s in 1..max_sampled
    int ms = a.model.max_sampled > 0 ? a.model.max_sampled : 0;
    out->table.assign((size_t)ms + 1, 0.0);
    for (int s = 1; s <= ms; s++) {
        double fl = floor((double)((a.itf.rows - 1) / s));
        out->table[s] = (loss == LOSS_KOS) ? log(fl) : log(fl > 1.0 ? fl : 1.0);
    }
    double* d_table = nullptr;
    rc = upload("fit.loss_table", out->table.data(), out->table.size(), &d_table, x);
    if (rc != LFM_OK) return rc;
    a.loss_table = d_table;
    out->table_f.assign(out->table.begin(), out->table.end());
    float* d_table_f = nullptr;
    rc = upload("fit.loss_table_f", out->table_f.data(), out->table_f.size(), &d_table_f, x);
    if (rc != LFM_OK) return rc;
    a.loss_table_f = d_table_f;

    rc = select_hot_rows(loss, a);
    if (rc != LFM_OK) return rc;

    void* p = nullptr;
    rc = arena_get("fit.counters", sizeof(DevCounters), &p);
    if (rc != LFM_OK) return rc;
    a.counters = (DevCounters*)p;
    rc = arena_get("fit.scales", sizeof(DevScales), &p);
    if (rc != LFM_OK) return rc;
    a.scales = (DevScales*)p;
    return LFM_OK;
}

// Launch one epoch on staged (device-resident) data.
int run_fit(Staged& st, int mode, uint32_t seed, int* launches, bool pack_aside = false, Tuple* prepacked = nullptr) {
    FitArgs& a = st.a;
    a.seed = seed;
    CU(cudaMemsetAsync(a.counters, 0, sizeof(DevCounters), g_stream));
    DevScales ones = {1.0, 1.0};
    CU(cudaMemcpyAsync(a.scales, &ones, sizeof(ones), cudaMemcpyHostToDevice, g_stream));
    if (mode == LFM_MODE_REPLAY) {
        if (a.n == 0) {
            CU(cudaEventRecord(g_ev[4], g_stream));
            CU(cudaEventRecord(g_ev[5], g_stream));
        }
        if (a.n > 0) {
            if (!a.shuffle) return fail(LFM_ERR_STATE, "replay mode needs the host shuffle order");
            // BPR / logistic with identity features: scratch for the dataflow walk (task list, row
            // versions, membership bitmap); without it the sequential kernels run
            a.replay_scratch = nullptr;
            a.replay_scratch_bytes = lfm_replay_dataflow_scratch_bytes(st.loss, a, g_bitmap_limit_bytes.load());
            if (a.replay_scratch_bytes) {
                int rc = arena_get("replay.scratch", a.replay_scratch_bytes, &a.replay_scratch);
                if (rc != LFM_OK) return rc;
            }
            CU(cudaEventRecord(g_ev[4], g_stream));
            CU(lfm_launch_replay(st.loss, a, g_stream));
            CU(cudaEventRecord(g_ev[5], g_stream));
            (*launches)++;
            DevScales h;
            CU(cudaMemcpyAsync(&h, a.scales, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
            CU(cudaStreamSynchronize(g_stream));
            // final regularize (T:910-912): x / 1.0 == x, so skip the sweep when both are exactly 1
            if (!(h.item_scale == 1.0 && h.user_scale == 1.0)) {
                CU(lfm_launch_regularize(a.model, a.scales, g_stream));
                (*launches)++;
            }
        }
    } else {
        void* p = nullptr;
        int rc = arena_get("fit.tuples", sizeof(Tuple) * (size_t)(a.n > 0 ? a.n : 1), &p);
        if (rc != LFM_OK) return rc;
        // The boundary call (no resident plan): build the exact membership bitmap of the positives for
        // this epoch when it fits -- one sweep over the CSR just uploaded (~0.5 ms at C2) buys the
        // slot kernels a one-load membership test instead of a search per violating negative.
        if (!a.pos_bitmap && a.pos.indptr && st.loss != LOSS_LOGISTIC && g_bitmap_limit_bytes > 0 && a.n >= (1 << 20) &&
            lfm_fast_path_eligible(st.loss, a, a.n)) {
            const int64_t n_cols = a.pos.cols > a.itf.rows ? a.pos.cols : a.itf.rows;
            const int64_t words = (n_cols + 31) / 32;
            const int64_t bytes = words * 4 * (int64_t)a.pos.rows;
            // worth it only when the sweep over the bitmap is small next to the epoch itself
            if (a.pos.rows > 0 && words > 0 && bytes <= g_bitmap_limit_bytes && bytes <= 64 * a.n) {
                void* bm = nullptr;
                rc = arena_get("fit.bitmap", (size_t)bytes, &bm);
                if (rc != LFM_OK) return rc;
                CU(lfm_launch_build_bitmap(a.pos, (uint32_t*)bm, (int32_t)words, g_stream));
                a.pos_bitmap = (const uint32_t*)bm;
                a.bitmap_words = (int32_t)words;
                (*launches)++;
            }
        }
        if (prepacked)
            CU(lfm_launch_hogwild(st.loss, a, prepacked, g_stream, launches, g_ev[4], g_ev[5], nullptr, nullptr,
                                  g_ev_prepack, true));
        else
            CU(lfm_launch_hogwild(st.loss, a, (Tuple*)p, g_stream, launches, g_ev[4], g_ev[5],
                                  pack_aside ? g_stream2 : nullptr, g_ev_side[0], g_ev_side[1]));
    }
    return LFM_OK;
}

uint32_t fold_seed(int loss, int mode, const uint32_t* seeds, int n_seeds) {
    // rand_r stream (replay: thread 0's seed) / philox key (hogwild: all seeds folded together)
    if (loss == LOSS_LOGISTIC || !seeds || n_seeds < 1) return 0x1f2e3d4cu;
    uint32_t seed = seeds[0];
    if (mode == LFM_MODE_HOGWILD)
        for (int i = 1; i < n_seeds; i++) seed = seed * 0x9E3779B1u + seeds[i];
    return seed;
}

void fill_counters(lfm_counters* c, const DevCounters& hc, const Xfer& x, int launches, int mode,
                   float ms_h2d, float ms_k, float ms_d2h) {
    if (!c) return;
    c->positives = (int64_t)hc.positives;
    c->negatives_drawn = (int64_t)hc.negatives;
    c->updates = (int64_t)hc.updates;
    c->rejected = (int64_t)hc.rejected;
    c->kernel_ms = ms_k;
    float ms_train = 0;
    cudaEventElapsedTime(&ms_train, g_ev[4], g_ev[5]);
    c->train_kernel_ms = ms_train;
    c->h2d_ms = ms_h2d;
    c->d2h_ms = ms_d2h;
    c->h2d_bytes = x.h2d;
    c->d2h_bytes = x.d2h;
    c->kernel_launches = launches;
    c->mode = mode;
}

int fit_common(int loss, const FitInputs& in, lfm_counters* counters) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (counters) memset(counters, 0, sizeof(*counters));
    if (loss != LOSS_LOGISTIC && (!in.seeds || in.n_seeds < 1))
        return fail(LFM_ERR_ARG, "random_states must hold at least one seed");
    int rc = ensure_init();
    if (rc != LFM_OK) return rc;
    g_cur = &g_arena;
    Staged st;
    CU(cudaEventRecord(g_ev[0], g_stream));
    rc = stage_fit(loss, in, true, &st);
    if (rc != LFM_OK) return rc;
    const int mode = resolve_mode(in.num_threads, loss, st.a.model.d, in.nkos);
    CU(cudaEventRecord(g_ev[1], g_stream));
    int launches = 0;
    rc = run_fit(st, mode, fold_seed(loss, mode, in.seeds, in.n_seeds), &launches);
    if (rc != LFM_OK) return rc;
    CU(cudaEventRecord(g_ev[2], g_stream));
    rc = download_model(in.model, st.a.model, st.x);
    if (rc != LFM_OK) return rc;
    DevCounters hc;
    CU(cudaMemcpyAsync(&hc, st.a.counters, sizeof(hc), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaEventRecord(g_ev[3], g_stream));
    CU(cudaStreamSynchronize(g_stream));
    float ms_h2d = 0, ms_k = 0, ms_d2h = 0;
    cudaEventElapsedTime(&ms_h2d, g_ev[0], g_ev[1]);
    cudaEventElapsedTime(&ms_k, g_ev[1], g_ev[2]);
    cudaEventElapsedTime(&ms_d2h, g_ev[2], g_ev[3]);
    fill_counters(counters, hc, st.x, launches, mode, ms_h2d, ms_k, ms_d2h);
    return LFM_OK;
}

}  // namespace

// ---- library state ----------------------------------------------------------------
extern "C" const char* lfm_last_error(void) { return g_err; }
extern "C" const char* lfm_version(void) { return "lightfm_b200 libfm_cuda 0.1 (sm_100a)"; }
extern "C" int lfm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" int lfm_set_device(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (device < 0 || device >= lfm_device_count()) return fail(LFM_ERR_ARG, "no such device %d", device);
    if (g_init && device != g_device) return fail(LFM_ERR_STATE, "device already initialised as %d", g_device);
    g_device = device;
    return LFM_OK;
}
extern "C" int lfm_set_mode(int mode) {
    if (mode < LFM_MODE_AUTO || mode > LFM_MODE_HOGWILD) return fail(LFM_ERR_ARG, "bad mode %d", mode);
    g_mode = mode;
    return LFM_OK;
}
extern "C" int lfm_get_mode(void) { return g_mode; }
// Size limit of the positives bitmap a resident plan may build (0 disables it; default 1 GiB).
extern "C" int lfm_set_bitmap_limit(int64_t bytes) {
    g_bitmap_limit_bytes = bytes < 0 ? 0 : bytes;
    return LFM_OK;
}
extern "C" int lfm_release_cache(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_arena)
        if (kv.second.p) cudaFree(kv.second.p);
    g_arena.clear();
    return LFM_OK;
}

// ---- training entry points ----------------------------------------------------------
extern "C" int lfm_fit_logistic(const lfm_csr* item_features, const lfm_csr* user_features,
                                const int32_t* user_ids, const int32_t* item_ids, const float* Y,
                                const float* sample_weight, const int32_t* shuffle_indices,
                                int64_t no_examples, lfm_model* model, double item_alpha,
                                double user_alpha, int32_t num_threads, lfm_counters* counters) {
    FitInputs in = {item_features, user_features, nullptr, user_ids, item_ids, Y, sample_weight,
                    shuffle_indices, no_examples, model, item_alpha, user_alpha, 0, 0, num_threads,
                    nullptr, 0};
    return fit_common(LOSS_LOGISTIC, in, counters);
}

extern "C" int lfm_fit_warp(const lfm_csr* item_features, const lfm_csr* user_features,
                            const lfm_csr* interactions, const int32_t* user_ids,
                            const int32_t* item_ids, const float* Y, const float* sample_weight,
                            const int32_t* shuffle_indices, int64_t no_examples, lfm_model* model,
                            double item_alpha, double user_alpha, int32_t num_threads,
                            const uint32_t* random_states, int32_t n_random_states,
                            lfm_counters* counters) {
    FitInputs in = {item_features, user_features, interactions, user_ids, item_ids, Y, sample_weight,
                    shuffle_indices, no_examples, model, item_alpha, user_alpha, 0, 0, num_threads,
                    random_states, n_random_states};
    return fit_common(LOSS_WARP, in, counters);
}

extern "C" int lfm_fit_bpr(const lfm_csr* item_features, const lfm_csr* user_features,
                           const lfm_csr* interactions, const int32_t* user_ids,
                           const int32_t* item_ids, const float* Y, const float* sample_weight,
                           const int32_t* shuffle_indices, int64_t no_examples, lfm_model* model,
                           double item_alpha, double user_alpha, int32_t num_threads,
                           const uint32_t* random_states, int32_t n_random_states,
                           lfm_counters* counters) {
    FitInputs in = {item_features, user_features, interactions, user_ids, item_ids, Y, sample_weight,
                    shuffle_indices, no_examples, model, item_alpha, user_alpha, 0, 0, num_threads,
                    random_states, n_random_states};
    return fit_common(LOSS_BPR, in, counters);
}

extern "C" int lfm_fit_warp_kos(const lfm_csr* item_features, const lfm_csr* user_features,
                                const lfm_csr* data, const int32_t* user_ids,
                                const int32_t* shuffle_indices, int64_t no_examples, lfm_model* model,
                                double item_alpha, double user_alpha, int32_t k, int32_t n,
                                int32_t num_threads, const uint32_t* random_states,
                                int32_t n_random_states, lfm_counters* counters) {
    if (k < 1 || n < 1) return fail(LFM_ERR_ARG, "k and n must be positive");
    FitInputs in = {item_features, user_features, data, user_ids, nullptr, nullptr, nullptr,
                    shuffle_indices, no_examples, model, item_alpha, user_alpha, k, n, num_threads,
                    random_states, n_random_states};
    return fit_common(LOSS_KOS, in, counters);
}

// ---- scoring entry points -------------------------------------------------------------
extern "C" int lfm_predict_lightfm(const lfm_csr* item_features, const lfm_csr* user_features,
                                   const int32_t* user_ids, const int32_t* item_ids,
                                   float* predictions, int64_t no_examples, const lfm_model* model,
                                   int32_t num_threads) {
    (void)num_threads;
    std::lock_guard<std::mutex> lock(g_mu);
    g_cur = &g_arena;
    if (no_examples < 0) return fail(LFM_ERR_ARG, "negative no_examples");
    if (no_examples > 0 && (!user_ids || !item_ids || !predictions)) return fail(LFM_ERR_ARG, "null array");
    int rc = ensure_init();
    if (rc != LFM_OK) return rc;
    Xfer x;
    DevCsr itf, usf;
    DevModel dm;
    rc = upload_csr("itf", item_features, true, false, &itf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("usf", user_features, true, false, &usf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_model(model, false, &dm, x);
    if (rc != LFM_OK) return rc;
    int32_t *du = nullptr, *di = nullptr;
    rc = upload("predict.user_ids", user_ids, (size_t)no_examples, &du, x);
    if (rc != LFM_OK) return rc;
    rc = upload("predict.item_ids", item_ids, (size_t)no_examples, &di, x);
    if (rc != LFM_OK) return rc;
    void* p = nullptr;
    rc = arena_get("predict.out", sizeof(float) * (size_t)no_examples, &p);
    if (rc != LFM_OK) return rc;
    CU(lfm_launch_predict(itf, usf, dm, du, di, (float*)p, no_examples, g_stream));
    rc = download(predictions, (const float*)p, (size_t)no_examples, x);
    if (rc != LFM_OK) return rc;
    CU(cudaStreamSynchronize(g_stream));
    return LFM_OK;
}

extern "C" int lfm_predict_ranks(const lfm_csr* item_features, const lfm_csr* user_features,
                                 const lfm_csr* test_interactions, const lfm_csr* train_interactions,
                                 float* ranks, const lfm_model* model, int32_t num_threads) {
    (void)num_threads;
    std::lock_guard<std::mutex> lock(g_mu);
    g_cur = &g_arena;
    int rc = check_csr(test_interactions, "test_interactions", false);
    if (rc != LFM_OK) return rc;
    rc = check_csr(train_interactions, "train_interactions", false);
    if (rc != LFM_OK) return rc;
    if (test_interactions->nnz > 0 && !ranks) return fail(LFM_ERR_ARG, "null ranks");
    if (train_interactions->rows < test_interactions->rows)
        return fail(LFM_ERR_ARG, "train_interactions has fewer rows than test_interactions");
    rc = ensure_init();
    if (rc != LFM_OK) return rc;
    Xfer x;
    DevCsr itf, usf, test, train;
    DevModel dm;
    rc = upload_csr("itf", item_features, true, false, &itf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("usf", user_features, true, false, &usf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("ranks.test", test_interactions, false, false, &test, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("ranks.train", train_interactions, false, false, &train, x);
    if (rc != LFM_OK) return rc;
    rc = upload_model(model, false, &dm, x);
    if (rc != LFM_OK) return rc;
    if (itf.rows < test.cols) return fail(LFM_ERR_ARG, "item_features has fewer rows than there are items");
    if (usf.rows < test.rows) return fail(LFM_ERR_ARG, "user_features has fewer rows than there are users");
    float* d_ranks = nullptr;
    rc = upload("ranks.out", ranks, (size_t)test.nnz, &d_ranks, x);
    if (rc != LFM_OK) return rc;
    void* p = nullptr;
    rc = arena_get("ranks.scratch", sizeof(float) * lfm_ranks_scratch_floats(test.cols, dm.d, test.rows), &p);
    if (rc != LFM_OK) return rc;
    int launches = 0;
    CU(cudaEventRecord(g_ev[4], g_stream));
    CU(lfm_launch_predict_ranks(itf, usf, test, train, dm, d_ranks, (float*)p, g_stream, &launches));
    CU(cudaEventRecord(g_ev[5], g_stream));
    g_scoring_timed = true;
    rc = download(ranks, (const float*)d_ranks, (size_t)test.nnz, x);
    if (rc != LFM_OK) return rc;
    CU(cudaStreamSynchronize(g_stream));
    return LFM_OK;
}

extern "C" int lfm_calculate_auc_from_rank(const lfm_csr* ranks, const int32_t* num_train_positives,
                                           float* rank_data, float* auc, int32_t num_threads) {
    (void)num_threads;
    std::lock_guard<std::mutex> lock(g_mu);
    g_cur = &g_arena;
    int rc = check_csr(ranks, "ranks", false);
    if (rc != LFM_OK) return rc;
    if (ranks->rows > 0 && (!num_train_positives || !auc)) return fail(LFM_ERR_ARG, "null array");
    if (ranks->nnz > 0 && !rank_data) return fail(LFM_ERR_ARG, "null rank_data");
    rc = ensure_init();
    if (rc != LFM_OK) return rc;
    Xfer x;
    DevCsr dr;
    rc = upload_csr("auc.ranks", ranks, false, false, &dr, x);
    if (rc != LFM_OK) return rc;
    int32_t* d_ntp = nullptr;
    float *d_rank = nullptr, *d_auc = nullptr;
    rc = upload("auc.ntp", num_train_positives, (size_t)ranks->rows, &d_ntp, x);
    if (rc != LFM_OK) return rc;
    rc = upload("auc.rank_data", (const float*)rank_data, (size_t)ranks->nnz, &d_rank, x);
    if (rc != LFM_OK) return rc;
    rc = upload("auc.auc", (const float*)auc, (size_t)ranks->rows, &d_auc, x);
    if (rc != LFM_OK) return rc;
    void* tmp = nullptr;
    rc = arena_get("auc.tmp", sizeof(float) * (size_t)ranks->nnz, &tmp);
    if (rc != LFM_OK) return rc;
    CU(lfm_launch_auc(dr, d_ntp, d_rank, d_auc, (float*)tmp, g_stream));
    rc = download(rank_data, (const float*)d_rank, (size_t)ranks->nnz, x);
    if (rc != LFM_OK) return rc;
    rc = download(auc, (const float*)d_auc, (size_t)ranks->rows, x);
    if (rc != LFM_OK) return rc;
    CU(cudaStreamSynchronize(g_stream));
    return LFM_OK;
}

// Fused evaluation (lightfm/evaluation.py:14-327): predict_ranks, then the per-user reductions on
// the device -- hits[u] = #{rank < k}, best_rank[u] = smallest rank (-1 for users without test
// interactions), auc[u] as calculate_auc_from_rank computes it (num_train_positives = the train
// rows' lengths).  Any of the three outputs may be NULL.  The nnz_test ranks never leave the GPU.
extern "C" int lfm_evaluate_ranks(const lfm_csr* item_features, const lfm_csr* user_features,
                                  const lfm_csr* test_interactions, const lfm_csr* train_interactions,
                                  const lfm_model* model, int32_t k, int32_t* hits, float* best_rank,
                                  float* auc, int32_t num_threads) {
    (void)num_threads;
    std::lock_guard<std::mutex> lock(g_mu);
    g_cur = &g_arena;
    int rc = check_csr(test_interactions, "test_interactions", false);
    if (rc != LFM_OK) return rc;
    rc = check_csr(train_interactions, "train_interactions", false);
    if (rc != LFM_OK) return rc;
    if (train_interactions->rows < test_interactions->rows)
        return fail(LFM_ERR_ARG, "train_interactions has fewer rows than test_interactions");
    if (k < 0) return fail(LFM_ERR_ARG, "negative k");
    rc = ensure_init();
    if (rc != LFM_OK) return rc;
    Xfer x;
    DevCsr itf, usf, test, train;
    DevModel dm;
    rc = upload_csr("itf", item_features, true, false, &itf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("usf", user_features, true, false, &usf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("ranks.test", test_interactions, false, false, &test, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("ranks.train", train_interactions, false, false, &train, x);
    if (rc != LFM_OK) return rc;
    rc = upload_model(model, false, &dm, x);
    if (rc != LFM_OK) return rc;
    if (itf.rows < test.cols) return fail(LFM_ERR_ARG, "item_features has fewer rows than there are items");
    if (usf.rows < test.rows) return fail(LFM_ERR_ARG, "user_features has fewer rows than there are users");
    const size_t nnz = (size_t)test.nnz, rows = (size_t)test.rows;
    void *p = nullptr, *d_ranks = nullptr, *d_hits = nullptr, *d_best = nullptr, *d_auc = nullptr, *d_ntp = nullptr,
         *d_tmp = nullptr;
    rc = arena_get("ranks.out", sizeof(float) * nnz, &d_ranks);
    if (rc != LFM_OK) return rc;
    rc = arena_get("ranks.scratch", sizeof(float) * lfm_ranks_scratch_floats(test.cols, dm.d, test.rows), &p);
    if (rc != LFM_OK) return rc;
    rc = arena_get("eval.hits", sizeof(int32_t) * rows, &d_hits);
    if (rc != LFM_OK) return rc;
    rc = arena_get("eval.best", sizeof(float) * rows, &d_best);
    if (rc != LFM_OK) return rc;
    rc = arena_get("eval.auc", sizeof(float) * rows, &d_auc);
    if (rc != LFM_OK) return rc;
    rc = arena_get("eval.ntp", sizeof(int32_t) * rows, &d_ntp);
    if (rc != LFM_OK) return rc;
    rc = arena_get("auc.tmp", sizeof(float) * nnz, &d_tmp);
    if (rc != LFM_OK) return rc;
    CU(cudaMemsetAsync(d_ranks, 0, sizeof(float) * nnz, g_stream));
    int launches = 0;
    CU(cudaEventRecord(g_ev[4], g_stream));
    CU(lfm_launch_predict_ranks(itf, usf, test, train, dm, (float*)d_ranks, (float*)p, g_stream, &launches));
    if (hits || best_rank) CU(lfm_launch_rank_metrics(test, (const float*)d_ranks, k, (int32_t*)d_hits, (float*)d_best, g_stream));
    if (auc) {
        CU(lfm_launch_row_counts(train, (int32_t*)d_ntp, test.rows, g_stream));
        CU(cudaMemsetAsync(d_auc, 0, sizeof(float) * rows, g_stream));
        CU(lfm_launch_auc(test, (const int32_t*)d_ntp, (float*)d_ranks, (float*)d_auc, (float*)d_tmp, g_stream));
    }
    CU(cudaEventRecord(g_ev[5], g_stream));
    g_scoring_timed = true;
    if (hits) { rc = download(hits, (const int32_t*)d_hits, rows, x); if (rc != LFM_OK) return rc; }
    if (best_rank) { rc = download(best_rank, (const float*)d_best, rows, x); if (rc != LFM_OK) return rc; }
    if (auc) { rc = download(auc, (const float*)d_auc, rows, x); if (rc != LFM_OK) return rc; }
    CU(cudaStreamSynchronize(g_stream));
    return LFM_OK;
}

// Top-k recommendation for a batch of users (the reference's documented idiom is
// np.argsort(-model.predict(user, np.arange(n_items))), doc/quickstart.rst:125-126): scores of all
// n_items items per user as predict_lightfm computes them (bit-identical), the k best in descending
// score order (ties: lower item id first), optionally skipping the items stored in `exclude`
// (the user's row of a train matrix).  out_items / out_scores are [n_users * k]; slots beyond the
// number of scorable items hold -1 / NaN.  k <= 1024.
extern "C" int lfm_recommend(const lfm_csr* item_features, const lfm_csr* user_features,
                             const lfm_csr* exclude, const int32_t* user_ids, int64_t n_users, int32_t n_items,
                             int32_t k, const lfm_model* model, int32_t* out_items, float* out_scores) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_cur = &g_arena;
    if (n_users < 0 || n_items < 0 || k < 0 || k > 1024) return fail(LFM_ERR_ARG, "bad n_users / n_items / k (k <= 1024)");
    if (n_users > 0 && (!user_ids || (k > 0 && (!out_items || !out_scores)))) return fail(LFM_ERR_ARG, "null array");
    if (n_users == 0 || k == 0 || n_items == 0) return LFM_OK;
    int rc = ensure_init();
    if (rc != LFM_OK) return rc;
    Xfer x;
    DevCsr itf, usf, exc;
    DevModel dm;
    rc = upload_csr("itf", item_features, true, false, &itf, x);
    if (rc != LFM_OK) return rc;
    rc = upload_csr("usf", user_features, true, false, &usf, x);
    if (rc != LFM_OK) return rc;
    if (exclude) {
        rc = upload_csr("rec.exclude", exclude, false, false, &exc, x);
        if (rc != LFM_OK) return rc;
    }
    rc = upload_model(model, false, &dm, x);
    if (rc != LFM_OK) return rc;
    if (itf.rows < n_items) return fail(LFM_ERR_ARG, "item_features has fewer rows than n_items");
    for (int64_t i = 0; i < n_users; i++)
        if (user_ids[i] < 0 || user_ids[i] >= usf.rows) return fail(LFM_ERR_ARG, "user id out of range");
    const int ld = (n_items + 3) & ~3;
    // users are processed in batches whose score rows fit ~2 GiB of scratch
    int64_t batch = ((int64_t)2 << 30) / ((int64_t)ld * 4);
    if (batch < 8) batch = 8;
    if (batch > n_users) batch = n_users;
    void *scratch = nullptr, *d_items = nullptr, *d_scores = nullptr;
    rc = arena_get("rec.scratch", sizeof(float) * ((size_t)ld * (dm.d + 1) + (size_t)batch * ld), &scratch);
    if (rc != LFM_OK) return rc;
    rc = arena_get("rec.items", sizeof(int32_t) * (size_t)batch * k, &d_items);
    if (rc != LFM_OK) return rc;
    rc = arena_get("rec.scores", sizeof(float) * (size_t)batch * k, &d_scores);
    if (rc != LFM_OK) return rc;
    int32_t* d_users = nullptr;
    rc = upload("rec.users", user_ids, (size_t)n_users, &d_users, x);
    if (rc != LFM_OK) return rc;
    CU(cudaEventRecord(g_ev[4], g_stream));
    for (int64_t b0 = 0; b0 < n_users; b0 += batch) {
        const int nb = (int)((n_users - b0 < batch) ? (n_users - b0) : batch);
        int launches = 0;
        CU(lfm_launch_recommend(itf, usf, exclude ? &exc : nullptr, dm, n_items, d_users + b0, nb, k,
                                (int32_t*)d_items, (float*)d_scores, (float*)scratch, g_stream, &launches));
        rc = download(out_items + b0 * k, (const int32_t*)d_items, (size_t)nb * k, x);
        if (rc != LFM_OK) return rc;
        rc = download(out_scores + b0 * k, (const float*)d_scores, (size_t)nb * k, x);
        if (rc != LFM_OK) return rc;
        CU(cudaStreamSynchronize(g_stream));
    }
    CU(cudaEventRecord(g_ev[5], g_stream));
    CU(cudaStreamSynchronize(g_stream));
    g_scoring_timed = true;
    return LFM_OK;
}

// Device time (ms) between the first and last kernel of the most recent predict_ranks /
// evaluate_ranks / recommend call (for recommend: including the per-batch result copies).
extern "C" int lfm_last_scoring_ms(double* ms) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!ms) return fail(LFM_ERR_ARG, "null argument");
    if (!g_init || !g_scoring_timed) return fail(LFM_ERR_STATE, "no scoring call has run yet");
    float f = 0;
    CU(cudaEventElapsedTime(&f, g_ev[4], g_ev[5]));
    *ms = f;
    return LFM_OK;
}

extern "C" int lfm_test_in_positives(int32_t row, int32_t col, const lfm_csr* mat) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_cur = &g_arena;
    int rc = check_csr(mat, "mat", false);
    if (rc != LFM_OK) return rc;
    if (row < 0 || row >= mat->rows) return fail(LFM_ERR_ARG, "row out of range");
    rc = ensure_init();
    if (rc != LFM_OK) return rc;
    Xfer x;
    DevCsr dm;
    rc = upload_csr("tip.mat", mat, false, false, &dm, x);
    if (rc != LFM_OK) return rc;
    void* p = nullptr;
    rc = arena_get("tip.out", sizeof(int32_t), &p);
    if (rc != LFM_OK) return rc;
    CU(lfm_launch_in_positives(dm, row, col, (int32_t*)p, g_stream));
    int32_t h = 0;
    CU(cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    if (h != 0 && h != 3) return fail(LFM_ERR_STATE, "membership searches disagree (%d)", h);
    return h ? 1 : 0;
}


// ---- resident plans: keep one training problem in HBM across epochs ----------------
// (the reference crosses the native boundary once per epoch and
//  re-wraps everything; a plan uploads interactions, features and the model once,
//  runs any number of epochs on the device, and writes the model back on request.)
struct lfm_plan {
    Arena arena;
    Staged st;
    int loss = 0;
    int nkos = 0;
    bool has_shuffle_buf = false;
    bool upload_in_flight = false;  // lfm_plan_upload_model_async: g_ev_side[0] marks the stream position before it
    // lfm_plan_epoch_next: the next epoch's tuples, packed on the side stream while this epoch trains
    bool prepacked = false;
    uint32_t prepacked_seed = 0;
    int64_t prepacked_n = 0;
    Tuple* prepacked_buf = nullptr;
    // delta exchange of a replicated table (lfm_plan_delta_*): per side (0 item, 1 user)
    DeltaSegs segs[2] = {};
    float* dS[2] = {nullptr, nullptr};
    float* dD[2] = {nullptr, nullptr};
    int dstate[2] = {0, 0};  // 0 idle, 1 snapshot taken, 2 delta made
};

extern "C" int lfm_plan_create(lfm_plan** out, int32_t loss, const lfm_csr* item_features,
                               const lfm_csr* user_features, const lfm_csr* interactions,
                               const int32_t* user_ids, const int32_t* item_ids, const float* Y,
                               const float* sample_weight, int64_t no_examples,
                               const lfm_model* model, double item_alpha, double user_alpha,
                               int32_t k, int32_t n) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!out) return fail(LFM_ERR_ARG, "null plan pointer");
    if (loss < LOSS_LOGISTIC || loss > LOSS_KOS) return fail(LFM_ERR_ARG, "bad loss %d", loss);
    int rc = ensure_init();
    if (rc != LFM_OK) return rc;
    // *out != NULL: refresh an existing plan in place (same device buffers, everything re-uploaded)
    const bool reuse = *out != nullptr;
    lfm_plan* p = reuse ? *out : new lfm_plan();
    p->loss = loss;
    p->nkos = n;
    g_cur = &p->arena;
    FitInputs in = {item_features, user_features, interactions, user_ids, item_ids, Y, sample_weight,
                    nullptr, no_examples, const_cast<lfm_model*>(model), item_alpha, user_alpha, k, n,
                    2, nullptr, 0};
    // interactions == NULL (warp / bpr only): no positives CSR at all -- membership comes from a
    // bitmap built on the device straight from the COO arrays, which saves the caller the
    // COO -> sorted CSR conversion (1.3 s of host time per call at 20 M interactions, L:684-686).
    const bool coo_only = interactions == nullptr && (loss == LOSS_WARP || loss == LOSS_BPR);
    in.allow_no_positives = coo_only;
    rc = stage_fit(loss, in, false, &p->st);
    if (rc == LFM_OK && coo_only) {
        FitArgs& a = p->st.a;
        const DevModel& m = a.model;
        const int64_t words = ((int64_t)a.itf.rows + 31) / 32;
        const int64_t bytes = words * 4 * (int64_t)a.usf.rows;
        const bool fast_ok = a.itf.identity && a.usf.identity && !m.adadelta && item_alpha == 0.0 &&
                             user_alpha == 0.0 && (m.d == 16 || m.d == 32 || m.d == 64 || m.d == 128);
        if (!fast_ok || bytes > g_bitmap_limit_bytes || a.usf.rows <= 0 || words <= 0) {
            rc = fail(LFM_ERR_ARG, "a plan without a positives CSR needs the bitmap fast path "
                                   "(identity features, adagrad, no L2, d in {16,32,64,128}, bitmap within the limit)");
        } else {
            void* bm = nullptr;
            rc = arena_get("pos.bitmap", (size_t)bytes, &bm);
            if (rc == LFM_OK) {
                cudaError_t e = lfm_launch_build_bitmap_coo(a.user_ids, a.item_ids, a.n, (uint32_t*)bm, a.usf.rows,
                                                            (int32_t)words, g_stream);
                if (e != cudaSuccess) rc = fail(LFM_ERR_CUDA, "bitmap build failed: %s", cudaGetErrorString(e));
                a.pos_bitmap = (const uint32_t*)bm;
                a.bitmap_words = (int32_t)words;
            }
        }
    }
    if (rc == LFM_OK && loss != LOSS_LOGISTIC && !coo_only && g_bitmap_limit_bytes > 0) {
        // exact membership bitmap of the positives (users x items bits) when it is small enough:
        // replaces the sorted-row search (several L2 sectors per violating negative) by one load
        const DevCsr& pos = p->st.a.pos;
        // negatives are drawn from [0, item_features.rows), which may exceed interactions.shape[1]
        // (lightfm.py:314-363 only requires >=): size the rows for the larger, extra columns stay 0
        const int64_t n_cols = pos.cols > p->st.a.itf.rows ? pos.cols : p->st.a.itf.rows;
        const int64_t words = (n_cols + 31) / 32;
        const int64_t bytes = words * 4 * (int64_t)pos.rows;
        if (pos.rows > 0 && words > 0 && bytes <= g_bitmap_limit_bytes) {
            void* bm = nullptr;
            rc = arena_get("pos.bitmap", (size_t)bytes, &bm);
            if (rc == LFM_OK) {
                cudaError_t e = lfm_launch_build_bitmap(pos, (uint32_t*)bm, (int32_t)words, g_stream);
                if (e != cudaSuccess) rc = fail(LFM_ERR_CUDA, "bitmap build failed: %s", cudaGetErrorString(e));
                p->st.a.pos_bitmap = (const uint32_t*)bm;
                p->st.a.bitmap_words = (int32_t)words;
            }
        }
    }
    if (rc == LFM_OK) {
        cudaError_t e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) rc = fail(LFM_ERR_CUDA, "plan upload failed: %s", cudaGetErrorString(e));
    }
    g_cur = &g_arena;
    if (rc != LFM_OK) {
        if (!reuse) {
            for (auto& kv : p->arena)
                if (kv.second.p) cudaFree(kv.second.p);
            delete p;
        }
        return rc;
    }
    *out = p;
    return LFM_OK;
}

// One epoch on resident data.  shuffle_indices == NULL: the visiting order is a fresh
// pseudo-random permutation generated on the device from `seed` (hogwild mode only).
static int plan_epoch_impl(lfm_plan* p, const int32_t* shuffle_indices, uint32_t seed, int32_t num_threads,
                           const uint32_t* next_seed,
                           int64_t begin, int64_t count, lfm_counters* counters) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p) return fail(LFM_ERR_ARG, "null plan");
    if (counters) memset(counters, 0, sizeof(*counters));
    g_cur = &p->arena;
    struct Restore { ~Restore() { g_cur = &g_arena; } } restore;
    Staged& st = p->st;
    const int64_t n_total = st.a.n_all;
    if (count < 0) count = n_total - begin;
    if (begin < 0 || begin + count > n_total) return fail(LFM_ERR_ARG, "interaction range out of bounds");
    if ((begin != 0 || count != n_total) && shuffle_indices)
        return fail(LFM_ERR_ARG, "a sub-range epoch uses the device-generated order (shuffle_indices must be NULL)");
    struct RangeGuard {  // the staged arguments describe the whole problem again when we leave
        FitArgs& a; int64_t n;
        ~RangeGuard() { a.n = n; a.row_offset = 0; }
    } range_guard{st.a, n_total};
    st.a.n = count;
    st.a.row_offset = begin;
    const int mode = resolve_mode(num_threads, p->loss, st.a.model.d, p->nkos);
    if (mode != LFM_MODE_HOGWILD && (begin != 0 || count != n_total))
        return fail(LFM_ERR_ARG, "a sub-range epoch runs in hogwild mode only");
    if (p->loss != LOSS_LOGISTIC && st.a.pos.indptr == nullptr &&
        (mode != LFM_MODE_HOGWILD || !lfm_fast_path_eligible(p->loss, st.a, st.a.n)))
        return fail(LFM_ERR_STATE, "this plan has no positives CSR (bitmap only): it can only run the "
                                   "hogwild slot kernels, not replay mode or the generic kernels");
    Xfer x;
    CU(cudaEventRecord(g_ev[0], g_stream));
    if (shuffle_indices) {
        int32_t* d = nullptr;
        int rc = upload("fit.shuffle", shuffle_indices, (size_t)st.a.n_all, &d, x);
        if (rc != LFM_OK) return rc;
        st.a.shuffle = d;
    } else {
        if (mode == LFM_MODE_REPLAY) return fail(LFM_ERR_ARG, "replay mode needs shuffle_indices");
        st.a.shuffle = nullptr;
    }
    CU(cudaEventRecord(g_ev[1], g_stream));
    int launches = 0;
    // an asynchronous model upload is still on the stream: let pack_kernel run beside it
    const bool aside = p->upload_in_flight && mode == LFM_MODE_HOGWILD && !shuffle_indices;
    p->upload_in_flight = false;
    const bool whole = begin == 0 && count == n_total;
    // tuples packed ahead for exactly this epoch?  (anything else packed ahead is waited for and dropped)
    Tuple* pre = nullptr;
    if (p->prepacked) {
        if (mode == LFM_MODE_HOGWILD && !shuffle_indices && whole && p->prepacked_seed == seed && p->prepacked_n == st.a.n)
            pre = p->prepacked_buf;
        else
            CU(cudaStreamWaitEvent(g_stream, g_ev_prepack, 0));
        p->prepacked = false;
    }
    int rc = run_fit(st, mode, seed, &launches, aside && !pre, pre);
    if (rc != LFM_OK) return rc;
    if (next_seed && mode == LFM_MODE_HOGWILD && !shuffle_indices && whole && st.a.n > 0) {
        // the next epoch's visiting order only depends on its seed: pack it now, beside this epoch's
        // SGD kernel, into the tuple buffer this epoch is not reading
        void *b0 = nullptr, *b1 = nullptr;
        rc = arena_get("fit.tuples", sizeof(Tuple) * (size_t)st.a.n, &b0);
        if (rc != LFM_OK) return rc;
        rc = arena_get("fit.tuples2", sizeof(Tuple) * (size_t)st.a.n, &b1);
        if (rc != LFM_OK) return rc;
        Tuple* target = (pre == (Tuple*)b1) ? (Tuple*)b0 : (Tuple*)b1;  // this epoch reads `pre`, or b0 when it packed itself
        CU(lfm_launch_pack(st.a, p->loss, target, *next_seed ^ 0x5bd1e995u, g_stream2));
        CU(cudaEventRecord(g_ev_prepack, g_stream2));
        launches++;
        p->prepacked = true;
        p->prepacked_seed = *next_seed;
        p->prepacked_n = st.a.n;
        p->prepacked_buf = target;
    }
    CU(cudaEventRecord(g_ev[2], g_stream));
    DevCounters hc;
    CU(cudaMemcpyAsync(&hc, st.a.counters, sizeof(hc), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaEventRecord(g_ev[3], g_stream));
    CU(cudaStreamSynchronize(g_stream));
    float ms_h2d = 0, ms_k = 0, ms_d2h = 0;
    cudaEventElapsedTime(&ms_h2d, g_ev[0], g_ev[1]);
    cudaEventElapsedTime(&ms_k, g_ev[1], g_ev[2]);
    cudaEventElapsedTime(&ms_d2h, g_ev[2], g_ev[3]);
    fill_counters(counters, hc, x, launches, mode, ms_h2d, ms_k, ms_d2h);
    return LFM_OK;
}

extern "C" int lfm_plan_epoch(lfm_plan* p, const int32_t* shuffle_indices, uint32_t seed,
                              int32_t num_threads, lfm_counters* counters) {
    return plan_epoch_impl(p, shuffle_indices, seed, num_threads, nullptr, 0, -1, counters);
}

// lfm_plan_epoch in the device-generated order, told the seed of the epoch that will follow: that
// epoch's pack kernel runs on a side stream beside this epoch's SGD kernel (into a second tuple
// buffer), and the following lfm_plan_epoch* call with that seed finds its tuples ready.
extern "C" int lfm_plan_epoch_next(lfm_plan* p, uint32_t seed, uint32_t next_seed, int32_t num_threads,
                                   lfm_counters* counters) {
    return plan_epoch_impl(p, nullptr, seed, num_threads, &next_seed, 0, -1, counters);
}

// One pass over interactions [begin, begin + count) of the uploaded list, in a device-generated
// random order (hogwild mode).  Lets a caller cut an epoch into phases (e.g. by user block) and
// interleave its own collectives between them.
extern "C" int lfm_plan_epoch_range(lfm_plan* p, uint32_t seed, int32_t num_threads, int64_t begin,
                                    int64_t count, lfm_counters* counters) {
    return plan_epoch_impl(p, nullptr, seed, num_threads, nullptr, begin, count, counters);
}

// ---- delta exchange of a replicated table (multi-GPU) -------------------------------
// W <- W0 + sum_g (W_g - W0) for rows [row_begin, row_begin + row_count) of one side's w, g, b, bg,
// with the subtract / add-back fused into two sweeps of our own kernels and ONE contiguous buffer
// for the caller's collective (NCCL all-reduce over torch.distributed):
//     lfm_plan_delta_begin   S = cur                       (before the local epoch)
//     lfm_plan_delta_make    D = cur - S, S = D            -> *dev_ptr = D, *count floats: all-reduce it (SUM) in place
//     lfm_plan_delta_apply   cur += D - S                  (what the other ranks did)
// side: 0 item table, 1 user table.
static int timed_delta(int mode, const DeltaSegs& sg, float* S, float* D, double* ms) {
    CU(cudaEventRecord(g_ev[0], g_stream));
    CU(lfm_launch_delta(mode, sg, S, D, g_stream));
    CU(cudaEventRecord(g_ev[1], g_stream));
    CU(cudaStreamSynchronize(g_stream));  // the caller's collective runs on another stream
    if (ms) {
        float f = 0;
        cudaEventElapsedTime(&f, g_ev[0], g_ev[1]);
        *ms = f;
    }
    return LFM_OK;
}

extern "C" int lfm_plan_delta_begin(lfm_plan* p, int32_t side, int64_t row_begin, int64_t row_count, double* ms) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || side < 0 || side > 1) return fail(LFM_ERR_ARG, "bad plan / side");
    const DevModel& m = p->st.a.model;
    const DevTable& t = side == 0 ? m.item : m.user;
    if (row_count < 0) row_count = t.n - row_begin;
    if (row_begin < 0 || row_begin + row_count > t.n) return fail(LFM_ERR_ARG, "row range out of bounds");
    if (m.adadelta) return fail(LFM_ERR_ARG, "the delta exchange covers the adagrad state (w, g, b, bg)");
    g_cur = &p->arena;
    struct Restore { ~Restore() { g_cur = &g_arena; } } restore;
    DeltaSegs& sg = p->segs[side];
    sg.p[0] = t.w + row_begin * m.d;  sg.n[0] = row_count * m.d;
    sg.p[1] = t.g + row_begin * m.d;  sg.n[1] = row_count * m.d;
    sg.p[2] = t.b + row_begin;        sg.n[2] = row_count;
    sg.p[3] = t.bg + row_begin;       sg.n[3] = row_count;
    const size_t total = (size_t)(2 * row_count * m.d + 2 * row_count);
    void *s = nullptr, *d = nullptr;
    int rc = arena_get(side == 0 ? "delta.S.item" : "delta.S.user", sizeof(float) * total, &s);
    if (rc != LFM_OK) return rc;
    rc = arena_get(side == 0 ? "delta.D.item" : "delta.D.user", sizeof(float) * total, &d);
    if (rc != LFM_OK) return rc;
    p->dS[side] = (float*)s;
    p->dD[side] = (float*)d;
    rc = timed_delta(0, sg, p->dS[side], p->dD[side], ms);
    if (rc != LFM_OK) return rc;
    p->dstate[side] = 1;
    return LFM_OK;
}

extern "C" int lfm_plan_delta_make(lfm_plan* p, int32_t side, void** dev_ptr, int64_t* count, double* ms) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || side < 0 || side > 1 || !dev_ptr || !count) return fail(LFM_ERR_ARG, "bad argument");
    if (p->dstate[side] != 1) return fail(LFM_ERR_STATE, "lfm_plan_delta_begin was not called");
    const DeltaSegs& sg = p->segs[side];
    int rc = timed_delta(1, sg, p->dS[side], p->dD[side], ms);
    if (rc != LFM_OK) return rc;
    *dev_ptr = p->dD[side];
    *count = sg.n[0] + sg.n[1] + sg.n[2] + sg.n[3];
    p->dstate[side] = 2;
    return LFM_OK;
}

extern "C" int lfm_plan_delta_apply(lfm_plan* p, int32_t side, double* ms) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || side < 0 || side > 1) return fail(LFM_ERR_ARG, "bad argument");
    if (p->dstate[side] != 2) return fail(LFM_ERR_STATE, "lfm_plan_delta_make was not called");
    int rc = timed_delta(2, p->segs[side], p->dS[side], p->dD[side], ms);
    if (rc != LFM_OK) return rc;
    p->dstate[side] = 0;
    return LFM_OK;
}

// Refresh the resident model state (and the scalar hyper-parameters) from the caller's arrays;
// the interactions, feature matrices and positives lookup stay as uploaded.
static int plan_upload_model_impl(lfm_plan* p, const lfm_model* model, bool wait) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || !model) return fail(LFM_ERR_ARG, "null plan / model");
    const DevModel& dm = p->st.a.model;
    if (model->no_components != dm.d || model->n_item_features != dm.item.n ||
        model->n_user_features != dm.user.n || (model->adadelta != 0) != (dm.adadelta != 0) ||
        model->max_sampled != dm.max_sampled)
        return fail(LFM_ERR_ARG, "model shape does not match the plan");
    g_cur = &p->arena;
    struct Restore { ~Restore() { g_cur = &g_arena; } } restore;
    Xfer x;
    DevModel fresh;
    if (!wait) CU(cudaEventRecord(g_ev_side[0], g_stream));  // everything before the state copies
    int rc = upload_model(model, true, &fresh, x);
    if (rc != LFM_OK) return rc;
    p->st.a.model = fresh;
    if (wait) CU(cudaStreamSynchronize(g_stream));
    p->upload_in_flight = !wait;
    return LFM_OK;
}

extern "C" int lfm_plan_upload_model(lfm_plan* p, const lfm_model* model) { return plan_upload_model_impl(p, model, true); }
// The same without waiting for the copies: the arrays must stay untouched until the next call on
// this plan that synchronises (lfm_plan_epoch*, lfm_plan_download, lfm_plan_check_finite).  The
// next lfm_plan_epoch runs its pack_kernel beside the copies.
extern "C" int lfm_plan_upload_model_async(lfm_plan* p, const lfm_model* model) { return plan_upload_model_impl(p, model, false); }

// Page-lock / unlock caller-owned host memory (cudaHostRegister) so that the copies of a
// long-lived buffer (the model's numpy arrays across fit_partial calls) run at PCIe speed.
extern "C" int lfm_pin_host(void* ptr, int64_t bytes) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!ptr || bytes <= 0) return fail(LFM_ERR_ARG, "bad host range");
    int rc = ensure_init();
    if (rc != LFM_OK) return rc;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return LFM_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LFM_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    return LFM_OK;
}
extern "C" int lfm_unpin_host(void* ptr) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!ptr || !g_init) return LFM_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();  // not registered / context gone: nothing to undo
    return LFM_OK;
}

// Copy the resident model state back into the caller's arrays.
extern "C" int lfm_plan_download(lfm_plan* p, lfm_model* model) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || !model) return fail(LFM_ERR_ARG, "null plan / model");
    const DevModel& dm = p->st.a.model;
    if (model->no_components != dm.d || model->n_item_features != dm.item.n ||
        model->n_user_features != dm.user.n || (model->adadelta != 0) != (dm.adadelta != 0))
        return fail(LFM_ERR_ARG, "model shape does not match the plan");
    Xfer x;
    int rc = download_model(model, dm, x);
    if (rc != LFM_OK) return rc;
    CU(cudaStreamSynchronize(g_stream));
    return LFM_OK;
}

// Divergence check on the device (lightfm.py:447-464 does isfinite(sum(x)) on the host after every
// epoch; with resident state that would cost a full model download per epoch).
__global__ void finite_kernel(const float* a, size_t na, const float* b, size_t nb, const float* c,
                              size_t nc, const float* d, size_t nd, int32_t* flag) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t i = tid; i < na; i += stride) bad |= !isfinite(a[i]);
    for (size_t i = tid; i < nb; i += stride) bad |= !isfinite(b[i]);
    for (size_t i = tid; i < nc; i += stride) bad |= !isfinite(c[i]);
    for (size_t i = tid; i < nd; i += stride) bad |= !isfinite(d[i]);
    if (bad) *flag = 0;
}

extern "C" int lfm_plan_check_finite(lfm_plan* p, int32_t* all_finite) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || !all_finite) return fail(LFM_ERR_ARG, "null argument");
    g_cur = &p->arena;
    struct Restore { ~Restore() { g_cur = &g_arena; } } restore;
    void* f = nullptr;
    int rc = arena_get("finite.flag", sizeof(int32_t), &f);
    if (rc != LFM_OK) return rc;
    const DevModel& m = p->st.a.model;
    int32_t one = 1, h = 0;
    CU(cudaMemcpyAsync(f, &one, sizeof(one), cudaMemcpyHostToDevice, g_stream));
    finite_kernel<<<148 * 8, 256, 0, g_stream>>>(m.item.w, (size_t)m.item.n * m.d, m.item.b, (size_t)m.item.n,
                                                  m.user.w, (size_t)m.user.n * m.d, m.user.b,
                                                  (size_t)m.user.n, (int32_t*)f);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&h, f, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    *all_finite = h;
    return LFM_OK;
}

// Item-sharded multi-GPU runs: negatives are drawn from the local shard but the
// WARP rank estimate floor((n_items - 1) / sampled) keeps the GLOBAL catalogue size.
extern "C" int lfm_plan_set_global_items(lfm_plan* p, int32_t n_items_global) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || n_items_global < 1) return fail(LFM_ERR_ARG, "bad argument");
    Staged& st = p->st;
    int ms = st.a.model.max_sampled > 0 ? st.a.model.max_sampled : 0;
    st.table.assign((size_t)ms + 1, 0.0);
    for (int s = 1; s <= ms; s++) {
        double fl = floor((double)((n_items_global - 1) / s));
        st.table[s] = (p->loss == LOSS_KOS) ? log(fl) : log(fl > 1.0 ? fl : 1.0);
    }
    st.table_f.assign(st.table.begin(), st.table.end());
    CU(cudaMemcpyAsync((void*)st.a.loss_table, st.table.data(), sizeof(double) * st.table.size(),
                       cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync((void*)st.a.loss_table_f, st.table_f.data(), sizeof(float) * st.table_f.size(),
                       cudaMemcpyHostToDevice, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return LFM_OK;
}

// Device address + element count of one resident state array, for callers that run their
// own collectives on it (multi-GPU delta all-reduce).  which: 0..5 item {w,g,m,b,bg,bm},
// 6..11 user {w,g,m,b,bg,bm}.  Arrays the plan does not hold (momentum under adagrad) give NULL.
extern "C" int lfm_plan_table(lfm_plan* p, int32_t which, void** dev_ptr, int64_t* count) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p || !dev_ptr || !count) return fail(LFM_ERR_ARG, "null argument");
    if (which < 0 || which > 11) return fail(LFM_ERR_ARG, "bad table index %d", which);
    const DevModel& dm = p->st.a.model;
    const DevTable& t = which < 6 ? dm.item : dm.user;
    float* ptrs[6] = {t.w, t.g, t.m, t.b, t.bg, t.bm};
    int k = which % 6;
    *dev_ptr = ptrs[k];
    *count = ptrs[k] ? (k < 3 ? (int64_t)t.n * dm.d : (int64_t)t.n) : 0;
    return LFM_OK;
}

extern "C" int lfm_plan_destroy(lfm_plan* p) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!p) return LFM_OK;
    if (g_init) cudaStreamSynchronize(g_stream);
    for (auto& kv : p->arena)
        if (kv.second.p) cudaFree(kv.second.p);
    delete p;
    return LFM_OK;
}
