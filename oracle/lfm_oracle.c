/*
 * lfm_oracle.c -- CPU restatement of LightFM's native hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this file.
 * The product (lightfm_b200/) never does; it fails loudly without CUDA.
 *
 * What it is: a plain-C, single-stream restatement of the algorithm in
 * /root/reference/lightfm/_lightfm_fast.pyx.template ("T:" below), following
 * the precision of every temporary as it appears in the Cython-generated C
 * (fp32 tables, fp64 temporaries, fp32 re-rounding at every store).  It
 * reproduces the reference at num_threads=1: same visiting order, same rand_r
 * stream, same per-element op order.  Compile with -ffp-contract=off (the
 * Makefile does) so no FMA contraction changes a rounding.
 *
 * Parity pinned: tests/test_oracle_vs_reference.py checks this file against
 * the real reference built from its own sources into oracle/_ref (bit-equal
 * weights for all four losses / both schedules / with and without features and
 * L2), and tests/golden/ holds vectors generated from that reference build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lfm_cuda.h"

#define MAX_REG_SCALE 1000000.0f /* T:19 (flt) */
#define MAX_LOSS 10.0            /* T:817 */

/* ---- RNG: T:64-90 --------------------------------------------------------- */
static uint32_t temper(uint32_t x) {
    x ^= x >> 11;
    x ^= (x << 7) & 0x9D2C5680u;
    x ^= (x << 15) & 0xEFC60000u;
    x ^= x >> 18;
    return x;
}
static int rand_r_musl(uint32_t *seed) { /* T:79-81 */
    *seed = *seed * 1103515245u + 12345u;
    return (int)(temper(*seed) / 2u);
}
static int sample_range(int min_val, int max_val, uint32_t *seed) { /* T:84-90 */
    return min_val + (rand_r_musl(seed) % (max_val - min_val));
}

/* ---- in_positives: T:270-284 (bsearch over the sorted CSR row) ------------ */
static int in_positives(int item_id, int user_id, const lfm_csr *m) {
    int lo = m->indptr[user_id], hi = m->indptr[user_id + 1];
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        int v = m->indices[mid];
        if (v == item_id) return 1;
        if (v < item_id) lo = mid + 1; else hi = mid;
    }
    return 0;
}

typedef struct {
    lfm_model *m;
    double item_scale, user_scale; /* T:213-214, reset per call T:254-255 */
} state_t;

/* ---- compute_representation: T:287-317 ------------------------------------ */
static void compute_representation(const lfm_csr *features, const float *emb,
                                   const float *biases, int d, int row_id,
                                   double scale, float *repr) {
    int start = features->indptr[row_id], stop = features->indptr[row_id + 1];
    for (int j = 0; j <= d; j++) repr[j] = 0.0f;
    for (int i = start; i < stop; i++) {
        int feature = features->indices[i];
        float fw = (float)((double)features->data[i] * scale); /* T:311 */
        const float *row = emb + (size_t)feature * d;
        for (int j = 0; j < d; j++) {
            float p = fw * row[j];
            repr[j] = repr[j] + p;
        }
        float pb = fw * biases[feature];
        repr[d] = repr[d] + pb;
    }
}

/* ---- compute_prediction_from_repr: T:320-334 ------------------------------ */
static float compute_prediction(const float *u, const float *v, int d) {
    float r = u[d] + v[d];
    for (int i = 0; i < d; i++) {
        float p = u[i] * v[i];
        r = r + p;
    }
    return r;
}

static float sigmoidf_ref(float v) { /* T:262-267 */
    return (float)(1.0 / (1.0 + exp((double)(-v))));
}

/* ---- per-element update, shared by biases (T:337-391) and rows (T:394-451) - */
static double step(float *theta, float *G, float *M, double fw, double gradient,
                   int adadelta, double lr, double alpha, float rho, float eps) {
    double llr;
    if (adadelta) {
        double sq = (fw * gradient) * (fw * gradient);
        float rg = rho * *G; /* float*float, T:364/422 */
        *G = (float)((double)rg + (1.0 - (double)rho) * sq);
        float me = *M + eps, ge = *G + eps;
        llr = sqrt((double)me) / sqrt((double)ge);
        double upd = (llr * gradient) * fw;
        float rm = rho * *M;
        *M = (float)((double)rm + (1.0 - (double)rho) * (upd * upd));
        *theta = (float)((double)*theta - upd);
    } else {
        llr = lr / sqrt((double)*G);
        *theta = (float)((double)*theta - (llr * fw) * gradient);
        double gw = gradient * fw;
        *G = (float)((double)*G + gw * gw);
    }
    *theta = (float)((double)*theta * (1.0 + alpha * llr));
    return llr;
}

static double update_biases(const lfm_csr *f, int start, int stop, float *b, float *g,
                            float *mom, double gradient, const lfm_model *m, double alpha) {
    double sum = 0.0;
    for (int i = start; i < stop; i++) {
        int ft = f->indices[i];
        sum += step(&b[ft], &g[ft], &mom[ft], (double)f->data[i], gradient, m->adadelta,
                    (double)m->learning_rate, alpha, m->rho, m->eps);
    }
    return sum;
}

static double update_features(const lfm_csr *f, float *w, float *g, float *mom, int comp,
                              int start, int stop, double gradient, const lfm_model *m,
                              double alpha) {
    double sum = 0.0;
    int d = m->no_components;
    for (int i = start; i < stop; i++) {
        size_t o = (size_t)f->indices[i] * d + comp;
        sum += step(&w[o], &g[o], &mom[o], (double)f->data[i], gradient, m->adadelta,
                    (double)m->learning_rate, alpha, m->rho, m->eps);
    }
    return sum;
}

/* ---- update: T:454-534 (logistic) ----------------------------------------- */
static void update(state_t *s, double loss, const lfm_csr *itf, const lfm_csr *usf, int user_id,
                   int item_id, const float *u, const float *it, double item_alpha,
                   double user_alpha) {
    lfm_model *m = s->m;
    int d = m->no_components;
    int is = itf->indptr[item_id], ie = itf->indptr[item_id + 1];
    int us = usf->indptr[user_id], ue = usf->indptr[user_id + 1];
    double avg = 0.0;
    avg += update_biases(itf, is, ie, m->item_biases, m->item_bias_gradients,
                         m->item_bias_momentum, loss, m, item_alpha);
    avg += update_biases(usf, us, ue, m->user_biases, m->user_bias_gradients,
                         m->user_bias_momentum, loss, m, user_alpha);
    for (int i = 0; i < d; i++) {
        float uc = u[i], ic = it[i];
        avg += update_features(itf, m->item_features, m->item_feature_gradients,
                               m->item_feature_momentum, i, is, ie, loss * (double)uc, m,
                               item_alpha);
        avg += update_features(usf, m->user_features, m->user_feature_gradients,
                               m->user_feature_momentum, i, us, ue, loss * (double)ic, m,
                               user_alpha);
    }
    avg /= (double)((d + 1) * (ue - us) + (d + 1) * (ie - is));
    s->item_scale *= (1.0 + item_alpha * avg);
    s->user_scale *= (1.0 + user_alpha * avg);
}

/* ---- warp_update: T:537-649 ------------------------------------------------ */
static void warp_update(state_t *s, double loss, const lfm_csr *itf, const lfm_csr *usf,
                        int user_id, int pos_id, int neg_id, const float *u, const float *pos,
                        const float *neg, double item_alpha, double user_alpha) {
    lfm_model *m = s->m;
    int d = m->no_components;
    int ps = itf->indptr[pos_id], pe = itf->indptr[pos_id + 1];
    int ns = itf->indptr[neg_id], ne = itf->indptr[neg_id + 1];
    int us = usf->indptr[user_id], ue = usf->indptr[user_id + 1];
    double avg = 0.0;
    avg += update_biases(itf, ps, pe, m->item_biases, m->item_bias_gradients,
                         m->item_bias_momentum, -loss, m, item_alpha);
    avg += update_biases(itf, ns, ne, m->item_biases, m->item_bias_gradients,
                         m->item_bias_momentum, loss, m, item_alpha);
    avg += update_biases(usf, us, ue, m->user_biases, m->user_bias_gradients,
                         m->user_bias_momentum, loss, m, user_alpha);
    for (int i = 0; i < d; i++) {
        float uc = u[i], pc = pos[i], nc = neg[i];
        avg += update_features(itf, m->item_features, m->item_feature_gradients,
                               m->item_feature_momentum, i, ps, pe, (-loss) * (double)uc, m,
                               item_alpha);
        avg += update_features(itf, m->item_features, m->item_feature_gradients,
                               m->item_feature_momentum, i, ns, ne, loss * (double)uc, m,
                               item_alpha);
        avg += update_features(usf, m->user_features, m->user_feature_gradients,
                               m->user_feature_momentum, i, us, ue,
                               loss * (double)(float)(nc - pc), m, user_alpha);
    }
    avg /= (double)((d + 1) * (ue - us) + (d + 1) * (pe - ps) + (d + 1) * (ne - ns));
    s->item_scale *= (1.0 + item_alpha * avg);
    s->user_scale *= (1.0 + user_alpha * avg);
}

/* ---- regularize: T:652-675 ------------------------------------------------- */
static void regularize(state_t *s) {
    lfm_model *m = s->m;
    int d = m->no_components;
    for (int i = 0; i < m->n_item_features; i++) {
        for (int j = 0; j < d; j++) {
            float *p = &m->item_features[(size_t)i * d + j];
            *p = (float)((double)*p / s->item_scale);
        }
        m->item_biases[i] = (float)((double)m->item_biases[i] / s->item_scale);
    }
    for (int i = 0; i < m->n_user_features; i++) {
        for (int j = 0; j < d; j++) {
            float *p = &m->user_features[(size_t)i * d + j];
            *p = (float)((double)*p / s->user_scale);
        }
        m->user_biases[i] = (float)((double)m->user_biases[i] / s->user_scale);
    }
    s->item_scale = 1.0;
    s->user_scale = 1.0;
}

static void maybe_regularize(state_t *s) { /* T:901-904 + T:678-691 */
    if (s->item_scale > MAX_REG_SCALE || s->user_scale > MAX_REG_SCALE) regularize(s);
}

static void zero_counters(lfm_counters *c) {
    if (c) memset(c, 0, sizeof(*c));
}

/* ---- fit_logistic: T:694-781 ----------------------------------------------- */
int oracle_fit_logistic(const lfm_csr *itf, const lfm_csr *usf, const int32_t *user_ids,
                        const int32_t *item_ids, const float *Y, const float *sample_weight,
                        const int32_t *shuffle, int64_t n, lfm_model *m, double item_alpha,
                        double user_alpha, int32_t num_threads, lfm_counters *c) {
    (void)num_threads;
    state_t s = {m, 1.0, 1.0};
    int d = m->no_components;
    float *u = (float *)malloc(sizeof(float) * (d + 1));
    float *it = (float *)malloc(sizeof(float) * (d + 1));
    zero_counters(c);
    for (int64_t i = 0; i < n; i++) {
        int row = shuffle[i];
        int user_id = user_ids[row], item_id = item_ids[row];
        float weight = sample_weight[row];
        compute_representation(usf, m->user_features, m->user_biases, d, user_id, s.user_scale, u);
        compute_representation(itf, m->item_features, m->item_biases, d, item_id, s.item_scale, it);
        double prediction = (double)sigmoidf_ref(compute_prediction(u, it, d));
        int y = (Y[row] <= 0) ? 0 : 1;
        double loss = (double)weight * (prediction - (double)y);
        update(&s, loss, itf, usf, user_id, item_id, u, it, item_alpha, user_alpha);
        maybe_regularize(&s);
        if (c) { c->positives++; c->updates++; }
    }
    free(u); free(it);
    regularize(&s);
    return 0;
}

/* ---- fit_warp: T:784-912 --------------------------------------------------- */
int oracle_fit_warp(const lfm_csr *itf, const lfm_csr *usf, const lfm_csr *inter,
                    const int32_t *user_ids, const int32_t *item_ids, const float *Y,
                    const float *sample_weight, const int32_t *shuffle, int64_t n,
                    lfm_model *m, double item_alpha, double user_alpha, int32_t num_threads,
                    const uint32_t *random_states, int32_t n_states, lfm_counters *c) {
    (void)num_threads; (void)n_states;
    state_t s = {m, 1.0, 1.0};
    int d = m->no_components;
    uint32_t seed = random_states[0];
    float *u = (float *)malloc(sizeof(float) * (d + 1));
    float *pos = (float *)malloc(sizeof(float) * (d + 1));
    float *neg = (float *)malloc(sizeof(float) * (d + 1));
    zero_counters(c);
    for (int64_t i = 0; i < n; i++) {
        int row = shuffle[i];
        int user_id = user_ids[row], pos_id = item_ids[row];
        if (!(Y[row] > 0)) continue;
        float weight = sample_weight[row];
        if (c) c->positives++;
        compute_representation(usf, m->user_features, m->user_biases, d, user_id, s.user_scale, u);
        compute_representation(itf, m->item_features, m->item_biases, d, pos_id, s.item_scale, pos);
        double pp = (double)compute_prediction(u, pos, d);
        int sampled = 0;
        while (sampled < m->max_sampled) {
            sampled++;
            int neg_id = rand_r_musl(&seed) % itf->rows;
            compute_representation(itf, m->item_features, m->item_biases, d, neg_id,
                                   s.item_scale, neg);
            double np_ = (double)compute_prediction(u, neg, d);
            if (c) c->negatives_drawn++;
            if (np_ > pp - 1) {
                if (in_positives(neg_id, user_id, inter)) { if (c) c->rejected++; continue; }
                double fl = floor((double)((itf->rows - 1) / sampled));
                double loss = (double)weight * log(fl > 1.0 ? fl : 1.0); /* T:881 */
                if (loss > MAX_LOSS) loss = MAX_LOSS;
                warp_update(&s, loss, itf, usf, user_id, pos_id, neg_id, u, pos, neg,
                            item_alpha, user_alpha);
                if (c) c->updates++;
                break;
            }
        }
        maybe_regularize(&s);
    }
    free(u); free(pos); free(neg);
    regularize(&s);
    return 0;
}

/* ---- fit_warp_kos: T:915-1071 ---------------------------------------------- */
typedef struct { int idx; float val; } pair_t; /* T:109-111 */

int oracle_fit_warp_kos(const lfm_csr *itf, const lfm_csr *usf, const lfm_csr *data,
                        const int32_t *user_ids, const int32_t *shuffle, int64_t n_ex,
                        lfm_model *m, double item_alpha, double user_alpha, int32_t k,
                        int32_t n, int32_t num_threads, const uint32_t *random_states,
                        int32_t n_states, lfm_counters *c) {
    (void)num_threads; (void)n_states;
    state_t s = {m, 1.0, 1.0};
    int d = m->no_components;
    uint32_t seed = random_states[0];
    float *u = (float *)malloc(sizeof(float) * (d + 1));
    float *pos = (float *)malloc(sizeof(float) * (d + 1));
    float *neg = (float *)malloc(sizeof(float) * (d + 1));
    pair_t *pairs = (pair_t *)malloc(sizeof(pair_t) * (n > 0 ? n : 1));
    zero_counters(c);
    for (int64_t i = 0; i < n_ex; i++) {
        int row = shuffle[i];
        int user_id = user_ids[row];
        compute_representation(usf, m->user_features, m->user_biases, d, user_id, s.user_scale, u);
        int ps = data->indptr[user_id], pe = data->indptr[user_id + 1];
        if (pe == ps) continue;
        if (c) c->positives++;
        int no_pos = (n < pe - ps) ? n : (pe - ps);
        for (int j = 0; j < no_pos; j++) {
            int sid = data->indices[sample_range(ps, pe, &seed)];
            compute_representation(itf, m->item_features, m->item_biases, d, sid, s.item_scale, pos);
            pairs[j].idx = sid;
            pairs[j].val = compute_prediction(u, pos, d);
        }
        /* qsort(..., reverse_pair_compare) T:997-1000 with glibc's stable merge
         * sort and a comparator that answers "a first" on ties == stable
         * descending insertion sort. */
        for (int a = 1; a < no_pos; a++) {
            pair_t key = pairs[a];
            int b = a - 1;
            while (b >= 0 && (pairs[b].val - key.val) < 0) { pairs[b + 1] = pairs[b]; b--; }
            pairs[b + 1] = key;
        }
        int sel = ((k < no_pos) ? k : no_pos) - 1;
        int pos_id = pairs[sel].idx;
        double pp = (double)pairs[sel].val;
        compute_representation(itf, m->item_features, m->item_biases, d, pos_id, s.item_scale, pos);
        int sampled = 0;
        while (sampled < m->max_sampled) {
            sampled++;
            int neg_id = rand_r_musl(&seed) % itf->rows;
            compute_representation(itf, m->item_features, m->item_biases, d, neg_id,
                                   s.item_scale, neg);
            double np_ = (double)compute_prediction(u, neg, d);
            if (c) c->negatives_drawn++;
            if (np_ > pp - 1) {
                if (in_positives(neg_id, user_id, data)) { if (c) c->rejected++; continue; }
                double loss = log(floor((double)((itf->rows - 1) / sampled))); /* T:1039 */
                if (loss > MAX_LOSS) loss = MAX_LOSS;
                warp_update(&s, loss, itf, usf, user_id, pos_id, neg_id, u, pos, neg,
                            item_alpha, user_alpha);
                if (c) c->updates++;
                break;
            }
        }
        maybe_regularize(&s);
    }
    free(u); free(pos); free(neg); free(pairs);
    regularize(&s);
    return 0;
}

/* ---- fit_bpr: T:1074-1182 -------------------------------------------------- */
int oracle_fit_bpr(const lfm_csr *itf, const lfm_csr *usf, const lfm_csr *inter,
                   const int32_t *user_ids, const int32_t *item_ids, const float *Y,
                   const float *sample_weight, const int32_t *shuffle, int64_t n, lfm_model *m,
                   double item_alpha, double user_alpha, int32_t num_threads,
                   const uint32_t *random_states, int32_t n_states, lfm_counters *c) {
    (void)num_threads; (void)n_states;
    state_t s = {m, 1.0, 1.0};
    int d = m->no_components;
    uint32_t seed = random_states[0];
    float *u = (float *)malloc(sizeof(float) * (d + 1));
    float *pos = (float *)malloc(sizeof(float) * (d + 1));
    float *neg = (float *)malloc(sizeof(float) * (d + 1));
    zero_counters(c);
    for (int64_t i = 0; i < n; i++) {
        int row = shuffle[i];
        if (!(Y[row] > 0)) continue;
        float weight = sample_weight[row];
        int user_id = user_ids[row], pos_id = item_ids[row];
        int neg_id = 0;
        for (int64_t j = 0; j < n; j++) { /* T:1123-1127 */
            neg_id = item_ids[rand_r_musl(&seed) % (int)n];
            if (c) c->negatives_drawn++;
            if (!in_positives(neg_id, user_id, inter)) break;
            if (c) c->rejected++;
        }
        compute_representation(usf, m->user_features, m->user_biases, d, user_id, s.user_scale, u);
        compute_representation(itf, m->item_features, m->item_biases, d, pos_id, s.item_scale, pos);
        compute_representation(itf, m->item_features, m->item_biases, d, neg_id, s.item_scale, neg);
        double pp = (double)compute_prediction(u, pos, d);
        double np_ = (double)compute_prediction(u, neg, d);
        double loss = (double)weight * (1.0 - (double)sigmoidf_ref((float)(pp - np_)));
        warp_update(&s, loss, itf, usf, user_id, pos_id, neg_id, u, pos, neg, item_alpha,
                    user_alpha);
        if (c) { c->positives++; c->updates++; }
        maybe_regularize(&s);
    }
    free(u); free(pos); free(neg);
    regularize(&s);
    return 0;
}

/* ---- predict_lightfm: T:1185-1229 ------------------------------------------ */
int oracle_predict_lightfm(const lfm_csr *itf, const lfm_csr *usf, const int32_t *user_ids,
                           const int32_t *item_ids, float *predictions, int64_t n,
                           const lfm_model *m, int32_t num_threads) {
    (void)num_threads;
    int d = m->no_components;
    float *u = (float *)malloc(sizeof(float) * (d + 1));
    float *it = (float *)malloc(sizeof(float) * (d + 1));
    for (int64_t i = 0; i < n; i++) {
        compute_representation(usf, m->user_features, m->user_biases, d, user_ids[i], 1.0, u);
        compute_representation(itf, m->item_features, m->item_biases, d, item_ids[i], 1.0, it);
        predictions[i] = compute_prediction(u, it, d);
    }
    free(u); free(it);
    return 0;
}

/* ---- predict_ranks: T:1232-1323 -------------------------------------------- */
int oracle_predict_ranks(const lfm_csr *itf, const lfm_csr *usf, const lfm_csr *test,
                         const lfm_csr *train, float *ranks, const lfm_model *m,
                         int32_t num_threads) {
    (void)num_threads;
    int d = m->no_components;
    int maxrow = 0;
    for (int u = 0; u < test->rows; u++) {
        int l = test->indptr[u + 1] - test->indptr[u];
        if (l > maxrow) maxrow = l;
    }
    float *ur = (float *)malloc(sizeof(float) * (d + 1));
    float *ir = (float *)malloc(sizeof(float) * (d + 1));
    int *ids = (int *)malloc(sizeof(int) * (maxrow + 1));
    float *preds = (float *)malloc(sizeof(float) * (maxrow + 1));
    for (int user = 0; user < test->rows; user++) {
        int rs = test->indptr[user], re = test->indptr[user + 1];
        if (re == rs) continue;
        compute_representation(usf, m->user_features, m->user_biases, d, user, 1.0, ur);
        for (int i = 0; i < re - rs; i++) {
            int item = test->indices[rs + i];
            compute_representation(itf, m->item_features, m->item_biases, d, item, 1.0, ir);
            ids[i] = item;
            preds[i] = compute_prediction(ur, ir, d);
        }
        for (int item = 0; item < test->cols; item++) {
            if (in_positives(item, user, train)) continue;
            compute_representation(itf, m->item_features, m->item_biases, d, item, 1.0, ir);
            float p = compute_prediction(ur, ir, d);
            for (int i = 0; i < re - rs; i++)
                if (item != ids[i] && p >= preds[i]) ranks[rs + i] += 1.0f;
        }
    }
    free(ur); free(ir); free(ids); free(preds);
    return 0;
}

/* ---- calculate_auc_from_rank: T:1326-1376 ----------------------------------- */
static int flt_cmp(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x - y > 0) - (x - y < 0);
}
int oracle_calculate_auc_from_rank(const lfm_csr *ranks, const int32_t *num_train_positives,
                                   float *rank_data, float *auc, int32_t num_threads) {
    (void)num_threads;
    for (int user = 0; user < ranks->rows; user++) {
        int rs = ranks->indptr[user], re = ranks->indptr[user + 1];
        int num_pos = re - rs;
        int num_neg = ranks->cols - ((re - rs) + num_train_positives[user]);
        if (num_pos == 0 || num_neg == ranks->cols) { auc[user] = 0.5f; continue; }
        qsort(&rank_data[rs], (size_t)num_pos, sizeof(float), flt_cmp);
        for (int i = 0; i < num_pos; i++) {
            float rank = rank_data[rs + i]; /* ranks.data aliases rank_data, E:247-249 */
            rank = rank - (float)i;
            if (rank < 0) rank = 0;
            /* auc[user] += 1.0 - rank / num_negatives : float/int -> float, 1.0 - float -> double */
            auc[user] = (float)((double)auc[user] + (1.0 - (double)(rank / (float)num_neg)));
        }
        if (num_pos != 0) auc[user] = auc[user] / (float)num_pos;
    }
    return 0;
}

int oracle_test_in_positives(int32_t row, int32_t col, const lfm_csr *mat) {
    return in_positives(col, row, mat);
}
