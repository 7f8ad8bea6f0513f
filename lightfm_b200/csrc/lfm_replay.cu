// lfm_replay.cu -- deterministic replay mode (num_threads == 1).
//
// One warp walks the shuffled interaction list in the reference's exact order
// with the reference's rand_r stream and per-element arithmetic, so the result
// equals the reference at num_threads=1 (oracle/lfm_oracle.c
// is the CPU statement of the same rules).  Lanes own embedding components
// (lane l owns components l, l+32, ...): within one SGD step the reference
// visits component i of the positive row, then the negative row, then the user
// row, and different components never touch the same address, so running the
// components on different lanes preserves every per-address op order.
//
// THIS FILE MUST BE COMPILED WITH --fmad=false: an FMA would remove a rounding
// the reference performs.
//
// Reference: lightfm/_lightfm_fast.pyx.template ("T:")
//   compute_representation T:287-317, compute_prediction_from_repr T:320-334,
//   update_biases T:337-391, update_features T:394-451, update T:454-534,
//   warp_update T:537-649, regularize T:652-675, fit_logistic T:694-781,
//   fit_warp T:784-912, fit_warp_kos T:915-1071, fit_bpr T:1074-1182.
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "lfm_common.cuh"

namespace {

struct ReplayShared {
    float* u;       // [d+1]
    float* pos;     // [d+1]
    float* neg;     // [d+1]
    double* lrsum;  // [3*d] per-component learning-rate sums (pos/item, neg, user)
    int* pidx;      // [nkos]
    float* pval;    // [nkos]
};

// T:287-317.  repr[j] = f32(repr[j] + f32(fw * E[f,j])), features in CSR order.
// Lane l owns components l, l+32, ...; lane 0 also owns the bias slot repr[d]
// (it is the lane that applies bias steps, so no cross-lane global dependency).
__device__ void gather(const DevCsr& f, const float* emb, const float* bias, int d, int row,
                       double scale, float* repr, int lane) {
    __syncwarp();  // previous readers of repr are done
    if (f.identity) {
        // identity features: the row has the single entry (row, 1.0f).  Same operations as the
        // loop below -- fw = f32(double(1.0f) * scale), repr = 0.0f + fw * E -- without the three
        // dependent CSR loads.
        const float fw = (float)((double)1.0f * scale);
        const float* r = emb + (size_t)row * d;
        for (int j = lane; j < d; j += 32) repr[j] = 0.0f + fw * r[j];
        if (lane == 0) repr[d] = 0.0f + fw * bias[row];
        __syncwarp();
        return;
    }
    int start = f.indptr[row], stop = f.indptr[row + 1];
    for (int j = lane; j < d; j += 32) repr[j] = 0.0f;
    if (lane == 0) repr[d] = 0.0f;
    for (int i = start; i < stop; i++) {
        int ft = f.indices[i];
        float fw = (float)((double)f.data[i] * scale);
        const float* r = emb + (size_t)ft * d;
        for (int j = lane; j < d; j += 32) repr[j] = repr[j] + fw * r[j];
        if (lane == 0) repr[d] = repr[d] + fw * bias[ft];
    }
    __syncwarp();
}

// Two independent gathers (user + positive item).  With identity features on both sides their
// row loads are issued together so the two L2 round trips overlap; otherwise one after the other.
__device__ void gather2(const DevCsr& fa, const float* ea, const float* ba, int rowa, double sa, float* ra,
                        const DevCsr& fb, const float* eb, const float* bb, int rowb, double sb, float* rb,
                        int d, int lane) {
    if (fa.identity && fb.identity) {
        __syncwarp();
        const float fwa = (float)((double)1.0f * sa), fwb = (float)((double)1.0f * sb);
        const float* pa = ea + (size_t)rowa * d;
        const float* pb = eb + (size_t)rowb * d;
        for (int j = lane; j < d; j += 32) {
            const float xa = pa[j], xb = pb[j];
            ra[j] = 0.0f + fwa * xa;
            rb[j] = 0.0f + fwb * xb;
        }
        if (lane == 0) {
            const float xa = ba[rowa], xb = bb[rowb];
            ra[d] = 0.0f + fwa * xa;
            rb[d] = 0.0f + fwb * xb;
        }
        __syncwarp();
        return;
    }
    gather(fa, ea, ba, d, rowa, sa, ra, lane);
    gather(fb, eb, bb, d, rowb, sb, rb, lane);
}

// T:320-334: strictly left-to-right fp32 sum; every lane computes the same value.
// The adds form one dependent chain by definition; the loads and products do not, so they are
// issued eight at a time ahead of the chain (same operations, same order of additions).
__device__ float score(const float* u, const float* v, int d) {
    float r = u[d] + v[d];
    int i = 0;
    for (; i + 8 <= d; i += 8) {
        float p[8];
#pragma unroll
        for (int k = 0; k < 8; k++) p[k] = u[i + k] * v[i + k];
#pragma unroll
        for (int k = 0; k < 8; k++) r = r + p[k];
    }
    for (; i < d; i++) r = r + u[i] * v[i];
    return r;
}

__device__ float sigmoid_ref(float v) {  // T:262-267
    return (float)(1.0 / (1.0 + exp((double)(-v))));
}

// One parameter step; returns the local learning rate (T:359-389 / T:417-449).
__device__ double step(float* theta, float* G, float* M, double fw, double gradient, int adadelta,
                       double lr, double alpha, float rho, float eps) {
    double llr;
    if (adadelta) {
        double t = fw * gradient;
        float rg = rho * *G;
        *G = (float)((double)rg + (1.0 - (double)rho) * (t * t));
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

__device__ double bias_steps(const DevCsr& f, int row, DevTable& t, double gradient,
                             const DevModel& m, double alpha) {
    double s = 0.0;
    if (f.identity)  // single entry (row, 1.0f): 0.0 + llr, as the loop below would compute
        return s + step(&t.b[row], &t.bg[row], t.bm ? &t.bm[row] : nullptr, (double)1.0f, gradient,
                        m.adadelta, (double)m.lr, alpha, m.rho, m.eps);
    int start = f.indptr[row], stop = f.indptr[row + 1];
    for (int i = start; i < stop; i++) {
        int ft = f.indices[i];
        s += step(&t.b[ft], &t.bg[ft], t.bm ? &t.bm[ft] : nullptr, (double)f.data[i], gradient,
                  m.adadelta, (double)m.lr, alpha, m.rho, m.eps);
    }
    return s;
}

__device__ double row_steps(const DevCsr& f, int row, DevTable& t, int comp, double gradient,
                            const DevModel& m, double alpha) {
    double s = 0.0;
    if (f.identity) {
        size_t o = (size_t)row * m.d + comp;
        return s + step(&t.w[o], &t.g[o], t.m ? &t.m[o] : nullptr, (double)1.0f, gradient,
                        m.adadelta, (double)m.lr, alpha, m.rho, m.eps);
    }
    int start = f.indptr[row], stop = f.indptr[row + 1];
    for (int i = start; i < stop; i++) {
        size_t o = (size_t)f.indices[i] * m.d + comp;
        s += step(&t.w[o], &t.g[o], t.m ? &t.m[o] : nullptr, (double)f.data[i], gradient,
                  m.adadelta, (double)m.lr, alpha, m.rho, m.eps);
    }
    return s;
}

__device__ int nnz_of(const DevCsr& f, int row) { return f.identity ? 1 : f.indptr[row + 1] - f.indptr[row]; }

// T:537-649.  Returns avg learning rate contribution via scales update.
__device__ void warp_update(FitArgs& a, double loss, int user, int pos_id, int neg_id,
                            ReplayShared& sh, double& item_scale, double& user_scale, int lane) {
    DevModel& m = a.model;
    int d = m.d;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    if (a.itf.identity && a.usf.identity && pos_id != neg_id) {
        // three different addresses: lanes 0 / 1 / 2 take one bias each, round trips overlap
        if (lane == 0) b0 = bias_steps(a.itf, pos_id, m.item, -loss, m, a.item_alpha);
        if (lane == 1) b1 = bias_steps(a.itf, neg_id, m.item, loss, m, a.item_alpha);
        if (lane == 2) b2 = bias_steps(a.usf, user, m.user, loss, m, a.user_alpha);
        b1 = __shfl_sync(LFM_FULL, b1, 1);
        b2 = __shfl_sync(LFM_FULL, b2, 2);
    } else if (lane == 0) {
        b0 = bias_steps(a.itf, pos_id, m.item, -loss, m, a.item_alpha);
        b1 = bias_steps(a.itf, neg_id, m.item, loss, m, a.item_alpha);
        b2 = bias_steps(a.usf, user, m.user, loss, m, a.user_alpha);
    }
    if (a.itf.identity && a.usf.identity && pos_id != neg_id && d <= 256) {
        // Identity features: the nine values a component's three steps touch (w, G, M of the
        // positive, negative and user rows) are private to this lane and distinct, so they are
        // fetched up front for all of the lane's components -- one overlapped L2 round trip
        // instead of a dependent one per step -- then stepped in the reference's order and
        // stored.  The arithmetic is step()'s, untouched.
        constexpr int KMAX = 8;
        float w[KMAX][3], g[KMAX][3], mo[KMAX][3];
        const size_t op = (size_t)pos_id * d, on = (size_t)neg_id * d, ou = (size_t)user * d;
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            const int i = lane + 32 * k;
            if (i < d) {
                w[k][0] = m.item.w[op + i]; g[k][0] = m.item.g[op + i];
                w[k][1] = m.item.w[on + i]; g[k][1] = m.item.g[on + i];
                w[k][2] = m.user.w[ou + i]; g[k][2] = m.user.g[ou + i];
                if (m.adadelta) {
                    mo[k][0] = m.item.m[op + i]; mo[k][1] = m.item.m[on + i]; mo[k][2] = m.user.m[ou + i];
                } else {
                    mo[k][0] = mo[k][1] = mo[k][2] = 0.0f;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            const int i = lane + 32 * k;
            if (i < d) {
                const float uc = sh.u[i], pc = sh.pos[i], nc = sh.neg[i];
                const double lr = (double)m.lr;
                sh.lrsum[3 * i + 0] = 0.0 + step(&w[k][0], &g[k][0], &mo[k][0], (double)1.0f, (-loss) * (double)uc,
                                                 m.adadelta, lr, a.item_alpha, m.rho, m.eps);
                sh.lrsum[3 * i + 1] = 0.0 + step(&w[k][1], &g[k][1], &mo[k][1], (double)1.0f, loss * (double)uc,
                                                 m.adadelta, lr, a.item_alpha, m.rho, m.eps);
                sh.lrsum[3 * i + 2] = 0.0 + step(&w[k][2], &g[k][2], &mo[k][2], (double)1.0f,
                                                 loss * (double)(float)(nc - pc), m.adadelta, lr, a.user_alpha,
                                                 m.rho, m.eps);
                m.item.w[op + i] = w[k][0]; m.item.g[op + i] = g[k][0];
                m.item.w[on + i] = w[k][1]; m.item.g[on + i] = g[k][1];
                m.user.w[ou + i] = w[k][2]; m.user.g[ou + i] = g[k][2];
                if (m.adadelta) {
                    m.item.m[op + i] = mo[k][0]; m.item.m[on + i] = mo[k][1]; m.user.m[ou + i] = mo[k][2];
                }
            }
        }
    } else {
        for (int i = lane; i < d; i += 32) {
            float uc = sh.u[i], pc = sh.pos[i], nc = sh.neg[i];
            sh.lrsum[3 * i + 0] = row_steps(a.itf, pos_id, m.item, i, (-loss) * (double)uc, m, a.item_alpha);
            sh.lrsum[3 * i + 1] = row_steps(a.itf, neg_id, m.item, i, loss * (double)uc, m, a.item_alpha);
            sh.lrsum[3 * i + 2] = row_steps(a.usf, user, m.user, i, loss * (double)(float)(nc - pc), m,
                                            a.user_alpha);
        }
    }
    __syncwarp();
    if (a.item_alpha != 0.0 || a.user_alpha != 0.0) {
        b0 = __shfl_sync(LFM_FULL, b0, 0);
        b1 = __shfl_sync(LFM_FULL, b1, 0);
        b2 = __shfl_sync(LFM_FULL, b2, 0);
        double avg = 0.0;
        avg += b0; avg += b1; avg += b2;
        for (int i = 0; i < 3 * d; i++) avg += sh.lrsum[i];  // same order as T:602-638
        avg /= (double)((d + 1) * nnz_of(a.usf, user) + (d + 1) * nnz_of(a.itf, pos_id) +
                        (d + 1) * nnz_of(a.itf, neg_id));
        item_scale *= (1.0 + a.item_alpha * avg);
        user_scale *= (1.0 + a.user_alpha * avg);
    }
    __syncwarp();
}

// T:454-534.
__device__ void logistic_update(FitArgs& a, double loss, int user, int item, ReplayShared& sh,
                                double& item_scale, double& user_scale, int lane) {
    DevModel& m = a.model;
    int d = m.d;
    double b0 = 0.0, b1 = 0.0;
    if (lane == 0) {
        b0 = bias_steps(a.itf, item, m.item, loss, m, a.item_alpha);
        b1 = bias_steps(a.usf, user, m.user, loss, m, a.user_alpha);
    }
    for (int i = lane; i < d; i += 32) {
        float uc = sh.u[i], ic = sh.pos[i];
        sh.lrsum[2 * i + 0] = row_steps(a.itf, item, m.item, i, loss * (double)uc, m, a.item_alpha);
        sh.lrsum[2 * i + 1] = row_steps(a.usf, user, m.user, i, loss * (double)ic, m, a.user_alpha);
    }
    __syncwarp();
    if (a.item_alpha != 0.0 || a.user_alpha != 0.0) {
        b0 = __shfl_sync(LFM_FULL, b0, 0);
        b1 = __shfl_sync(LFM_FULL, b1, 0);
        double avg = 0.0;
        avg += b0; avg += b1;
        for (int i = 0; i < 2 * d; i++) avg += sh.lrsum[i];
        avg /= (double)((d + 1) * nnz_of(a.usf, user) + (d + 1) * nnz_of(a.itf, item));
        item_scale *= (1.0 + a.item_alpha * avg);
        user_scale *= (1.0 + a.user_alpha * avg);
    }
    __syncwarp();
}

// T:652-675 executed by the single replay warp (mid-epoch rescale, T:901-904).
// Same ownership as everywhere else: lane l touches components l, l+32, ...; lane 0 the biases.
__device__ void regularize_inline(DevModel& m, double& item_scale, double& user_scale, int lane) {
    int d = m.d;
    for (int r = 0; r < m.item.n; r++) {
        float* w = m.item.w + (size_t)r * d;
        for (int j = lane; j < d; j += 32) w[j] = (float)((double)w[j] / item_scale);
        if (lane == 0) m.item.b[r] = (float)((double)m.item.b[r] / item_scale);
    }
    for (int r = 0; r < m.user.n; r++) {
        float* w = m.user.w + (size_t)r * d;
        for (int j = lane; j < d; j += 32) w[j] = (float)((double)w[j] / user_scale);
        if (lane == 0) m.user.b[r] = (float)((double)m.user.b[r] / user_scale);
    }
    item_scale = 1.0;
    user_scale = 1.0;
    __syncwarp();
}

template <int LOSS>
__global__ void __launch_bounds__(32, 1) replay_kernel(FitArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    int lane = threadIdx.x;
    DevModel& m = a.model;
    int d = m.d;
    ReplayShared sh;
    sh.lrsum = (double*)smem;
    sh.u = (float*)(sh.lrsum + 3 * d);
    sh.pos = sh.u + (d + 1);
    sh.neg = sh.pos + (d + 1);
    sh.pval = sh.neg + (d + 1);
    sh.pidx = (int*)(sh.pval + (a.nkos > 0 ? a.nkos : 1));

    double item_scale = 1.0, user_scale = 1.0;  // T:254-255
    uint32_t seed = a.seed;
    unsigned long long c_pos = 0, c_neg = 0, c_upd = 0, c_rej = 0;

    for (int64_t i = 0; i < a.n; i++) {
        int row = a.shuffle[i];
        int user = a.user_ids[row];

        if (LOSS == LOSS_LOGISTIC) {
            int item = a.item_ids[row];
            float weight = a.sample_weight[row];
            gather2(a.usf, m.user.w, m.user.b, user, user_scale, sh.u,
                    a.itf, m.item.w, m.item.b, item, item_scale, sh.pos, d, lane);
            double prediction = (double)sigmoid_ref(score(sh.u, sh.pos, d));
            int y = (a.y[row] <= 0) ? 0 : 1;
            double loss = (double)weight * (prediction - (double)y);
            logistic_update(a, loss, user, item, sh, item_scale, user_scale, lane);
            c_pos++; c_upd++;
        } else if (LOSS == LOSS_WARP) {
            int pos_id = a.item_ids[row];
            if (!(a.y[row] > 0)) continue;
            float weight = a.sample_weight[row];
            c_pos++;
            gather2(a.usf, m.user.w, m.user.b, user, user_scale, sh.u,
                    a.itf, m.item.w, m.item.b, pos_id, item_scale, sh.pos, d, lane);
            double pp = (double)score(sh.u, sh.pos, d);
            int sampled = 0;
            while (sampled < m.max_sampled) {
                sampled++;
                int neg_id = lfm_rand_r(seed) % a.itf.rows;
                gather(a.itf, m.item.w, m.item.b, d, neg_id, item_scale, sh.neg, lane);
                double np = (double)score(sh.u, sh.neg, d);
                c_neg++;
                if (np > pp - 1) {
                    if (lfm_warp_member(a.pos.indices, a.pos.indptr[user], a.pos.indptr[user + 1], neg_id, lane)) {
                        c_rej++;
                        continue;
                    }
                    double loss = (double)weight * a.loss_table[sampled];  // T:881
                    if (loss > LFM_MAX_LOSS) loss = LFM_MAX_LOSS;
                    warp_update(a, loss, user, pos_id, neg_id, sh, item_scale, user_scale, lane);
                    c_upd++;
                    break;
                }
            }
        } else if (LOSS == LOSS_BPR) {
            if (!(a.y[row] > 0)) continue;
            float weight = a.sample_weight[row];
            int pos_id = a.item_ids[row];
            int neg_id = 0;
            for (int64_t j = 0; j < a.n; j++) {  // T:1123-1127
                neg_id = a.item_ids[lfm_rand_r(seed) % (int)a.n];
                c_neg++;
                if (!lfm_warp_member(a.pos.indices, a.pos.indptr[user], a.pos.indptr[user + 1], neg_id, lane)) break;
                c_rej++;
            }
            gather2(a.usf, m.user.w, m.user.b, user, user_scale, sh.u,
                    a.itf, m.item.w, m.item.b, pos_id, item_scale, sh.pos, d, lane);
            gather(a.itf, m.item.w, m.item.b, d, neg_id, item_scale, sh.neg, lane);
            double pp = (double)score(sh.u, sh.pos, d);
            double np = (double)score(sh.u, sh.neg, d);
            double loss = (double)weight * (1.0 - (double)sigmoid_ref((float)(pp - np)));
            warp_update(a, loss, user, pos_id, neg_id, sh, item_scale, user_scale, lane);
            c_pos++; c_upd++;
        } else {  // LOSS_KOS, T:957-1062
            gather(a.usf, m.user.w, m.user.b, d, user, user_scale, sh.u, lane);
            int ps = a.pos.indptr[user], pe = a.pos.indptr[user + 1];
            if (pe == ps) continue;
            c_pos++;
            int no_pos = (a.nkos < pe - ps) ? a.nkos : (pe - ps);
            for (int j = 0; j < no_pos; j++) {
                int sid = a.pos.indices[ps + (lfm_rand_r(seed) % (pe - ps))];  // T:84-90
                gather(a.itf, m.item.w, m.item.b, d, sid, item_scale, sh.pos, lane);
                float v = score(sh.u, sh.pos, d);
                if (lane == 0) { sh.pidx[j] = sid; sh.pval[j] = v; }
                __syncwarp();
            }
            if (lane == 0) {  // stable descending insertion sort == qsort(reverse_pair_compare)
                for (int x = 1; x < no_pos; x++) {
                    int ki = sh.pidx[x]; float kv = sh.pval[x];
                    int y = x - 1;
                    while (y >= 0 && (sh.pval[y] - kv) < 0) {
                        sh.pidx[y + 1] = sh.pidx[y]; sh.pval[y + 1] = sh.pval[y]; y--;
                    }
                    sh.pidx[y + 1] = ki; sh.pval[y + 1] = kv;
                }
            }
            __syncwarp();
            int sel = ((a.k < no_pos) ? a.k : no_pos) - 1;
            int pos_id = sh.pidx[sel];
            double pp = (double)sh.pval[sel];
            __syncwarp();
            gather(a.itf, m.item.w, m.item.b, d, pos_id, item_scale, sh.pos, lane);
            int sampled = 0;
            while (sampled < m.max_sampled) {
                sampled++;
                int neg_id = lfm_rand_r(seed) % a.itf.rows;
                gather(a.itf, m.item.w, m.item.b, d, neg_id, item_scale, sh.neg, lane);
                double np = (double)score(sh.u, sh.neg, d);
                c_neg++;
                if (np > pp - 1) {
                    if (lfm_warp_member(a.pos.indices, ps, pe, neg_id, lane)) { c_rej++; continue; }
                    double loss = a.loss_table[sampled];  // T:1039: no weight, no max(1, .)
                    if (loss > LFM_MAX_LOSS) loss = LFM_MAX_LOSS;
                    warp_update(a, loss, user, pos_id, neg_id, sh, item_scale, user_scale, lane);
                    c_upd++;
                    break;
                }
            }
        }
        if (item_scale > LFM_MAX_REG_SCALE || user_scale > LFM_MAX_REG_SCALE)
            regularize_inline(m, item_scale, user_scale, lane);
    }
    if (lane == 0) {
        a.scales->item_scale = item_scale;  // final regularize (T:910-912) runs as its own kernel
        a.scales->user_scale = user_scale;
        a.counters->positives = c_pos;
        a.counters->negatives = c_neg;
        a.counters->updates = c_upd;
        a.counters->rejected = c_rej;
    }
}

// T:652-675 as a full-width sweep; x / 1.0 == x so the launch is skipped by the
// host when both scales are exactly 1 (alpha == 0).
__global__ void regularize_kernel(DevModel m, DevScales* scales) {
    double is = scales->item_scale, us = scales->user_scale;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t ni = (size_t)m.item.n * m.d, nu = (size_t)m.user.n * m.d;
    for (size_t i = tid; i < ni; i += stride) m.item.w[i] = (float)((double)m.item.w[i] / is);
    for (size_t i = tid; i < (size_t)m.item.n; i += stride) m.item.b[i] = (float)((double)m.item.b[i] / is);
    for (size_t i = tid; i < nu; i += stride) m.user.w[i] = (float)((double)m.user.w[i] / us);
    for (size_t i = tid; i < (size_t)m.user.n; i += stride) m.user.b[i] = (float)((double)m.user.b[i] / us);
}

}  // namespace

#include "lfm_replay_fast.cuh"
#include "lfm_replay_dataflow.cuh"

static std::atomic<int> g_replay_fast{1};
static std::atomic<int> g_replay_dataflow{1};
extern "C" int lfm_set_replay_fast(int enabled) { return g_replay_fast.exchange(enabled ? 1 : 0); }
extern "C" int lfm_set_replay_dataflow(int enabled) { return g_replay_dataflow.exchange(enabled ? 1 : 0); }

// Timing of the last dataflow epoch: schedule kernel, execute kernel (ms), tasks (-1: it declined).
extern "C" int lfm_last_replay_dataflow(double* schedule_ms, double* execute_ms, int32_t* tasks) {
    if (schedule_ms) *schedule_ms = g_rdf_ms[0];
    if (execute_ms) *execute_ms = g_rdf_ms[1];
    if (tasks) *tasks = g_rdf_tasks;
    return LFM_OK;
}

size_t lfm_replay_dataflow_scratch_bytes(int loss, const FitArgs& a, int64_t bitmap_limit_bytes) {
    if (!g_replay_fast.load() || !g_replay_dataflow.load()) return 0;
    return rdf_scratch_bytes(loss, a, bitmap_limit_bytes);
}

cudaError_t lfm_launch_replay(int loss, const FitArgs& a, cudaStream_t st) {
    if (g_replay_fast.load()) {
        cudaError_t e = g_replay_dataflow.load() ? lfm_try_launch_replay_dataflow(loss, a, st) : cudaErrorNotSupported;
        if (e != cudaErrorNotSupported) return e;
        cudaGetLastError();
        e = lfm_try_launch_replay_fast(loss, a, st);
        if (e != cudaErrorNotSupported) return e;
        cudaGetLastError();
    }
    int d = a.model.d;
    int nk = a.nkos > 0 ? a.nkos : 1;
    size_t smem = sizeof(double) * 3 * d + sizeof(float) * 3 * (d + 1) + (sizeof(float) + sizeof(int)) * nk + 16;
    switch (loss) {
#define LFM_CASE(L)                                                                              \
    case L:                                                                                      \
        if (smem > 48 * 1024)                                                                    \
            cudaFuncSetAttribute(replay_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                 (int)smem);                                                     \
        replay_kernel<L><<<1, 32, smem, st>>>(a);                                                \
        break;
        LFM_CASE(LOSS_LOGISTIC)
        LFM_CASE(LOSS_WARP)
        LFM_CASE(LOSS_BPR)
        LFM_CASE(LOSS_KOS)
#undef LFM_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t lfm_launch_regularize(const DevModel& m, DevScales* scales, cudaStream_t st) {
    regularize_kernel<<<148 * 8, 256, 0, st>>>(m, scales);
    return cudaGetLastError();
}
