// lfm_replay_fast.cuh -- replay mode, WARP loss, identity features, adagrad, alpha == 0: the
// reference's single-thread order, rand_r stream and per-element arithmetic (bit-equal to the
// oracle, like replay_kernel<LOSS_WARP>), with every L2 round trip that the sequential walk does
// not depend on taken off the critical path (included by lfm_replay.cu).
//
// Why it is legal: the walk is one warp; an interaction writes exactly three rows (user, positive,
// negative) and their biases.  Everything else it reads can be fetched early as long as it is
// re-fetched when one of those three rows turns out to be the same row:
//   * the interaction list is read three interactions ahead (shuffle -> tuple -> rows), and the
//     next interaction's user / positive rows, accumulators, biases and CSR row bounds are in
//     registers before it starts (re-read if the current update wrote them);
//   * every rand_r draw of fit_warp (T:860-861) becomes a negative candidate, in order, whatever
//     the control flow does, so the next RING draws are made early and their rows stream into a
//     shared-memory ring with cp.async (an entry written by an update in the meantime is marked
//     stale and read again when its turn comes).
// The arithmetic is replay_kernel's: score() and step() are the same functions.
#pragma once

namespace {

#define RP_RING 4  // negative candidates in flight

__device__ __forceinline__ void rp_cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void rp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int K>
struct RpRow {      // one table row as the lanes hold it: components lane, lane + 32, ...
    float w[K], g[K];
    float b, bg;    // bias and its accumulator (every lane holds a copy: broadcast loads)
    int id;
};

template <int K>
__device__ __forceinline__ void rp_load_row(RpRow<K>& r, const DevTable& t, int id, int d, int lane) {
    r.id = id;
    const size_t o = (size_t)id * d;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = lane + 32 * k;
        r.w[k] = j < d ? t.w[o + j] : 0.0f;
        r.g[k] = j < d ? t.g[o + j] : 1.0f;
    }
    r.b = t.b[id];
    r.bg = t.bg[id];
}

struct RpTuple {
    int user, item;
    float y, weight;
};

template <int K>
__global__ void __launch_bounds__(32, 1) replay_warp_fast_kernel(FitArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    DevModel& m = a.model;
    const int d = m.d;
    float* su = (float*)smem_raw;          // [d+1] user representation
    float* sp = su + (d + 1);              // [d+1] positive
    float* sn = sp + (d + 1);              // [d+1] negative
    float* ring = sn + (d + 1);            // [RP_RING][2 * d + 2]: w row, g row, b, bg
    const int rstride = 2 * d + 2;

    const int64_t n = a.n;
    const int n_items = a.itf.rows;
    const float fw = (float)((double)1.0f * 1.0);  // f32(double(1.0f) * scale), scale == 1 (alpha == 0)
    uint32_t seed = a.seed;
    unsigned long long c_pos = 0, c_neg = 0, c_upd = 0, c_rej = 0;

    // ---- candidate ring: ids drawn ahead, rows streamed in by cp.async ------------------------
    int ring_id[RP_RING];
    unsigned stale = 0;  // bit e: entry e was written by an update after it was requested
    auto ring_fetch = [&](int e, int id) {
        float* dst = ring + e * rstride;
        const size_t o = (size_t)id * d;
        for (int j = lane; j < d; j += 32) {
            rp_cp_async4(dst + j, m.item.w + o + j);
            rp_cp_async4(dst + d + j, m.item.g + o + j);
        }
        if (lane == 0) {
            rp_cp_async4(dst + 2 * d, m.item.b + id);
            rp_cp_async4(dst + 2 * d + 1, m.item.bg + id);
        }
        rp_commit();
    };
    int head = 0;  // ring slot of the next candidate to consume
#pragma unroll
    for (int e = 0; e < RP_RING; e++) {
        ring_id[e] = lfm_rand_r(seed) % n_items;
        ring_fetch(e, ring_id[e]);
    }

    // ---- interaction pipeline: shuffle index -> tuple -> rows ----------------------------------
    auto load_tuple = [&](int row) {
        RpTuple t;
        t.user = a.user_ids[row];
        t.item = a.item_ids[row];
        t.y = a.y[row];
        t.weight = a.sample_weight[row];
        return t;
    };
    RpTuple t0 = {0, 0, 0.f, 0.f}, t1 = t0;
    int row2 = 0;
    if (n > 0) t0 = load_tuple(a.shuffle[0]);
    if (n > 1) t1 = load_tuple(a.shuffle[1]);
    if (n > 2) row2 = a.shuffle[2];
    RpRow<K> U, P, NU, NP;
    int ps = 0, pe = 0, nps = 0, npe = 0;
    if (n > 0) {
        rp_load_row<K>(U, m.user, t0.user, d, lane);
        rp_load_row<K>(P, m.item, t0.item, d, lane);
        ps = a.pos.indptr[t0.user];
        pe = a.pos.indptr[t0.user + 1];
    }

    for (int64_t i = 0; i < n; i++) {
        // stage the next two interactions (their loads complete while this one computes)
        RpTuple t2 = {0, 0, 0.f, 0.f};
        if (i + 2 < n) t2 = load_tuple(row2);
        int row3 = 0;
        if (i + 3 < n) row3 = a.shuffle[i + 3];
        const bool have_next = i + 1 < n;
        if (have_next) {
            rp_load_row<K>(NU, m.user, t1.user, d, lane);
            rp_load_row<K>(NP, m.item, t1.item, d, lane);
            nps = a.pos.indptr[t1.user];
            npe = a.pos.indptr[t1.user + 1];
        }

        if (t0.y > 0) {
            c_pos++;
            const int user = t0.user, pos_id = t0.item;
            // representations (T:302-317 with the single identity feature): 0.0f + fw * E
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int j = lane + 32 * k;
                if (j < d) {
                    su[j] = 0.0f + fw * U.w[k];
                    sp[j] = 0.0f + fw * P.w[k];
                }
            }
            if (lane == 0) {
                su[d] = 0.0f + fw * U.b;
                sp[d] = 0.0f + fw * P.b;
            }
            __syncwarp();
            const double pp = (double)score(su, sp, d);
            int sampled = 0;
            while (sampled < m.max_sampled) {
                sampled++;
                // ---- consume the next candidate of the rand_r stream ----
                const int e = head;
                const int neg_id = ring_id[e];
                rp_wait<RP_RING - 1>();   // this thread's oldest outstanding group (entry e) has landed
                __syncwarp();
                float* slot = ring + e * rstride;
                RpRow<K> C;
                C.id = neg_id;
                if ((stale >> e) & 1u) {  // an update wrote this row after it was requested: read it again
                    rp_load_row<K>(C, m.item, neg_id, d, lane);
                } else {
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int j = lane + 32 * k;
                        C.w[k] = j < d ? slot[j] : 0.0f;
                        C.g[k] = j < d ? slot[d + j] : 1.0f;
                    }
                    C.b = slot[2 * d];
                    C.bg = slot[2 * d + 1];
                }
                __syncwarp();  // every lane has read the slot before it is refilled
                stale &= ~(1u << e);
                ring_id[e] = lfm_rand_r(seed) % n_items;  // draws beyond the epoch's last one are never used
                ring_fetch(e, ring_id[e]);
                head = (head + 1) % RP_RING;

#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int j = lane + 32 * k;
                    if (j < d) sn[j] = 0.0f + fw * C.w[k];
                }
                if (lane == 0) sn[d] = 0.0f + fw * C.b;
                __syncwarp();
                const double np = (double)score(su, sn, d);
                c_neg++;
                if (np > pp - 1) {
                    if (lfm_warp_member(a.pos.indices, ps, pe, neg_id, lane)) {
                        c_rej++;
                        continue;
                    }
                    double loss = (double)t0.weight * a.loss_table[sampled];  // T:881
                    if (loss > LFM_MAX_LOSS) loss = LFM_MAX_LOSS;
                    // ---- warp_update (T:537-649) on the prefetched copies; identical step()s ----
                    const double lr = (double)m.lr;
                    if (lane == 0) {
                        step(&P.b, &P.bg, nullptr, (double)1.0f, -loss, 0, lr, 0.0, m.rho, m.eps);
                        m.item.b[pos_id] = P.b;
                        m.item.bg[pos_id] = P.bg;
                    }
                    if (lane == 1) {
                        step(&C.b, &C.bg, nullptr, (double)1.0f, loss, 0, lr, 0.0, m.rho, m.eps);
                        m.item.b[neg_id] = C.b;
                        m.item.bg[neg_id] = C.bg;
                    }
                    if (lane == 2) {
                        step(&U.b, &U.bg, nullptr, (double)1.0f, loss, 0, lr, 0.0, m.rho, m.eps);
                        m.user.b[user] = U.b;
                        m.user.bg[user] = U.bg;
                    }
                    const size_t op = (size_t)pos_id * d, on = (size_t)neg_id * d, ou = (size_t)user * d;
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int j = lane + 32 * k;
                        if (j < d) {
                            const float uc = su[j], pc = sp[j], nc = sn[j];
                            step(&P.w[k], &P.g[k], nullptr, (double)1.0f, (-loss) * (double)uc, 0, lr, 0.0, m.rho, m.eps);
                            step(&C.w[k], &C.g[k], nullptr, (double)1.0f, loss * (double)uc, 0, lr, 0.0, m.rho, m.eps);
                            step(&U.w[k], &U.g[k], nullptr, (double)1.0f, loss * (double)(float)(nc - pc), 0, lr, 0.0,
                                 m.rho, m.eps);
                            m.item.w[op + j] = P.w[k]; m.item.g[op + j] = P.g[k];
                            m.item.w[on + j] = C.w[k]; m.item.g[on + j] = C.g[k];
                            m.user.w[ou + j] = U.w[k]; m.user.g[ou + j] = U.g[k];
                        }
                    }
                    __syncwarp();  // the stores above are ordered before the re-reads below, for all lanes
                    // ---- hazards: anything fetched early that this update has just rewritten ----
#pragma unroll
                    for (int q = 0; q < RP_RING; q++)
                        if (ring_id[q] == pos_id || ring_id[q] == neg_id) stale |= 1u << q;
                    if (have_next) {
                        if (NU.id == user) rp_load_row<K>(NU, m.user, user, d, lane);
                        if (NP.id == pos_id || NP.id == neg_id) rp_load_row<K>(NP, m.item, NP.id, d, lane);
                    }
                    c_upd++;
                    break;
                }
            }
        }
        // rotate the pipeline
        t0 = t1;
        t1 = t2;
        row2 = row3;
        if (have_next) {
            U = NU;
            P = NP;
            ps = nps;
            pe = npe;
        }
    }
    rp_wait<0>();
    if (lane == 0) {
        a.scales->item_scale = 1.0;  // alpha == 0: the scales never leave 1 (T:528-534)
        a.scales->user_scale = 1.0;
        a.counters->positives = c_pos;
        a.counters->negatives = c_neg;
        a.counters->updates = c_upd;
        a.counters->rejected = c_rej;
    }
}

}  // namespace

// Returns cudaErrorNotSupported when the inputs are outside this kernel's scope (the caller then
// launches the general replay kernel).
static cudaError_t lfm_try_launch_replay_fast(int loss, const FitArgs& a, cudaStream_t st) {
    const DevModel& m = a.model;
    if (loss != LOSS_WARP || !a.itf.identity || !a.usf.identity || m.adadelta || a.item_alpha != 0.0 ||
        a.user_alpha != 0.0 || m.d > 256 || a.n > 0x7fffffffLL || m.max_sampled < 1)
        return cudaErrorNotSupported;
    const int d = m.d;
    const size_t smem = sizeof(float) * (3 * (d + 1) + RP_RING * (2 * d + 2)) + 16;
#define RP_CASE(KK)                                                                                       \
    do {                                                                                                  \
        if (smem > 48 * 1024)                                                                             \
            cudaFuncSetAttribute(replay_warp_fast_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem);                                                              \
        replay_warp_fast_kernel<KK><<<1, 32, smem, st>>>(a);                                              \
    } while (0)
    if (d <= 32) RP_CASE(1);
    else if (d <= 64) RP_CASE(2);
    else if (d <= 128) RP_CASE(4);
    else RP_CASE(8);
#undef RP_CASE
    return cudaGetLastError();
}
