// lfm_hogwild_fast.cuh -- specialised hogwild kernels (included by lfm_hogwild.cu).
//
// Eligibility: identity user AND item features, adagrad, item_alpha == user_alpha == 0,
// no_components d in {16, 32, 64, 128}.  This is BASELINE configs C1, C2, C4, C5.
//
// Layout: a row of d floats is read by LPR = d/4 lanes as one float4 each, so a
// warp covers NS = 32/LPR rows ("slots") per load instruction:
//     d = 16 -> 8 slots, 32 -> 4, 64 -> 2, 128 -> 1.
// WARP / k-OS: one warp per interaction.  The user and positive rows are loaded
// by every slot (same addresses -> one L2 request), then NS negative candidates
// are scored per round, one per slot, speculatively: the first violating one in
// draw order wins, exactly as if they had been drawn one at a time (candidates
// after the winner are discarded and do not count as sampled).
// Logistic / BPR have no rank-sampling loop: one slot per interaction, NS
// interactions per warp.
// Updates: G row is re-read (ld.global.cg.v4), deltas go out as
// red.global.add.v4.f32 (fire-and-forget vector reductions performed in L2).
#pragma once

namespace {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg((const float4*)p); }

template <int LPR>
__device__ __forceinline__ float slot_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(LFM_FULL, v, o);
    return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Adagrad step on 4 consecutive parameters: w -= lr/sqrt(G) * g ; G += g^2   (alpha == 0)
__device__ __forceinline__ void adagrad_row4(float* w, float* G, float lr, float gx, float gy,
                                             float gz, float gw) {
    float4 g0 = ldcg4(G);
    red_add_v4(w, -lr * rsqrtf(g0.x) * gx, -lr * rsqrtf(g0.y) * gy, -lr * rsqrtf(g0.z) * gz,
               -lr * rsqrtf(g0.w) * gw);
    red_add_v4(G, gx * gx, gy * gy, gz * gz, gw * gw);
}
__device__ __forceinline__ void adagrad_scalar(float* b, float* G, float lr, float g) {
    float g0 = __ldcg(G);
    red_add(b, -lr * rsqrtf(g0) * g);
    red_add(G, g * g);
}

// ---- WARP and k-OS: one warp per interaction --------------------------------
template <int LOSS, int LPR>
__global__ void __launch_bounds__(256) fast_rank_kernel(FitArgs a, const Tuple* __restrict__ tuples) {
    constexpr int D = 4 * LPR;
    constexpr int NS = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, slot = lane / LPR;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const DevModel& m = a.model;
    const float lr = m.lr;
    const int n_items = a.itf.rows;
    const int max_sampled = m.max_sampled;
    unsigned long long c_pos = 0, c_neg = 0, c_upd = 0, c_rej = 0;

    for (int64_t t = warp; t < a.n; t += nwarps) {
        const Tuple tp = tuples[t];
        const int user = tp.user;
        if (user < 0) continue;
        const int ps = __ldg(a.pos.indptr + user), pe = __ldg(a.pos.indptr + user + 1);
        if (LOSS == LOSS_KOS && pe == ps) continue;
        c_pos++;
        const float4 u4 = ldcg4(m.user.w + (size_t)user * D + sub * 4);
        const float ub = __ldcg(m.user.b + user);

        uint32_t ctr = 0;
        int pos_id = tp.item;
        float pp;
        float4 p4;
        float pb;
        if (LOSS == LOSS_KOS) {
            // T:975-1011: draw min(n, nnz_u) positives with replacement, take the k-th best
            const int no_pos = min(a.nkos, pe - ps);
            int my_idx = 0;
            float my_val = 0.0f;
            for (int j0 = 0; j0 < no_pos; j0 += NS) {
                Philox4 r4 = lfm_philox((uint32_t)t, (uint32_t)(t >> 32), ctr + (slot >> 2), 1u, a.seed, 0x4c464d31u);
                ctr += (NS + 3) / 4;
                uint32_t r = (slot & 3) == 0 ? r4.x : (slot & 3) == 1 ? r4.y : (slot & 3) == 2 ? r4.z : r4.w;
                const bool act = (j0 + slot) < no_pos;
                int sid = act ? __ldg(a.pos.indices + ps + lfm_bounded(r, (uint32_t)(pe - ps))) : 0;
                float4 s4 = ldcg4(m.item.w + (size_t)sid * D + sub * 4);
                float sb = __ldcg(m.item.b + sid);
                float sc = slot_sum<LPR>(dot4(u4, s4)) + ub + sb;
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    int idx_s = __shfl_sync(LFM_FULL, sid, s * LPR);
                    float val_s = __shfl_sync(LFM_FULL, sc, s * LPR);
                    if (lane == j0 + s && j0 + s < no_pos) { my_idx = idx_s; my_val = val_s; }
                }
            }
            int rank = 0;
            for (int j = 0; j < no_pos; j++) {
                float vj = __shfl_sync(LFM_FULL, my_val, j);
                rank += (vj > my_val || (vj == my_val && j < lane)) ? 1 : 0;
            }
            const int sel = min(a.k, no_pos) - 1;
            unsigned hit = __ballot_sync(LFM_FULL, lane < no_pos && rank == sel);
            int src = __ffs(hit) - 1;
            if (src < 0) src = 0;
            pos_id = __shfl_sync(LFM_FULL, my_idx, src);
            pp = __shfl_sync(LFM_FULL, my_val, src);
            p4 = ldcg4(m.item.w + (size_t)pos_id * D + sub * 4);
            pb = __ldcg(m.item.b + pos_id);
        } else {
            p4 = ldcg4(m.item.w + (size_t)pos_id * D + sub * 4);
            pb = __ldcg(m.item.b + pos_id);
            pp = slot_sum<LPR>(dot4(u4, p4)) + ub + pb;
        }

        // ---- rank sampling (T:855-899): NS speculative candidates per round ----
        int sampled = 0;
        int neg_id = -1;
        int neg_lane = 0;
        float loss = 0.0f;
        float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
        while (sampled < max_sampled && neg_id < 0) {
            const int nb = min(NS, max_sampled - sampled);
            Philox4 r4 = lfm_philox((uint32_t)t, (uint32_t)(t >> 32), ctr + (slot >> 2), 0u, a.seed, 0x4c464d31u);
            ctr += (NS + 3) / 4;
            uint32_t r = (slot & 3) == 0 ? r4.x : (slot & 3) == 1 ? r4.y : (slot & 3) == 2 ? r4.z : r4.w;
            const int cand = lfm_bounded(r, (uint32_t)n_items);
            const bool act = slot < nb;
            if (act) q4 = ldcg4(m.item.w + (size_t)cand * D + sub * 4);
            const float qb = act ? __ldcg(m.item.b + cand) : 0.0f;
            const float np = slot_sum<LPR>(dot4(u4, q4)) + ub + qb;
            unsigned vm = __ballot_sync(LFM_FULL, act && sub == 0 && np > pp - 1.0f);
            int consumed = nb;
            while (vm) {
                const int first = __ffs(vm) - 1;  // lane = slot_k * LPR
                const int ck = __shfl_sync(LFM_FULL, cand, first);
                if (lfm_warp_member(a.pos.indices, ps, pe, ck, lane)) {
                    c_rej++;
                    vm &= vm - 1;
                    continue;
                }
                const int k = first / LPR;
                consumed = k + 1;
                neg_id = ck;
                neg_lane = first;
                float l = (float)a.loss_table[sampled + k + 1];
                loss = (LOSS == LOSS_KOS) ? l : tp.weight * l;
                loss = fminf(loss, (float)LFM_MAX_LOSS);
                break;
            }
            sampled += consumed;
            c_neg += consumed;
        }
        if (neg_id < 0) continue;
        c_upd++;

        // ---- update (T:537-649): three rows + three biases ----------------------
        float4 n4;
        n4.x = __shfl_sync(LFM_FULL, q4.x, neg_lane + sub);
        n4.y = __shfl_sync(LFM_FULL, q4.y, neg_lane + sub);
        n4.z = __shfl_sync(LFM_FULL, q4.z, neg_lane + sub);
        n4.w = __shfl_sync(LFM_FULL, q4.w, neg_lane + sub);
#pragma unroll
        for (int task = slot; task < 3; task += NS) {
            if (task == 0) {  // positive item row: gradient -loss * u
                size_t o = (size_t)pos_id * D + sub * 4;
                adagrad_row4(m.item.w + o, m.item.g + o, lr, -loss * u4.x, -loss * u4.y, -loss * u4.z,
                             -loss * u4.w);
            } else if (task == 1) {  // negative item row: +loss * u
                size_t o = (size_t)neg_id * D + sub * 4;
                adagrad_row4(m.item.w + o, m.item.g + o, lr, loss * u4.x, loss * u4.y, loss * u4.z,
                             loss * u4.w);
            } else {  // user row: loss * (neg - pos)
                size_t o = (size_t)user * D + sub * 4;
                adagrad_row4(m.user.w + o, m.user.g + o, lr, loss * (n4.x - p4.x), loss * (n4.y - p4.y),
                             loss * (n4.z - p4.z), loss * (n4.w - p4.w));
            }
        }
        if (lane == 0) adagrad_scalar(m.item.b + pos_id, m.item.bg + pos_id, lr, -loss);
        if (lane == 1) adagrad_scalar(m.item.b + neg_id, m.item.bg + neg_id, lr, loss);
        if (lane == 2) adagrad_scalar(m.user.b + user, m.user.bg + user, lr, loss);
    }
    if (lane == 0) {
        atomicAdd(&a.counters->positives, c_pos);
        atomicAdd(&a.counters->negatives, c_neg);
        atomicAdd(&a.counters->updates, c_upd);
        atomicAdd(&a.counters->rejected, c_rej);
    }
}

// ---- logistic and BPR: one slot per interaction ------------------------------
template <int LOSS, int LPR>
__global__ void __launch_bounds__(256) fast_pair_kernel(FitArgs a, const Tuple* __restrict__ tuples) {
    constexpr int D = 4 * LPR;
    constexpr int NS = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, slot = lane / LPR;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const DevModel& m = a.model;
    const float lr = m.lr;
    unsigned long long c_pos = 0, c_neg = 0, c_upd = 0, c_rej = 0;

    for (int64_t base = warp * NS; base < a.n; base += nwarps * NS) {
        const int64_t t = base + slot;
        Tuple tp = {-1, 0, 0.0f, 0.0f};
        if (t < a.n) tp = tuples[t];
        const bool act = tp.user >= 0;
        const int user = act ? tp.user : 0;
        const int item = act ? tp.item : 0;
        const float4 u4 = ldcg4(m.user.w + (size_t)user * D + sub * 4);
        const float ub = __ldcg(m.user.b + user);
        const float4 p4 = ldcg4(m.item.w + (size_t)item * D + sub * 4);
        const float pb = __ldcg(m.item.b + item);
        const float pp = slot_sum<LPR>(dot4(u4, p4)) + ub + pb;
        if (LOSS == LOSS_LOGISTIC) {
            const float pred = 1.0f / (1.0f + __expf(-pp));
            const float loss = tp.weight * (pred - (tp.y > 0 ? 1.0f : 0.0f));
            if (act) {
                size_t oi = (size_t)item * D + sub * 4, ou = (size_t)user * D + sub * 4;
                adagrad_row4(m.item.w + oi, m.item.g + oi, lr, loss * u4.x, loss * u4.y, loss * u4.z,
                             loss * u4.w);
                adagrad_row4(m.user.w + ou, m.user.g + ou, lr, loss * p4.x, loss * p4.y, loss * p4.z,
                             loss * p4.w);
                if (sub == 0) adagrad_scalar(m.item.b + item, m.item.bg + item, lr, loss);
                if (sub == 1) adagrad_scalar(m.user.b + user, m.user.bg + user, lr, loss);
                if (sub == 0) { c_pos++; c_upd++; }
            }
        } else {  // BPR, T:1113-1169
            int neg_id = 0;
            if (act) {
                const int ps = __ldg(a.pos.indptr + user), pe = __ldg(a.pos.indptr + user + 1);
                uint32_t ctr = 0;
                int rpos = 4;
                Philox4 r4 = {0u, 0u, 0u, 0u};
                for (int tries = 0; tries < 256; tries++) {
                    if (rpos == 4) {
                        r4 = lfm_philox((uint32_t)t, (uint32_t)(t >> 32), ctr++, 2u, a.seed, 0x4c464d31u);
                        rpos = 0;
                    }
                    uint32_t r = rpos == 0 ? r4.x : rpos == 1 ? r4.y : rpos == 2 ? r4.z : r4.w;
                    rpos++;
                    int64_t j = (int64_t)(((unsigned long long)r * (unsigned long long)a.n) >> 32);
                    neg_id = __ldg(a.item_ids + j);
                    if (sub == 0) c_neg++;
                    if (!lfm_bsearch(a.pos.indices, ps, pe, neg_id)) break;
                    if (sub == 0) c_rej++;
                }
            }
            const float4 n4 = ldcg4(m.item.w + (size_t)neg_id * D + sub * 4);
            const float nb = __ldcg(m.item.b + neg_id);
            const float np = slot_sum<LPR>(dot4(u4, n4)) + ub + nb;
            const float loss = tp.weight * (1.0f - 1.0f / (1.0f + __expf(-(pp - np))));
            if (act) {
                size_t op = (size_t)item * D + sub * 4, on = (size_t)neg_id * D + sub * 4,
                       ou = (size_t)user * D + sub * 4;
                adagrad_row4(m.item.w + op, m.item.g + op, lr, -loss * u4.x, -loss * u4.y, -loss * u4.z,
                             -loss * u4.w);
                adagrad_row4(m.item.w + on, m.item.g + on, lr, loss * u4.x, loss * u4.y, loss * u4.z,
                             loss * u4.w);
                adagrad_row4(m.user.w + ou, m.user.g + ou, lr, loss * (n4.x - p4.x), loss * (n4.y - p4.y),
                             loss * (n4.z - p4.z), loss * (n4.w - p4.w));
                if (sub == 0) adagrad_scalar(m.item.b + item, m.item.bg + item, lr, -loss);
                if (sub == 1) adagrad_scalar(m.item.b + neg_id, m.item.bg + neg_id, lr, loss);
                if (sub == 2) adagrad_scalar(m.user.b + user, m.user.bg + user, lr, loss);
                if (sub == 0) { c_pos++; c_upd++; }
            }
        }
    }
    // per-slot counters live on the sub == 0 lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_pos += __shfl_xor_sync(LFM_FULL, c_pos, o);
        c_neg += __shfl_xor_sync(LFM_FULL, c_neg, o);
        c_upd += __shfl_xor_sync(LFM_FULL, c_upd, o);
        c_rej += __shfl_xor_sync(LFM_FULL, c_rej, o);
    }
    if (lane == 0) {
        atomicAdd(&a.counters->positives, c_pos);
        atomicAdd(&a.counters->negatives, c_neg);
        atomicAdd(&a.counters->updates, c_upd);
        atomicAdd(&a.counters->rejected, c_rej);
    }
}

struct FastGrid {
    int blocks, threads;
};
template <typename K>
FastGrid fast_grid(K kernel, int64_t warps_wanted, int64_t warps_cap) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0);
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (int64_t)sms * per_sm;  // one full wave of resident CTAs (persistent warps)
    int64_t need = (warps_wanted + 7) / 8;
    if (blocks > need) blocks = need;
    int64_t cap = (warps_cap + 7) / 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    FastGrid g = {(int)blocks, 256};
    return g;
}

template <int LOSS, int LPR>
cudaError_t launch_fast(const FitArgs& a, const Tuple* tuples, int64_t begin, int64_t count,
                        cudaStream_t st) {
    FitArgs b = a;
    b.n = count;
    const Tuple* tp = tuples + begin;
    if constexpr (LOSS == LOSS_WARP || LOSS == LOSS_KOS) {
        FastGrid g = fast_grid(fast_rank_kernel<LOSS, LPR>, count, lfm_inflight_cap(count));
        fast_rank_kernel<LOSS, LPR><<<g.blocks, g.threads, 0, st>>>(b, tp);
    } else {
        constexpr int NS = 32 / LPR;
        FastGrid g = fast_grid(fast_pair_kernel<LOSS, LPR>, (count + NS - 1) / NS,
                               (lfm_inflight_cap(count) + NS - 1) / NS);
        fast_pair_kernel<LOSS, LPR><<<g.blocks, g.threads, 0, st>>>(b, tp);
    }
    return cudaGetLastError();
}

template <int LOSS>
cudaError_t launch_fast_d(const FitArgs& a, const Tuple* tuples, int64_t begin, int64_t count,
                          cudaStream_t st, bool* done) {
    *done = true;
    switch (a.model.d) {
        case 16: return launch_fast<LOSS, 4>(a, tuples, begin, count, st);
        case 32: return launch_fast<LOSS, 8>(a, tuples, begin, count, st);
        case 64: return launch_fast<LOSS, 16>(a, tuples, begin, count, st);
        case 128: return launch_fast<LOSS, 32>(a, tuples, begin, count, st);
        default: *done = false; return cudaSuccess;
    }
}

}  // namespace

// Set by lfm_set_fast_path (tests use it to exercise the generic kernels on fast-eligible inputs).
static int g_fast_enabled = 1;
extern "C" int lfm_set_fast_path(int enabled) {
    int old = g_fast_enabled;
    g_fast_enabled = enabled ? 1 : 0;
    return old;
}

static cudaError_t lfm_try_launch_fast(int loss, const FitArgs& a, const Tuple* tuples,
                                       int64_t begin, int64_t count, cudaStream_t st, bool* done) {
    *done = false;
    const DevModel& m = a.model;
    if (!g_fast_enabled) return cudaSuccess;
    if (!a.itf.identity || !a.usf.identity) return cudaSuccess;
    if (m.adadelta || a.item_alpha != 0.0 || a.user_alpha != 0.0) return cudaSuccess;
    if (loss == LOSS_KOS && a.nkos > 32) return cudaSuccess;
    // float4 path needs 16-byte aligned rows
    if ((((uintptr_t)m.item.w | (uintptr_t)m.item.g | (uintptr_t)m.user.w | (uintptr_t)m.user.g) & 15) != 0)
        return cudaSuccess;
    switch (loss) {
        case LOSS_LOGISTIC: return launch_fast_d<LOSS_LOGISTIC>(a, tuples, begin, count, st, done);
        case LOSS_WARP: return launch_fast_d<LOSS_WARP>(a, tuples, begin, count, st, done);
        case LOSS_BPR: return launch_fast_d<LOSS_BPR>(a, tuples, begin, count, st, done);
        case LOSS_KOS: return launch_fast_d<LOSS_KOS>(a, tuples, begin, count, st, done);
        default: return cudaSuccess;
    }
}
