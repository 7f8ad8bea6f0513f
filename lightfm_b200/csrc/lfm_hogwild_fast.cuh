// lfm_hogwild_fast.cuh -- specialised hogwild kernels (included by lfm_hogwild.cu).
//
// Eligibility: identity user AND item features, adagrad, item_alpha == user_alpha == 0,
// no_components d in {16, 32, 64, 128}.  This is BASELINE configs C1, C2, C4, C5.
//
// fast_slot_kernel (WARP / BPR / logistic / k-OS): every interaction owns a SLOT of
// LPR = d / (4 * VPL) lanes (a lane holds VPL float4 chunks of each row) and NS = 32 / LPR
// interactions run per warp in lockstep:
//     d = 64: VPL 2 -> 8 lanes per interaction, 4 interactions per warp (default for WARP)
//             VPL 1 -> 16 lanes,                2 interactions per warp
// The user / positive rows and their Adagrad accumulator rows are staged one group ahead with
// cp.async.cg into double-buffered shared memory; negatives are drawn with Philox4x32-7 and
// scored one per slot per round; membership in the user's positives is one load from an exact
// bitmap (resident plans) or an LPR-ary / 4*LPR-ary search of the sorted CSR row; updates are
// red.global.add.v4.f32 reductions performed in L2 (fire and forget).
//
// fast_rank_kernel (first generation; k-OS with n > LPR, and WARP under lfm_set_tuning(0)):
// one warp per interaction, the user and positive rows loaded by every slot, NS negative
// candidates scored per round speculatively -- the first violating one in draw order wins,
// exactly as if they had been drawn one at a time.
//
// How the design moved from the second to the first is recorded with the ncu captures in
// profiles/README.md.
#pragma once

namespace {

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
    red_add_v4(w, -lr * rsqrt_ftz(g0.x) * gx, -lr * rsqrt_ftz(g0.y) * gy, -lr * rsqrt_ftz(g0.z) * gz,
               -lr * rsqrt_ftz(g0.w) * gw);
    red_add_v4(G, gx * gx, gy * gy, gz * gz, gw * gw);
}
__device__ __forceinline__ void adagrad_scalar(float* b, float* G, float lr, float g) {
    float g0 = __ldcg(G);
    red_add(b, -lr * rsqrt_ftz(g0) * g);
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

// ---- staging helpers (cp.async = LDGSTS, .cg: L2-coherent, bypasses L1) ---------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ Philox4 philox7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 7; r++) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o = {c0, c1, c2, c3};
    return o;
}

__device__ __forceinline__ void adagrad_row4_g(float* w, float* G, const float4& g0, float lr,
                                               float gx, float gy, float gz, float gw) {
    red_add_v4(w, -lr * rsqrt_ftz(g0.x) * gx, -lr * rsqrt_ftz(g0.y) * gy, -lr * rsqrt_ftz(g0.z) * gz,
               -lr * rsqrt_ftz(g0.w) * gw);
    red_add_v4(G, gx * gx, gy * gy, gz * gz, gw * gw);
}

template <int KSR>
struct TupleScalars {
    float ub, pb;    // user / positive-item bias
    int ps, pe;      // bounds of the user's row in the positives CSR
    int probe;       // lane's first-level probe of that row (key independent, see slot_member)
    int sid[KSR];    // k-OS: the positive items this lane sampled from the user's row
                     // (sample j of the min(n, nnz_u) lives on lane j % LPR, register j / LPR)
};

// ---- WARP / BPR / logistic: one SLOT per interaction --------------------------------------
// ncu (profiles/r1_ncu_v3_summary.txt): once the dependent-load chain of the warp-per-interaction
// kernel was pipelined it became
// issue-bound (66% issue-active, ~600 warp-instructions per interaction), because with one
// warp per interaction every scalar step (Philox, membership search, control flow, the
// user.positive dot) is executed by 32 lanes for ONE interaction.  Here each interaction gets
// LPR = d / (4*VPL) lanes, each lane holding VPL float4 chunks of a row, and NS = 32/LPR
// interactions run per warp in lockstep: one instruction stream serves NS interactions, no
// candidate row is ever loaded speculatively, and the sampling rounds of the NS interactions
// overlap in time.  A slot that has found its negative (or exhausted max_sampled) idles until
// its warp-mates finish; E[max of NS geometric draws] / NS < E[one draw].
//     d = 64 : VPL 1 -> 2 interactions / warp (v4) ; VPL 2 -> 4 interactions / warp (v5)
//
// Staging (cp.async.cg into double-buffered shared memory, issued one group ahead) and the
// prefetched first membership probe are per slot.
//
// Membership of `key` in the sorted CSR row idx[lo, hi) by the LPR lanes of a slot.  Level 1
// uses LPR probes that do not depend on the key and were prefetched with the row bounds
// (len <= LPR: the row itself; else pivots lo + (len*l >> log2 LPR)).  Every further level is
// 4*LPR-ary: each lane issues four independent probes, so a row of length L costs
// ceil(log_{4 LPR}(L / LPR)) dependent L2 round trips instead of log2 L (T:270-284).
template <int LPR>
__device__ __forceinline__ bool slot_member(const int32_t* __restrict__ idx, int lo, int hi, int probe,
                                            int key, bool need, int sub, unsigned slotmask) {
    constexpr int LOG = LPR == 32 ? 5 : LPR == 16 ? 4 : LPR == 8 ? 3 : 2;
    constexpr int W = 4 * LPR;  // fan-out of the deeper levels
    bool found = false, busy = need;
    if (!__any_sync(LFM_FULL, busy)) return false;
    int len = hi - lo;
    {   // level 1, from registers
        const bool inr = busy && (len > LPR || sub < len);
        const unsigned le = __ballot_sync(LFM_FULL, inr && probe <= key) & slotmask;
        const unsigned eq = __ballot_sync(LFM_FULL, inr && probe == key) & slotmask;
        if (busy) {
            const int c = __popc(le);
            if (eq != 0) { found = true; busy = false; }
            else if (len <= LPR || c == 0) { busy = false; }
            else {
                const int nlo = lo + (int)(((unsigned long long)(unsigned)len * (unsigned)(c - 1)) >> LOG);
                const int nhi = c == LPR ? hi : lo + (int)(((unsigned long long)(unsigned)len * (unsigned)c) >> LOG);
                lo = nlo + 1;
                hi = nhi;
                len = hi - lo;
                if (len <= 0) busy = false;
            }
        }
    }
    while (__any_sync(LFM_FULL, busy)) {
        // this lane's four probe positions: elements (len <= W) or pivots q_j = lo + (len*j / W)
        int v[4];
        bool in[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = sub * 4 + i;
            const int pos = len <= W ? lo + j : lo + (int)(((unsigned long long)(unsigned)len * (unsigned)j) / W);
            in[i] = busy && (len > W || j < len);
            v[i] = in[i] ? __ldg(idx + pos) : 0;
        }
        int cnt = 0;
        bool hit = false;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cnt += (in[i] && v[i] <= key) ? 1 : 0;
            hit |= in[i] && v[i] == key;
        }
        const bool anyhit = (__ballot_sync(LFM_FULL, hit) & slotmask) != 0;
        int c = cnt;  // pivots are sorted, so the count of pivots <= key is the sum over lanes
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) c += __shfl_xor_sync(LFM_FULL, c, o);
        if (busy) {
            if (anyhit) { found = true; busy = false; }
            else if (len <= W || c == 0) { busy = false; }
            else {
                const int nlo = lo + (int)(((unsigned long long)(unsigned)len * (unsigned)(c - 1)) / W);
                const int nhi = c == W ? hi : lo + (int)(((unsigned long long)(unsigned)len * (unsigned)c) / W);
                lo = nlo + 1;
                hi = nhi;
                len = hi - lo;
                if (len <= 0) busy = false;
            }
        }
    }
    return found;
}
template <int LPR>
__device__ __forceinline__ int slot_probe_index(int lo, int hi, int sub) {
    constexpr int LOG = LPR == 32 ? 5 : LPR == 16 ? 4 : LPR == 8 ? 3 : 2;
    const int len = hi - lo;
    if (len <= LPR) return sub < len ? lo + sub : -1;
    return lo + (int)(((unsigned long long)(unsigned)len * (unsigned)sub) >> LOG);
}

//
// PROBE (tests only, lfm_set_probe): the same kernel body run as ONE warp with ONE interaction in
// flight (slot 0; tuple i sits in group i), rows staged at the top of the iteration instead of one
// group ahead, and negatives drawn from the reference's sequential rand_r stream instead of
// Philox.  What is left to differ from the oracle is exactly this kernel's arithmetic (fp32
// temporaries, FMA, lr * rsqrt.approx(G), float log table) -- tests/test_gpu_probe.py states by
// how much.
//
// SPEC (WARP / k-OS, lfm_set_tuning(9)): two candidates per slot per round -- both rows are
// requested together and judged in draw order, so the outcome is exactly that of drawing them one
// at a time (the second one is simply not consumed when the first is taken), while the number of
// dependent L2 round trips per interaction is halved.
//
// ATOMG (lfm_set_atomic_accumulators, default on): the accumulator update is an atomic add that
// RETURNS the old value, and the step is scaled by that value -- lr / sqrt(G) with G containing
// every earlier update of the element, as in the sequential algorithm.  With a plain
// read-then-reduce, the ~10^2 interactions in flight that touch the same popular item all scale
// their step by the same stale G; while accumulators are still small that overshoots, and it
// costs top-of-ranking precision at C2 shape (tests/test_gpu_tierb.py: p@10 outside the
// reference's own band without it).  The accumulator rows of the user and the positive item are
// then no longer staged, which halves the staging traffic.
template <int LOSS, int D, int VPL, int MINB, bool BITMAP = false, int KSR = 1, bool PROBE = false, bool SPEC = false,
          bool ATOMG = false>
__global__ void __launch_bounds__(256, MINB) fast_slot_kernel(FitArgs a, const Tuple* __restrict__ tuples) {
    constexpr bool PAIRWISE = LOSS != LOSS_LOGISTIC;  // has a negative item and a positives CSR
    constexpr bool KOS = LOSS == LOSS_KOS;            // positive item chosen in-kernel (T:975-1011)
    static_assert(!KOS || VPL == 1, "k-OS keeps its sampled positives KSR per lane: VPL must be 1");
    static_assert(KOS || KSR == 1, "KSR only matters for k-OS");
    typedef TupleScalars<KSR> Scalars;
    constexpr int LPR = D / (4 * VPL);
    constexpr int NS = 32 / LPR;
    constexpr int BUFF = (ATOMG ? 2 : 4) * D;  // floats per slot per buffer: u, p (+ Gu, Gp rows unless ATOMG)
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane % LPR, slot = lane / LPR;
    const unsigned slotmask = (LPR == 32 ? 0xffffffffu : ((1u << LPR) - 1u)) << (slot * LPR);
    float* sbuf = smem + ((size_t)wib * NS + slot) * 2 * BUFF;  // this slot's two buffers
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n_tuples = (int)a.n;
    const DevModel& m = a.model;
    const float lr = m.lr;
    const int n_items = a.itf.rows;
    const int max_sampled = m.max_sampled;
    unsigned c_pos = 0, c_neg = 0, c_upd = 0, c_rej = 0;  // per slot, kept on its sub == 0 lane

    // lane's chunk v of a row starts at float offset (sub + LPR*v)*4: consecutive lanes read
    // consecutive 16 B, so every load instruction of a slot is one contiguous LPR*16 B piece
    auto stage = [&](const Tuple& tp, float* buf, Scalars& sc) {
        if (tp.user < 0) return;
#pragma unroll
        for (int v = 0; v < VPL; v++) {
            const int o = (sub + LPR * v) * 4;
            cp_async16(buf + 0 * D + o, m.user.w + (size_t)tp.user * D + o);
            if (!ATOMG) cp_async16(buf + 2 * D + o, m.user.g + (size_t)tp.user * D + o);
            if (!KOS) {
                cp_async16(buf + 1 * D + o, m.item.w + (size_t)tp.item * D + o);
                if (!ATOMG) cp_async16(buf + 3 * D + o, m.item.g + (size_t)tp.item * D + o);
            }
        }
        sc.ub = __ldcg(m.user.b + tp.user);
        if (!KOS) sc.pb = __ldcg(m.item.b + tp.item);
        if (PAIRWISE && (!BITMAP || KOS)) {
            sc.ps = __ldg(a.pos.indptr + tp.user);
            sc.pe = __ldg(a.pos.indptr + tp.user + 1);
            if (!BITMAP) {
                const int pi = slot_probe_index<LPR>(sc.ps, sc.pe, sub);
                sc.probe = pi >= 0 ? __ldg(a.pos.indices + pi) : -1;
            }
        }
        if (KOS) {
            // lane `sub` draws samples sub, sub + LPR, ... of the min(n, nnz_u) positives sampled with
            // replacement (T:976-980); the draws only need the row bounds, so they are prefetched too.
            // tp.item carries the tuple's index in the epoch (pack_kernel), the Philox counter.
            const int len = sc.pe - sc.ps;
            sc.pb = 0.0f;
#pragma unroll
            for (int rr = 0; rr < KSR; rr++) {
                const int j = rr * LPR + sub;
                const Philox4 r4 = philox7((uint32_t)tp.item, 0u, (uint32_t)(j >> 2), 1u, a.seed, 0x4c464d31u);
                const uint32_t r = (j & 3) == 0 ? r4.x : (j & 3) == 1 ? r4.y : (j & 3) == 2 ? r4.z : r4.w;
                sc.sid[rr] = (len > 0 && j < min(a.nkos, len)) ? __ldg(a.pos.indices + sc.ps + lfm_bounded(r, (uint32_t)len)) : 0;
            }
        }
    };
    // PROBE: tuple i is "group" i and only slot 0 carries it
    const int n_end = PROBE ? n_tuples * NS : n_tuples;
    auto fetch = [&](int base) -> Tuple {
        Tuple tp = {-1, 0, 0.0f, 0.0f};
        if (PROBE) {
            if (base >= 0 && base < n_end && slot == 0) tp = tuples[base / NS];
            return tp;
        }
        const int t = base + slot;
        if (base >= 0 && t < n_tuples) tp = tuples[t];
        return tp;
    };
    uint32_t rr_state = a.seed;  // PROBE: the reference's rand_r stream (T:64-81)

    int base = warp * NS;  // first tuple of this warp's group; groups are nwarps*NS apart
    Tuple cur = fetch(base < n_end ? base : -1);
    Scalars cs = {0.f, 0.f, 0, 0, -1, {0}};
    if (!PROBE) {
        stage(cur, sbuf, cs);
        cp_async_commit();
    }
    int flip = 0;

    for (; base < n_end; base += nwarps * NS, flip ^= 1) {
        const int nbase = base + nwarps * NS;
        Tuple nxt = fetch((nbase > 0 && nbase < n_end) ? nbase : -1);
        Scalars ns = {0.f, 0.f, 0, 0, -1, {0}};
        float* buf = sbuf + flip * BUFF;
        float* nbuf = sbuf + (flip ^ 1) * BUFF;
        if (PROBE) {  // stage now: the previous interaction's update is complete and visible
            stage(cur, buf, cs);
            cp_async_commit();
        }
        cp_async_wait_all();
        __syncwarp();
        const bool valid = cur.user >= 0 && (!KOS || cs.pe > cs.ps);  // k-OS skips empty rows (T:972-973)
        const int t = base + slot;
        float4 u[VPL];
        float4 pk = make_float4(0.f, 0.f, 0.f, 0.f);  // k-OS: the chosen positive's row (not staged)
        float pp = 0.0f;
#pragma unroll
        for (int v = 0; v < VPL; v++) {
            u[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                u[v] = *(const float4*)(buf + 0 * D + (sub + LPR * v) * 4);
                if (!KOS) pp += dot4(u[v], *(const float4*)(buf + 1 * D + (sub + LPR * v) * 4));
            }
        }
        int pos_id = cur.item;
        if constexpr (KOS) {
            // ---- T:975-1011: score the sampled positives (two rows in flight per round), keep the
            //      score of sample j on lane j, take the k-th best in stable descending order ----
            const int no_pos = valid ? min(a.nkos, cs.pe - cs.ps) : 0;
            const int slot0 = slot * LPR;
            float my_val[KSR];
#pragma unroll
            for (int rr = 0; rr < KSR; rr++) my_val[rr] = 0.0f;
            int maxn = no_pos;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(LFM_FULL, maxn, o));
#pragma unroll
            for (int rr = 0; rr < KSR; rr++) {
                for (int l = 0; l < LPR && rr * LPR + l < maxn; l += 2) {
                    const int j = rr * LPR + l;
                    const int sa = __shfl_sync(LFM_FULL, cs.sid[rr], slot0 + l);
                    const int sb = __shfl_sync(LFM_FULL, cs.sid[rr], slot0 + min(l + 1, LPR - 1));
                    const bool aa = j < no_pos, ab = l + 1 < LPR && j + 1 < no_pos;
                    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
                    float ba = 0.0f, bb = 0.0f;
                    if (aa) { ra = ldcg4(m.item.w + (size_t)sa * D + sub * 4); ba = __ldcg(m.item.b + sa); }
                    if (ab) { rb = ldcg4(m.item.w + (size_t)sb * D + sub * 4); bb = __ldcg(m.item.b + sb); }
                    const float va = slot_sum<LPR>(dot4(u[0], ra)) + cs.ub + ba;
                    const float vb = slot_sum<LPR>(dot4(u[0], rb)) + cs.ub + bb;
                    if (sub == l) my_val[rr] = va;
                    if (sub == l + 1) my_val[rr] = vb;
                }
            }
            // position of every sample in the stable descending order == qsort(reverse_pair_compare)
            int rank[KSR];
#pragma unroll
            for (int rr = 0; rr < KSR; rr++) rank[rr] = 0;
#pragma unroll
            for (int r2 = 0; r2 < KSR; r2++) {
                for (int l = 0; l < LPR && r2 * LPR + l < maxn; l++) {
                    const int j2 = r2 * LPR + l;
                    const float vj = __shfl_sync(LFM_FULL, my_val[r2], slot0 + l);
#pragma unroll
                    for (int rr = 0; rr < KSR; rr++) {
                        const int j = rr * LPR + sub;
                        rank[rr] += (j2 < no_pos && (vj > my_val[rr] || (vj == my_val[rr] && j2 < j))) ? 1 : 0;
                    }
                }
            }
            const int sel = min(a.k, no_pos) - 1;
            int hit_sid = cs.sid[0];
            float hit_val = my_val[0];
            bool has = false;
#pragma unroll
            for (int rr = 0; rr < KSR; rr++) {
                if (!has && valid && rr * LPR + sub < no_pos && rank[rr] == sel) {
                    has = true; hit_sid = cs.sid[rr]; hit_val = my_val[rr];
                }
            }
            const unsigned hit = __ballot_sync(LFM_FULL, has) & slotmask;
            const int src = hit ? __ffs(hit) - 1 : slot0;  // NaN scores: fall back to the first sample
            pos_id = __shfl_sync(LFM_FULL, hit_sid, src);
            pp = __shfl_sync(LFM_FULL, hit_val, src);
            if (valid) pk = ldcg4(m.item.w + (size_t)pos_id * D + sub * 4);  // T:1005-1011
        } else {
            pp = slot_sum<LPR>(pp) + cs.ub + cs.pb;
        }

        int sampled = 0, neg_id = -1;
        float loss = 0.0f;
        float4 q[VPL];
#pragma unroll
        for (int v = 0; v < VPL; v++) q[v] = make_float4(0.f, 0.f, 0.f, 0.f);

        if constexpr ((LOSS == LOSS_WARP || KOS) && SPEC && !PROBE) {
            // ---- rank sampling (T:855-899), two draws per round, judged in draw order ----
            Philox4 r4 = {0u, 0u, 0u, 0u};
            bool active = valid && max_sampled > 0;
            for (int round = 0; __any_sync(LFM_FULL, active); round++) {
                if ((round & 1) == 0)
                    r4 = philox7((uint32_t)t, 0u, (uint32_t)(round >> 1), 0u, a.seed, 0x4c464d31u);
                const uint32_t ra = (round & 1) ? r4.z : r4.x, rb = (round & 1) ? r4.w : r4.y;
                const int ca = lfm_bounded(ra, (uint32_t)n_items), cb = lfm_bounded(rb, (uint32_t)n_items);
                const bool act_b = active && sampled + 1 < max_sampled;
                float4 qb[VPL];
                float ba = 0.0f, bb = 0.0f;
#pragma unroll
                for (int v = 0; v < VPL; v++) qb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                // the membership words of both candidates leave with their rows (they depend on the draw
                // and the user only): a violating candidate then needs no further round trip
                uint32_t wa = 0u, wb = 0u;
                const uint32_t* brow = BITMAP ? a.pos_bitmap + (size_t)cur.user * a.bitmap_words : nullptr;
                if (active) {
#pragma unroll
                    for (int v = 0; v < VPL; v++) q[v] = ldcg4(m.item.w + (size_t)ca * D + (sub + LPR * v) * 4);
                    ba = __ldcg(m.item.b + ca);
                    if (BITMAP) wa = __ldg(brow + (ca >> 5));
                }
                if (act_b) {
#pragma unroll
                    for (int v = 0; v < VPL; v++) qb[v] = ldcg4(m.item.w + (size_t)cb * D + (sub + LPR * v) * 4);
                    bb = __ldcg(m.item.b + cb);
                    if (BITMAP) wb = __ldg(brow + (cb >> 5));
                }
                float pa = 0.0f, pb2 = 0.0f;
#pragma unroll
                for (int v = 0; v < VPL; v++) { pa += dot4(u[v], q[v]); pb2 += dot4(u[v], qb[v]); }
                const float npa = slot_sum<LPR>(pa) + cs.ub + ba;
                const float npb = slot_sum<LPR>(pb2) + cs.ub + bb;
                const bool va = active && npa > pp - 1.0f, vb = act_b && npb > pp - 1.0f;
                bool ma, mb;
                if (BITMAP) {
                    ma = va && ((wa >> (ca & 31)) & 1u);
                    mb = vb && ((wb >> (cb & 31)) & 1u);
                } else {
                    ma = slot_member<LPR>(a.pos.indices, cs.ps, cs.pe, cs.probe, ca, va, sub, slotmask);
                    mb = slot_member<LPR>(a.pos.indices, cs.ps, cs.pe, cs.probe, cb, vb, sub, slotmask);
                }
                if (active) {
                    sampled++;
                    if (va && !ma) {
                        neg_id = ca;
                        loss = fminf((KOS ? 1.0f : cur.weight) * __ldg(a.loss_table_f + sampled), (float)LFM_MAX_LOSS);
                    } else {
                        if (va && sub == 0) c_rej++;
                        if (act_b) {
                            sampled++;
                            if (vb && !mb) {
                                neg_id = cb;
                                loss = fminf((KOS ? 1.0f : cur.weight) * __ldg(a.loss_table_f + sampled), (float)LFM_MAX_LOSS);
#pragma unroll
                                for (int v = 0; v < VPL; v++) q[v] = qb[v];
                            } else if (vb && sub == 0) {
                                c_rej++;
                            }
                        }
                    }
                    active = neg_id < 0 && sampled < max_sampled;
                }
            }
        } else if constexpr (LOSS == LOSS_WARP || KOS) {
            // ---- rank sampling (T:855-899): every slot draws its own candidates in lockstep ----
            Philox4 r4 = {0u, 0u, 0u, 0u};
            bool active = valid && max_sampled > 0;
            for (int round = 0; __any_sync(LFM_FULL, active); round++) {
                if ((round & 3) == 0)
                    r4 = philox7((uint32_t)t, 0u, (uint32_t)(round >> 2), 0u, a.seed, 0x4c464d31u);
                const int w = round & 3;
                const uint32_t r = w == 0 ? r4.x : w == 1 ? r4.y : w == 2 ? r4.z : r4.w;
                const int cand = PROBE ? lfm_rand_r(rr_state) % n_items : lfm_bounded(r, (uint32_t)n_items);
                float qb = 0.0f;
                uint32_t wm = 0u;  // the candidate's membership word leaves with its row
                if (active) {
#pragma unroll
                    for (int v = 0; v < VPL; v++) q[v] = ldcg4(m.item.w + (size_t)cand * D + (sub + LPR * v) * 4);
                    qb = __ldcg(m.item.b + cand);
                    if (BITMAP) wm = __ldg(a.pos_bitmap + (size_t)cur.user * a.bitmap_words + (cand >> 5));
                }
                float part = 0.0f;
#pragma unroll
                for (int v = 0; v < VPL; v++) part += dot4(u[v], q[v]);
                const float np = slot_sum<LPR>(part) + cs.ub + qb;
                const bool viol = active && np > pp - 1.0f;
                bool member;
                if (BITMAP) {  // exact bitmap of the positives: one 4 B load instead of a search
                    member = viol && ((wm >> (cand & 31)) & 1u);
                } else {
                    member = slot_member<LPR>(a.pos.indices, cs.ps, cs.pe, cs.probe, cand, viol, sub, slotmask);
                }
                if (active) {
                    sampled++;
                    if (viol) {
                        if (member) {
                            if (sub == 0) c_rej++;
                        } else {
                            neg_id = cand;
                            // T:881 (weight * log term) / T:1039 (k-OS: no weight)
                            loss = fminf((KOS ? 1.0f : cur.weight) * __ldg(a.loss_table_f + sampled), (float)LFM_MAX_LOSS);
                        }
                    }
                    active = neg_id < 0 && sampled < max_sampled;
                }
            }
        } else if constexpr (LOSS == LOSS_BPR) {
            // ---- T:1123-1127: popularity-weighted draw from the interaction list until it is
            //      not one of the user's positives (almost always the first try) ----
            Philox4 r4 = {0u, 0u, 0u, 0u};
            bool active = valid;
            float qb = 0.0f;
            for (int round = 0; __any_sync(LFM_FULL, active); round++) {
                if ((round & 3) == 0)
                    r4 = philox7((uint32_t)t, 0u, (uint32_t)(round >> 2), 2u, a.seed, 0x4c464d31u);
                const int w = round & 3;
                const uint32_t r = w == 0 ? r4.x : w == 1 ? r4.y : w == 2 ? r4.z : r4.w;
                const int64_t j = PROBE ? (int64_t)(lfm_rand_r(rr_state) % (int)a.n_all)
                                        : (int64_t)(((unsigned long long)r * (unsigned long long)a.n_all) >> 32);
                const int cand = active ? __ldg(a.item_ids + j) : 0;
                bool member;
                if (BITMAP) {
                    // the candidate's row leaves together with its membership word: the draw is almost
                    // always kept, and then nothing is left to fetch
                    uint32_t wm = 0u;
                    if (active) {
                        wm = __ldg(a.pos_bitmap + (size_t)cur.user * a.bitmap_words + (cand >> 5));
#pragma unroll
                        for (int v = 0; v < VPL; v++) q[v] = ldcg4(m.item.w + (size_t)cand * D + (sub + LPR * v) * 4);
                        qb = __ldcg(m.item.b + cand);
                    }
                    member = active && ((wm >> (cand & 31)) & 1u);
                } else {
                    member = slot_member<LPR>(a.pos.indices, cs.ps, cs.pe, cs.probe, cand, active, sub, slotmask);
                }
                if (active) {
                    sampled++;
                    neg_id = cand;
                    if (member && sub == 0) c_rej++;
                    active = member && sampled < 256;
                }
            }
            if (!BITMAP && valid) {
#pragma unroll
                for (int v = 0; v < VPL; v++) q[v] = ldcg4(m.item.w + (size_t)neg_id * D + (sub + LPR * v) * 4);
                qb = __ldcg(m.item.b + neg_id);
            }
            float part = 0.0f;
#pragma unroll
            for (int v = 0; v < VPL; v++) part += dot4(u[v], q[v]);
            const float np = slot_sum<LPR>(part) + cs.ub + qb;
            loss = cur.weight * (1.0f - 1.0f / (1.0f + __expf(-(pp - np))));
        } else {  // logistic, T:747-759
            const float pred = 1.0f / (1.0f + __expf(-pp));
            loss = cur.weight * (pred - (cur.y > 0 ? 1.0f : 0.0f));
        }
        if (valid && sub == 0) { c_pos++; c_neg += sampled; }

        // ---- prefetch the next group while this one updates ----
        if (!PROBE) {
            stage(nxt, nbuf, ns);
            cp_async_commit();
        }

        // ---- update (T:454-534 / T:537-649): rows + biases of a slot, one instruction stream ----
        const bool upd = (LOSS == LOSS_WARP || KOS) ? neg_id >= 0 : valid;
        if constexpr (ATOMG) {
            if (upd) {
                const size_t op = (size_t)pos_id * D, ou = (size_t)cur.user * D;
                const size_t on = PAIRWISE ? (size_t)neg_id * D : 0;
                // biases first: sub 0 positive / item, 1 negative (pairwise) or user (logistic), 2 user
                float* bp = nullptr;
                float* bgp = nullptr;
                float bgrad = 0.0f;
                if constexpr (PAIRWISE) {
                    if (sub < 3) {
                        bp = sub == 0 ? m.item.b + pos_id : sub == 1 ? m.item.b + neg_id : m.user.b + cur.user;
                        bgp = sub == 0 ? m.item.bg + pos_id : sub == 1 ? m.item.bg + neg_id : m.user.bg + cur.user;
                        bgrad = sub == 0 ? -loss : loss;
                    }
                } else {
                    if (sub < 2) {
                        bp = sub == 0 ? m.item.b + cur.item : m.user.b + cur.user;
                        bgp = sub == 0 ? m.item.bg + cur.item : m.user.bg + cur.user;
                        bgrad = loss;
                    }
                }
                float bold = 1.0f;
                if (bp) bold = atomicAdd(bgp, bgrad * bgrad);
#pragma unroll
                for (int v = 0; v < VPL; v++) {
                    const int o = (sub + LPR * v) * 4;
                    const float4 p4 = KOS ? pk : *(const float4*)(buf + 1 * D + o);
                    const float lx = loss * u[v].x, ly = loss * u[v].y, lz = loss * u[v].z, lw = loss * u[v].w;
                    float4 gu;   // gradient of the user row
                    if constexpr (PAIRWISE)
                        gu = make_float4(loss * (q[v].x - p4.x), loss * (q[v].y - p4.y), loss * (q[v].z - p4.z),
                                         loss * (q[v].w - p4.w));
                    else
                        gu = make_float4(loss * p4.x, loss * p4.y, loss * p4.z, loss * p4.w);
                    // accumulate g^2 and get the accumulators as every earlier update left them
                    const float4 oP = atom_add_v4(m.item.g + op + o, lx * lx, ly * ly, lz * lz, lw * lw);
                    const float4 oU = atom_add_v4(m.user.g + ou + o, gu.x * gu.x, gu.y * gu.y, gu.z * gu.z, gu.w * gu.w);
                    float4 oN = make_float4(1.f, 1.f, 1.f, 1.f);
                    if constexpr (PAIRWISE) oN = atom_add_v4(m.item.g + on + o, lx * lx, ly * ly, lz * lz, lw * lw);
                    const float sg = PAIRWISE ? -1.0f : 1.0f;  // positive item: -loss * u (pairwise), +loss * u (logistic)
                    red_add_v4(m.item.w + op + o, -lr * rsqrt_ftz(oP.x) * sg * lx, -lr * rsqrt_ftz(oP.y) * sg * ly,
                               -lr * rsqrt_ftz(oP.z) * sg * lz, -lr * rsqrt_ftz(oP.w) * sg * lw);
                    red_add_v4(m.user.w + ou + o, -lr * rsqrt_ftz(oU.x) * gu.x, -lr * rsqrt_ftz(oU.y) * gu.y,
                               -lr * rsqrt_ftz(oU.z) * gu.z, -lr * rsqrt_ftz(oU.w) * gu.w);
                    if constexpr (PAIRWISE)
                        red_add_v4(m.item.w + on + o, -lr * rsqrt_ftz(oN.x) * lx, -lr * rsqrt_ftz(oN.y) * ly,
                                   -lr * rsqrt_ftz(oN.z) * lz, -lr * rsqrt_ftz(oN.w) * lw);
                }
                if (bp) red_add(bp, -lr * rsqrt_ftz(bold) * bgrad);
                if (sub == 0) c_upd++;
            }
        } else if (__any_sync(LFM_FULL, upd)) {
            float4 gn[VPL];
            float4 pk_g = make_float4(1.f, 1.f, 1.f, 1.f);  // k-OS: the chosen positive's accumulator row
            float bgv = 1.0f;  // bias accumulator: sub 0 item (positive), 1 negative / user, 2 user
            if (upd) {         // the only fetches left on the critical path, all issued together
                if constexpr (PAIRWISE) {
                    const size_t on = (size_t)neg_id * D;
#pragma unroll
                    for (int v = 0; v < VPL; v++) gn[v] = ldcg4(m.item.g + on + (sub + LPR * v) * 4);
                    if (sub < 3)
                        bgv = __ldcg(sub == 0 ? m.item.bg + pos_id : sub == 1 ? m.item.bg + neg_id
                                                                               : m.user.bg + cur.user);
                    if (KOS) pk_g = ldcg4(m.item.g + (size_t)pos_id * D + sub * 4);
                } else {
                    if (sub < 2) bgv = __ldcg(sub == 0 ? m.item.bg + cur.item : m.user.bg + cur.user);
                }
            }
            if (upd) {
                const size_t op = (size_t)pos_id * D, ou = (size_t)cur.user * D;
#pragma unroll
                for (int v = 0; v < VPL; v++) {
                    const int o = (sub + LPR * v) * 4;
                    const float4 p4 = KOS ? pk : *(const float4*)(buf + 1 * D + o);
                    const float4 gu = *(const float4*)(buf + 2 * D + o);
                    const float4 gp = KOS ? pk_g : *(const float4*)(buf + 3 * D + o);
                    const float lx = loss * u[v].x, ly = loss * u[v].y, lz = loss * u[v].z, lw = loss * u[v].w;
                    if constexpr (PAIRWISE) {
                        const size_t on = (size_t)neg_id * D;
                        adagrad_row4_g(m.item.w + op + o, m.item.g + op + o, gp, lr, -lx, -ly, -lz, -lw);
                        adagrad_row4_g(m.user.w + ou + o, m.user.g + ou + o, gu, lr, loss * (q[v].x - p4.x),
                                       loss * (q[v].y - p4.y), loss * (q[v].z - p4.z), loss * (q[v].w - p4.w));
                        adagrad_row4_g(m.item.w + on + o, m.item.g + on + o, gn[v], lr, lx, ly, lz, lw);
                    } else {
                        adagrad_row4_g(m.item.w + op + o, m.item.g + op + o, gp, lr, lx, ly, lz, lw);
                        adagrad_row4_g(m.user.w + ou + o, m.user.g + ou + o, gu, lr, loss * p4.x, loss * p4.y,
                                       loss * p4.z, loss * p4.w);
                    }
                }
                if constexpr (PAIRWISE) {
                    if (sub < 3) {  // biases: sub 0 positive (-loss), 1 negative (+loss), 2 user (+loss)
                        float* b = sub == 0 ? m.item.b + pos_id : sub == 1 ? m.item.b + neg_id : m.user.b + cur.user;
                        float* bg = sub == 0 ? m.item.bg + pos_id : sub == 1 ? m.item.bg + neg_id : m.user.bg + cur.user;
                        const float g = sub == 0 ? -loss : loss;
                        red_add(b, -lr * rsqrt_ftz(bgv) * g);
                        red_add(bg, g * g);
                    }
                } else {
                    if (sub < 2) {  // biases: sub 0 item, 1 user, both +loss
                        float* b = sub == 0 ? m.item.b + cur.item : m.user.b + cur.user;
                        float* bg = sub == 0 ? m.item.bg + cur.item : m.user.bg + cur.user;
                        red_add(b, -lr * rsqrt_ftz(bgv) * loss);
                        red_add(bg, loss * loss);
                    }
                }
                if (sub == 0) c_upd++;
            }
        }
        cur = nxt;
        cs = ns;
        if (PROBE) {  // make this interaction's reductions visible to every lane's next loads
            __threadfence();
            __syncwarp();
        }
    }
    cp_async_wait_all();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_pos += __shfl_xor_sync(LFM_FULL, c_pos, o);
        c_neg += __shfl_xor_sync(LFM_FULL, c_neg, o);
        c_upd += __shfl_xor_sync(LFM_FULL, c_upd, o);
        c_rej += __shfl_xor_sync(LFM_FULL, c_rej, o);
    }
    if (lane == 0) {
        atomicAdd(&a.counters->positives, (unsigned long long)c_pos);
        atomicAdd(&a.counters->negatives, (unsigned long long)c_neg);
        atomicAdd(&a.counters->updates, (unsigned long long)c_upd);
        atomicAdd(&a.counters->rejected, (unsigned long long)c_rej);
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

template <int LOSS, int D, int VPL, int MINB, bool BITMAP, int KSR = 1, bool SPEC = false, bool ATOMG = false>
cudaError_t launch_slot_impl(const FitArgs& b, const Tuple* tp, int64_t count, cudaStream_t st) {
    constexpr int LPR = D / (4 * VPL);
    constexpr int NS = 32 / LPR;
    constexpr int BT = 256, WPB = BT / 32;  // threads / warps per block
    const size_t smem = (size_t)WPB * NS * 2 * (ATOMG ? 2 : 4) * D * sizeof(float);
    auto kern = fast_slot_kernel<LOSS, D, VPL, MINB, BITMAP, KSR, false, SPEC, ATOMG>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem);
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t groups = (count + NS - 1) / NS;
    int64_t blocks = (int64_t)sms * per_sm;
    int64_t need = (groups + WPB - 1) / WPB, cap = ((lfm_inflight_cap(count) + NS - 1) / NS + WPB - 1) / WPB;
    if (blocks > need) blocks = need;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(int)blocks, BT, smem, st>>>(b, tp);
    return cudaGetLastError();
}

static std::atomic<int> g_atomg{1};

template <int LOSS, int D, int VPL, int MINB, int KSR = 1, bool SPEC = false>
cudaError_t launch_slot(const FitArgs& b, const Tuple* tp, int64_t count, cudaStream_t st) {
    if (g_atomg.load()) {
        if constexpr (LOSS != LOSS_LOGISTIC) {
            if (b.pos_bitmap) return launch_slot_impl<LOSS, D, VPL, MINB, true, KSR, SPEC, true>(b, tp, count, st);
        }
        return launch_slot_impl<LOSS, D, VPL, MINB, false, KSR, SPEC, true>(b, tp, count, st);
    }
    if constexpr (LOSS != LOSS_LOGISTIC) {
        if (b.pos_bitmap) return launch_slot_impl<LOSS, D, VPL, MINB, true, KSR, SPEC>(b, tp, count, st);
    }
    return launch_slot_impl<LOSS, D, VPL, MINB, false, KSR, SPEC>(b, tp, count, st);
}

// Test hook (lfm_set_probe): one warp, one interaction in flight, rand_r negatives.
static std::atomic<int> g_probe{0};
template <int LOSS, int D, int VPL, int MINB>
cudaError_t launch_probe(const FitArgs& b, const Tuple* tp, cudaStream_t st) {
    constexpr int NS = 32 / (D / (4 * VPL));
    const size_t smem = (size_t)NS * 2 * 2 * D * sizeof(float);
    auto kern = fast_slot_kernel<LOSS, D, VPL, MINB, false, 1, true, false, true>;  // PROBE, ATOMG
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 32, smem, st>>>(b, tp);
    return cudaGetLastError();
}

// Kernel variant (lfm_set_tuning), measured on C2 in profiles/README.md:
//   0      fast_rank_kernel for WARP as well (warp per interaction, speculative negatives)
//   4 / 5  fast_slot_kernel, one float4 per lane, 3 / 4 CTAs per SM
//   6/7/8  fast_slot_kernel, two float4 per lane (twice the interactions per warp; d >= 32),
//          2 / 3 / 4 CTAs per SM
//   9/10   as 7 / 6 with two speculative candidates per slot per round (SPEC)
//          [9 = default: 9.84 ms vs 10.65 ms for 7 on C2 with atomic accumulators]
// Also tried and dropped (within +-1.5 % of variant 7 once the bitmap was in): 128-thread blocks
// at 6-8 CTAs per SM (up to 28 warps / SM), and L2 evict-first cache hints on the tuple stream
// and the CSR probes.
static std::atomic<int> g_tuning{9};

template <int LOSS, int LPR>
cudaError_t launch_fast(const FitArgs& a, const Tuple* tuples, int64_t begin, int64_t count,
                        cudaStream_t st) {
    FitArgs b = a;
    b.n = count;
    const Tuple* tp = tuples + begin;
    constexpr int DD = 4 * LPR;
    if (g_probe.load()) {  // the templates the probe instantiates are the ones the benches time
        if (b.pos.indptr == nullptr && LOSS != LOSS_LOGISTIC) return cudaErrorInvalidValue;
        if constexpr (LOSS == LOSS_WARP && DD == 64) return launch_probe<LOSS, 64, 2, 3>(b, tp, st);
        if constexpr (LOSS == LOSS_BPR && DD == 64) return launch_probe<LOSS, 64, 1, 4>(b, tp, st);
        if constexpr (LOSS == LOSS_LOGISTIC && DD == 32) return launch_probe<LOSS, 32, 1, 4>(b, tp, st);
        return cudaErrorInvalidValue;
    }
    if constexpr (LOSS == LOSS_KOS) {
        // slot kernel keeps the n sampled positives KSR per lane of a slot: needs n <= KSR * LPR
        // (KSR = 3 is instantiated for d = 16 / 32, where the class default n = 10 exceeds LPR)
        if (g_tuning != 0 && a.nkos <= LPR) return launch_slot<LOSS, DD, 1, 4>(b, tp, count, st);
        if constexpr (LPR <= 8) {
            if (g_tuning != 0 && a.nkos <= 3 * LPR) return launch_slot<LOSS, DD, 1, 4, 3>(b, tp, count, st);
        }
        FastGrid g = fast_grid(fast_rank_kernel<LOSS, LPR>, count, lfm_inflight_cap(count));
        fast_rank_kernel<LOSS, LPR><<<g.blocks, g.threads, 0, st>>>(b, tp);
        return cudaGetLastError();
    } else {
        if constexpr (LOSS == LOSS_WARP) {
            if (g_tuning == 0 && b.pos.indptr != nullptr) {  // the first-generation kernel searches the CSR
                FastGrid g = fast_grid(fast_rank_kernel<LOSS, LPR>, count, lfm_inflight_cap(count));
                fast_rank_kernel<LOSS, LPR><<<g.blocks, g.threads, 0, st>>>(b, tp);
                return cudaGetLastError();
            }
        }
        if constexpr (LOSS != LOSS_WARP) {
            // no sampling loop to amortise: one float4 per lane at 4 CTAs per SM measured fastest
            // (C5 logistic: 20.9 ms vs 28.7 ms with two chunks per lane)
            return launch_slot<LOSS, DD, 1, 4>(b, tp, count, st);
        }
        if (g_tuning == 4) return launch_slot<LOSS, DD, 1, 3>(b, tp, count, st);
        if (g_tuning == 5 || g_tuning == 0) return launch_slot<LOSS, DD, 1, 4>(b, tp, count, st);
        if constexpr (DD >= 32) {
            if (g_tuning == 9) return launch_slot<LOSS, DD, 2, 3, 1, true>(b, tp, count, st);
            if (g_tuning == 10) return launch_slot<LOSS, DD, 2, 2, 1, true>(b, tp, count, st);
            if (g_tuning == 6) return launch_slot<LOSS, DD, 2, 2>(b, tp, count, st);
            if (g_tuning == 8) return launch_slot<LOSS, DD, 2, 4>(b, tp, count, st);
            return launch_slot<LOSS, DD, 2, 3>(b, tp, count, st);
        } else {
            return launch_slot<LOSS, DD, 1, 4>(b, tp, count, st);
        }
    }
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

extern "C" int lfm_set_tuning(int variant) {
    int old = g_tuning.load();
    if (variant == 0 || (variant >= 4 && variant <= 10)) g_tuning.store(variant);
    return old;
}

extern "C" int lfm_set_probe(int enabled) { return g_probe.exchange(enabled ? 1 : 0); }
extern "C" int lfm_set_atomic_accumulators(int enabled) { return g_atomg.exchange(enabled ? 1 : 0); }

// Set by lfm_set_fast_path (tests use it to exercise the generic kernels on fast-eligible inputs).
static std::atomic<int> g_fast_enabled{1};
extern "C" int lfm_set_fast_path(int enabled) {
    return g_fast_enabled.exchange(enabled ? 1 : 0);
}

int lfm_fast_path_eligible(int loss, const FitArgs& a, int64_t count) {
    const DevModel& m = a.model;
    if (!a.itf.identity || !a.usf.identity) return 0;
    if (m.adadelta || a.item_alpha != 0.0 || a.user_alpha != 0.0) return 0;
    if (loss == LOSS_KOS && a.nkos > 32) return 0;
    if (count > 0x7ff00000LL) return 0;  // the fast kernels index tuples with int32
    if (m.d != 16 && m.d != 32 && m.d != 64 && m.d != 128) return 0;
    // float4 path needs 16-byte aligned rows
    if ((((uintptr_t)m.item.w | (uintptr_t)m.item.g | (uintptr_t)m.user.w | (uintptr_t)m.user.g) & 15) != 0)
        return 0;
    return 1;
}

static cudaError_t lfm_try_launch_fast(int loss, const FitArgs& a, const Tuple* tuples,
                                       int64_t begin, int64_t count, cudaStream_t st, bool* done) {
    *done = false;
    const bool csr_free = loss != LOSS_LOGISTIC && a.pos.indptr == nullptr;  // bitmap-only plan
    if (!g_fast_enabled.load() && !csr_free) return cudaSuccess;
    if (!lfm_fast_path_eligible(loss, a, count)) return cudaSuccess;
    switch (loss) {
        case LOSS_LOGISTIC: return launch_fast_d<LOSS_LOGISTIC>(a, tuples, begin, count, st, done);
        case LOSS_WARP: return launch_fast_d<LOSS_WARP>(a, tuples, begin, count, st, done);
        case LOSS_BPR: return launch_fast_d<LOSS_BPR>(a, tuples, begin, count, st, done);
        case LOSS_KOS: return launch_fast_d<LOSS_KOS>(a, tuples, begin, count, st, done);
        default: return cudaSuccess;
    }
}
