// lfm_replay_dataflow.cuh -- replay mode (num_threads == 1) for BPR and logistic with identity
// features and alpha == 0: the reference's single-thread RESULT, bit for bit, computed by many
// warps at once (included by lfm_replay.cu; same score() / step() / sigmoid_ref() as replay_kernel).
//
// Why it is legal.  For these two losses nothing an interaction does depends on the weights except
// the arithmetic on the rows it touches:
//   * fit_logistic (T:694-781) touches the user row and the item row of the interaction;
//   * fit_bpr (T:1074-1182) additionally touches one negative item row, and WHICH row that is
//     depends only on the rand_r stream and on in_positives (T:1123-1127) -- data, not weights.
// So the whole epoch is a dependency graph known before any arithmetic runs: interaction t must
// see its rows exactly as the interactions before it (in visiting order) left them, and two
// interactions that share no row commute exactly (they read and write disjoint addresses).
//
//   1. rdf_schedule_kernel (one warp) walks the shuffled list in order.  It consumes the rand_r
//      stream exactly as the reference does (32 draws speculated per round through an LCG
//      jump-ahead; a rejected draw re-aligns the round), and gives every interaction a task
//      {user, item, negative, versions}: `version` of a row = how many earlier interactions touch it.
//   2. rdf_execute_kernel (all SMs, one warp per task, tasks handed out in visiting order): wait
//      until the task's rows have reached their versions (ld.acquire spin), run the reference's
//      arithmetic on them, publish version + 1 (st.release).  The earliest unfinished task never
//      waits on anything, so the walk always advances (cooperative launch: every warp is resident).
//
// The result equals replay_kernel<BPR / LOGISTIC> bit for bit (tests/test_gpu_replay_dataflow.py),
// which in turn is the oracle's / the reference's single-thread result.
#pragma once

namespace {

struct __align__(16) RdfTask {  // 32 B
    int32_t user, item, neg;  // neg: BPR only
    int32_t eu, ei, en;       // versions the user / item / negative rows must have reached
    float weight, y;
};

struct RdfScratch {
    int32_t* header;   // [0] number of tasks, or -1: not representable (caller runs replay_kernel); [1] stall flag
    int32_t* cnt_user; // schedule: touches so far            [n_users]
    int32_t* cnt_item; //                                     [n_items]
    int32_t* ver_user; // execute: published row versions     [n_users]
    int32_t* ver_item; //                                     [n_items]
    RdfTask* tasks;    // [n]
    const uint32_t* bitmap;  // optional exact membership bitmap of the positives CSR (else null)
    int32_t bitmap_words;
};

#define RDF_WARPS 8  // warps per CTA of the execute kernel
#define RDF_RING 128  // candidate ring entries (schedule kernel)
#define RDF_AHEAD 96  // draws requested ahead of the one being judged (3 cp.async groups)

__device__ __forceinline__ int rdf_ld_acquire(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rdf_st_release(int32_t* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int rdf_ld_volatile(const int32_t* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- 1. schedule ---------------------------------------------------------------------------
template <int LOSS>
__global__ void __launch_bounds__(32, 1) rdf_schedule_kernel(FitArgs a, RdfScratch s) {
    const int lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t n = a.n;
    // LCG jump-ahead (rand_r's state update, T:76): lane l holds (A^(l+1), C_(l+1)) with
    // state after l+1 draws = A^(l+1) * state + C_(l+1)
    uint32_t ja = 1103515245u, jc = 12345u;
    for (int i = 0; i < lane; i++) {
        jc = jc * 1103515245u + 12345u;
        ja = ja * 1103515245u;
    }
    auto jump = [&](uint32_t st, int k) -> uint32_t {  // state after k more draws (0 <= k <= 32)
        const int src = k > 0 ? k - 1 : 0;
        const uint32_t A = __shfl_sync(LFM_FULL, ja, src), C = __shfl_sync(LFM_FULL, jc, src);
        return k > 0 ? A * st + C : st;
    };
    uint32_t base = a.seed;  // rand_r state before the next draw (draw number qbase)
    // Candidate ring (BPR): which item a draw names does not depend on who consumes it, so the
    // item_ids loads of the next RDF_AHEAD draws are always in flight (cp.async into shared memory)
    // and only the membership test of a round is a dependent load.
    __shared__ int ring[RDF_RING];
    uint32_t fbase = a.seed;     // rand_r state before draw number `filled`
    uint32_t filled = 0, qbase = 0;
    auto fetch = [&](int count) {  // lanes < count request the candidates of draws filled + lane
        const uint32_t st = jump(fbase, lane + 1);
        if (lane < count) {
            const int r = (int)(lfm_temper(st) >> 1);
            rp_cp_async4(&ring[(filled + lane) & (RDF_RING - 1)], a.item_ids + (r % (int)n));
        }
        rp_commit();
        fbase = jump(fbase, count);
        filled += (uint32_t)count;
    };
    if (LOSS == LOSS_BPR && n > 0)
        for (int g = 0; g < RDF_AHEAD / 32; g++) fetch(32);
    unsigned long long c_neg = 0, c_rej = 0;
    int out_base = 0;
    bool bad = false;

    // tuple pipeline: the next chunk's tuples are loaded while this one is scheduled
    int nrow = 0, nuser = 0, nitem = 0;
    float ny = 0.f, nw = 0.f;
    auto load_tuple = [&](int64_t t) {
        if (t < n) {
            nrow = a.shuffle[t];
            nuser = a.user_ids[nrow];
            nitem = a.item_ids[nrow];
            ny = a.y[nrow];
            nw = a.sample_weight[nrow];
        }
    };
    load_tuple(lane);
    for (int64_t t0 = 0; t0 < n; t0 += 32) {
        const bool in = t0 + lane < n;
        const int user = nuser, item = nitem;
        const float y = ny, w = nw;
        load_tuple(t0 + 32 + lane);
        const bool valid = in && (LOSS == LOSS_LOGISTIC || y > 0);
        const unsigned V = __ballot_sync(LFM_FULL, valid);
        const int nvalid = __popc(V);
        const int rank = __popc(V & lt);
        int neg = -1;
        if (LOSS == LOSS_BPR) {
            int ps = 0, pe = 0;
            if (valid && !s.bitmap) {
                ps = a.pos.indptr[user];
                pe = a.pos.indptr[user + 1];
            }
            int start = 0;     // valid interactions of this chunk already given their negative
            int attempts = 0;  // draws the interaction at rank == start has already rejected
            while (start < nvalid) {
                const bool active = valid && rank >= start;
                const int k = active ? rank - start : 0;  // this lane's draw, counted from base
                rp_wait<RDF_AHEAD / 32 - 1>();  // the groups holding draws qbase .. qbase + 31 have landed
                __syncwarp();
                const int cand = ring[(qbase + (uint32_t)k) & (RDF_RING - 1)];
                bool mem = false;
                if (active) {
                    if (s.bitmap)
                        mem = (s.bitmap[(size_t)user * s.bitmap_words + (cand >> 5)] >> (cand & 31)) & 1u;
                    else
                        mem = lfm_bsearch(a.pos.indices, ps, pe, cand);
                }
                // T:1123-1127: the loop gives up after no_examples draws and keeps the last one
                const int64_t att = (rank == start) ? attempts : 0;
                const bool rej = active && mem && att < n - 1;
                const unsigned R = __ballot_sync(LFM_FULL, rej);
                int consumed, f;
                if (R == 0) {
                    f = nvalid;
                    consumed = nvalid - start;
                } else {
                    const int fl = __ffs(R) - 1;
                    f = __shfl_sync(LFM_FULL, rank, fl);
                    consumed = f - start + 1;
                    if (lane == fl) attempts = (rank == start ? attempts : 0) + 1;
                    c_rej++;
                }
                const bool accepted = active && rank < f;
                if (accepted) neg = cand;
                // a kept draw that IS a positive (the give-up case) may equal the positive item:
                // not a three-distinct-rows task any more
                if (__ballot_sync(LFM_FULL, accepted && mem)) bad = true;
                c_neg += (unsigned long long)consumed;
                start = f;
                base = jump(base, consumed);
                qbase += (uint32_t)consumed;
                __syncwarp();  // every lane has read its candidate before the ring moves on
                fetch(consumed);
            }
        }
        // versions: touches of my rows by earlier interactions (earlier chunks: counters; this
        // chunk: lanes below me)
        int eu = 0, ei = 0, en = 0;
        if (valid) {
            eu = rdf_ld_volatile(s.cnt_user + user);
            ei = rdf_ld_volatile(s.cnt_item + item);
            if (LOSS == LOSS_BPR) en = rdf_ld_volatile(s.cnt_item + neg);
        }
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const int uj = __shfl_sync(LFM_FULL, user, j), pj = __shfl_sync(LFM_FULL, item, j);
            const int nj = __shfl_sync(LFM_FULL, neg, j);
            if (((V >> j) & 1u) && j < lane && valid) {
                eu += (uj == user);
                ei += (pj == item);
                if (LOSS == LOSS_BPR) {
                    ei += (nj == item);
                    en += (pj == neg) + (nj == neg);
                }
            }
        }
        __syncwarp();  // every lane has read the counters before they move
        if (valid) {
            atomicAdd(s.cnt_user + user, 1);
            atomicAdd(s.cnt_item + item, 1);
            if (LOSS == LOSS_BPR) atomicAdd(s.cnt_item + neg, 1);
            RdfTask tk;
            tk.user = user; tk.item = item; tk.neg = neg;
            tk.eu = eu; tk.ei = ei; tk.en = en;
            tk.weight = w; tk.y = y;
            s.tasks[out_base + rank] = tk;
        }
        __threadfence();
        __syncwarp();
        out_base += nvalid;
    }
    if (lane == 0) {
        s.header[0] = bad ? -1 : out_base;
        a.scales->item_scale = 1.0;  // alpha == 0: the scales never leave 1
        a.scales->user_scale = 1.0;
        a.counters->positives = (unsigned long long)out_base;
        a.counters->negatives = c_neg;
        a.counters->updates = (unsigned long long)out_base;
        a.counters->rejected = c_rej;
    }
}

// ---- 2. execute ----------------------------------------------------------------------------
template <int K, int AD>
struct RdfRow {
    float w[K], g[K], m[K];
    float b, bg, bm;
};

template <int K, int AD>
__device__ __forceinline__ void rdf_load_row(RdfRow<K, AD>& r, const DevTable& t, int id, int d, int lane) {
    const size_t o = (size_t)id * d;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = lane + 32 * k;
        r.w[k] = j < d ? __ldcg(t.w + o + j) : 0.0f;
        r.g[k] = j < d ? __ldcg(t.g + o + j) : 1.0f;
        r.m[k] = (AD && j < d) ? __ldcg(t.m + o + j) : 0.0f;
    }
    r.b = __ldcg(t.b + id);
    r.bg = __ldcg(t.bg + id);
    r.bm = AD ? __ldcg(t.bm + id) : 0.0f;
}

template <int K, int AD>
__device__ __forceinline__ void rdf_store_row(const RdfRow<K, AD>& r, const DevTable& t, int id, int d, int lane,
                                              int bias_lane) {
    const size_t o = (size_t)id * d;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = lane + 32 * k;
        if (j < d) {
            t.w[o + j] = r.w[k];
            t.g[o + j] = r.g[k];
            if (AD) t.m[o + j] = r.m[k];
        }
    }
    if (lane == bias_lane) {
        t.b[id] = r.b;
        t.bg[id] = r.bg;
        if (AD) t.bm[id] = r.bm;
    }
}

template <int LOSS, int K, int AD>
__global__ void __launch_bounds__(RDF_WARPS * 32) rdf_execute_kernel(FitArgs a, RdfScratch s) {
    extern __shared__ __align__(16) float rdf_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    DevModel& m = a.model;
    const int d = m.d;
    float* su = rdf_smem + (size_t)wib * 3 * (d + 1);
    float* sp = su + (d + 1);
    float* sn = sp + (d + 1);
    const int n_tasks = s.header[0];  // -1: nothing to do here
    const int W = gridDim.x * RDF_WARPS;
    const float fw = (float)((double)1.0f * 1.0);  // f32(double(1.0f) * scale), scale == 1
    const double lr = (double)m.lr;
    const int nwait = LOSS == LOSS_BPR ? 3 : 2;

    // consecutive tasks go to different SMs: the runnable ones are always the earliest ones
    int t = wib * gridDim.x + blockIdx.x;
    RdfTask tk;
    if (t < n_tasks) tk = s.tasks[t];
    for (; t < n_tasks; t += W) {
        RdfTask nx = tk;
        if (t + W < n_tasks) nx = s.tasks[t + W];
        if (lane < nwait) {
            const int32_t* p = lane == 0 ? s.ver_user + tk.user : (lane == 1 ? s.ver_item + tk.item : s.ver_item + tk.neg);
            const int e = lane == 0 ? tk.eu : (lane == 1 ? tk.ei : tk.en);
            // a version that never arrives would be a scheduling bug: give up loudly, not forever
            for (unsigned spins = 0; rdf_ld_acquire(p) != e; spins++)
                if (spins > (1u << 23)) {
                    atomicExch(s.header + 1, 1);
                    break;
                }
        }
        __syncwarp();
        RdfRow<K, AD> U, P, N;
        rdf_load_row<K, AD>(U, m.user, tk.user, d, lane);
        rdf_load_row<K, AD>(P, m.item, tk.item, d, lane);
        if (LOSS == LOSS_BPR) rdf_load_row<K, AD>(N, m.item, tk.neg, d, lane);
        // representations (T:302-317 with the single identity feature): 0.0f + fw * E
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int j = lane + 32 * k;
            if (j < d) {
                su[j] = 0.0f + fw * U.w[k];
                sp[j] = 0.0f + fw * P.w[k];
                if (LOSS == LOSS_BPR) sn[j] = 0.0f + fw * N.w[k];
            }
        }
        if (lane == 0) {
            su[d] = 0.0f + fw * U.b;
            sp[d] = 0.0f + fw * P.b;
            if (LOSS == LOSS_BPR) sn[d] = 0.0f + fw * N.b;
        }
        __syncwarp();
        if (LOSS == LOSS_BPR) {
            const double pp = (double)score(su, sp, d);
            const double np = (double)score(su, sn, d);
            const double loss = (double)tk.weight * (1.0 - (double)sigmoid_ref((float)(pp - np)));  // T:1160-1165
            // warp_update (T:537-649): biases, then per component positive / negative / user
            if (lane == 0) step(&P.b, &P.bg, &P.bm, (double)1.0f, -loss, AD, lr, 0.0, m.rho, m.eps);
            if (lane == 1) step(&N.b, &N.bg, &N.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
            if (lane == 2) step(&U.b, &U.bg, &U.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int j = lane + 32 * k;
                if (j < d) {
                    const float uc = su[j], pc = sp[j], nc = sn[j];
                    step(&P.w[k], &P.g[k], &P.m[k], (double)1.0f, (-loss) * (double)uc, AD, lr, 0.0, m.rho, m.eps);
                    step(&N.w[k], &N.g[k], &N.m[k], (double)1.0f, loss * (double)uc, AD, lr, 0.0, m.rho, m.eps);
                    step(&U.w[k], &U.g[k], &U.m[k], (double)1.0f, loss * (double)(float)(nc - pc), AD, lr, 0.0,
                         m.rho, m.eps);
                }
            }
            rdf_store_row<K, AD>(P, m.item, tk.item, d, lane, 0);
            rdf_store_row<K, AD>(N, m.item, tk.neg, d, lane, 1);
            rdf_store_row<K, AD>(U, m.user, tk.user, d, lane, 2);
        } else {
            const double prediction = (double)sigmoid_ref(score(su, sp, d));  // T:745-760
            const int y = (tk.y <= 0) ? 0 : 1;
            const double loss = (double)tk.weight * (prediction - (double)y);
            // update (T:454-534): item bias, user bias, then per component item / user
            if (lane == 0) step(&P.b, &P.bg, &P.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
            if (lane == 1) step(&U.b, &U.bg, &U.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int j = lane + 32 * k;
                if (j < d) {
                    const float uc = su[j], ic = sp[j];
                    step(&P.w[k], &P.g[k], &P.m[k], (double)1.0f, loss * (double)uc, AD, lr, 0.0, m.rho, m.eps);
                    step(&U.w[k], &U.g[k], &U.m[k], (double)1.0f, loss * (double)ic, AD, lr, 0.0, m.rho, m.eps);
                }
            }
            rdf_store_row<K, AD>(P, m.item, tk.item, d, lane, 0);
            rdf_store_row<K, AD>(U, m.user, tk.user, d, lane, 1);
        }
        __threadfence();  // this lane's stores are visible device-wide ...
        __syncwarp();     // ... for every lane, before the versions move
        if (lane < nwait) {
            int32_t* p = lane == 0 ? s.ver_user + tk.user : (lane == 1 ? s.ver_item + tk.item : s.ver_item + tk.neg);
            const int e = lane == 0 ? tk.eu : (lane == 1 ? tk.ei : tk.en);
            rdf_st_release(p, e + 1);
        }
        tk = nx;
    }
}

}  // namespace

// Scratch the dataflow path needs for (loss, a); 0 when (loss, a) is outside its scope.
static size_t rdf_scratch_bytes(int loss, const FitArgs& a, int64_t bitmap_limit_bytes) {
    const DevModel& m = a.model;
    if ((loss != LOSS_BPR && loss != LOSS_LOGISTIC) || !a.itf.identity || !a.usf.identity || a.item_alpha != 0.0 ||
        a.user_alpha != 0.0 || m.d > 256 || m.d < 1 || a.n > 0x7fffffffLL || a.n < 1)
        return 0;
    if (loss == LOSS_BPR && !a.pos.indptr) return 0;
    size_t b = 256 + sizeof(int32_t) * 2 * ((size_t)m.user.n + (size_t)m.item.n + 64) + sizeof(RdfTask) * (size_t)a.n + 256;
    if (loss == LOSS_BPR) {
        const size_t words = ((size_t)a.pos.cols + 31) / 32;
        const size_t bm = sizeof(uint32_t) * words * (size_t)a.pos.rows;
        if ((int64_t)bm <= bitmap_limit_bytes) b += bm + 256;
    }
    return b;
}

// Returns cudaErrorNotSupported when the epoch has to run in replay_kernel instead (out of scope,
// or the schedule found the one case it does not represent); nothing has been modified then.
static cudaEvent_t g_rdf_ev[3] = {nullptr, nullptr, nullptr};
static double g_rdf_ms[2] = {0.0, 0.0};  // schedule kernel, execute kernel of the last dataflow epoch
static int g_rdf_tasks = -1;

static cudaError_t lfm_try_launch_replay_dataflow(int loss, const FitArgs& a, cudaStream_t st) {
    if (!a.replay_scratch || a.replay_scratch_bytes == 0) return cudaErrorNotSupported;
    for (int i = 0; i < 3; i++)
        if (!g_rdf_ev[i]) {
            cudaError_t ee = cudaEventCreate(&g_rdf_ev[i]);
            if (ee != cudaSuccess) return ee;
        }
    const DevModel& m = a.model;
    const int d = m.d;
    const size_t nu = (size_t)m.user.n, ni = (size_t)m.item.n;
    unsigned char* base = (unsigned char*)a.replay_scratch;
    RdfScratch s;
    s.header = (int32_t*)base;
    s.cnt_user = (int32_t*)(base + 256);
    s.cnt_item = s.cnt_user + nu;
    s.ver_user = s.cnt_item + ni;
    s.ver_item = s.ver_user + nu;
    size_t off = 256 + sizeof(int32_t) * 2 * (nu + ni + 64);
    off = (off + 255) & ~(size_t)255;
    s.tasks = (RdfTask*)(base + off);
    off += sizeof(RdfTask) * (size_t)a.n;
    off = (off + 255) & ~(size_t)255;
    s.bitmap = nullptr;
    s.bitmap_words = 0;
    cudaError_t e = cudaMemsetAsync(base, 0, 256 + sizeof(int32_t) * 2 * (nu + ni + 64), st);
    if (e != cudaSuccess) return e;
    if (loss == LOSS_BPR) {
        const size_t words = ((size_t)a.pos.cols + 31) / 32;
        const size_t bm = sizeof(uint32_t) * words * (size_t)a.pos.rows;
        if (off + bm <= a.replay_scratch_bytes) {
            e = lfm_launch_build_bitmap(a.pos, (uint32_t*)(base + off), (int32_t)words, st);
            if (e != cudaSuccess) return e;
            s.bitmap = (const uint32_t*)(base + off);
            s.bitmap_words = (int32_t)words;
        }
    }
    cudaEventRecord(g_rdf_ev[0], st);
    if (loss == LOSS_BPR) rdf_schedule_kernel<LOSS_BPR><<<1, 32, 0, st>>>(a, s);
    else rdf_schedule_kernel<LOSS_LOGISTIC><<<1, 32, 0, st>>>(a, s);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cudaEventRecord(g_rdf_ev[1], st);

    const size_t smem = sizeof(float) * 3 * (d + 1) * RDF_WARPS;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    void* args[2] = {(void*)&a, (void*)&s};
#define RDF_LAUNCH(L, KK, AA)                                                                              \
    do {                                                                                                   \
        auto kern = rdf_execute_kernel<L, KK, AA>;                                                         \
        int per_sm = 0;                                                                                    \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RDF_WARPS * 32, smem);            \
        if (e != cudaSuccess) return e;                                                                    \
        if (per_sm < 1) return cudaErrorNotSupported;                                                      \
        if (per_sm > 2) per_sm = 2; /* a few thousand warps: far more than the graph is wide */            \
        int64_t blocks = (int64_t)per_sm * sms;                                                            \
        const int64_t need = (a.n + RDF_WARPS - 1) / RDF_WARPS;                                            \
        if (blocks > need) blocks = need;                                                                  \
        e = cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)blocks), dim3(RDF_WARPS * 32), args, \
                                        smem, st);                                                         \
    } while (0)
#define RDF_BY_K(L, AA)                       \
    do {                                      \
        if (d <= 32) RDF_LAUNCH(L, 1, AA);    \
        else if (d <= 64) RDF_LAUNCH(L, 2, AA); \
        else if (d <= 128) RDF_LAUNCH(L, 4, AA); \
        else RDF_LAUNCH(L, 8, AA);            \
    } while (0)
    if (loss == LOSS_BPR) {
        if (m.adadelta) RDF_BY_K(LOSS_BPR, 1);
        else RDF_BY_K(LOSS_BPR, 0);
    } else {
        if (m.adadelta) RDF_BY_K(LOSS_LOGISTIC, 1);
        else RDF_BY_K(LOSS_LOGISTIC, 0);
    }
#undef RDF_BY_K
#undef RDF_LAUNCH
    if (e != cudaSuccess) return e;
    cudaEventRecord(g_rdf_ev[2], st);
    // the schedule's verdict (one int): -1 means it met the give-up case of T:1123-1127 and the
    // execute kernel did nothing
    int32_t hdr[2] = {0, 0};
    e = cudaMemcpyAsync(hdr, s.header, sizeof(hdr), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (hdr[1] != 0) return cudaErrorLaunchTimeout;  // a task waited for a version that never came
    const int32_t verdict = hdr[0];
    float ms0 = 0.f, ms1 = 0.f;
    cudaEventElapsedTime(&ms0, g_rdf_ev[0], g_rdf_ev[1]);
    cudaEventElapsedTime(&ms1, g_rdf_ev[1], g_rdf_ev[2]);
    g_rdf_ms[0] = ms0;
    g_rdf_ms[1] = ms1;
    g_rdf_tasks = verdict;
    return verdict < 0 ? cudaErrorNotSupported : cudaSuccess;
}
