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
// One cooperative kernel (rdf_kernel), two roles:
//   * the SCHEDULER (three warps of CTA 0, which has an SM to itself, pipelined over chunks of 32
//     interactions) walks the shuffled list in order.  Stage S consumes the rand_r stream exactly as
//     the reference does (the state of draw k is an LCG jump-ahead, the candidates' item ids stream
//     in ahead of time through a cp.async ring, several draws are judged per interaction per round
//     and the lanes resolved in order); stage V counts, for every row an interaction touches, how
//     many earlier interactions touch it (its `version`); stage E emits one task per interaction:
//     {user, item, negative, versions};
//   * the EXECUTORS (every other SM, one warp per task, tasks handed out in visiting order) wait
//     until the task exists and its rows have reached their versions, run the reference's
//     arithmetic on them and publish version + 1 (st.release).  The earliest unfinished task never
//     waits on another task, and the scheduler never waits at all, so the walk always advances
//     (cooperative launch: every warp is resident).
//
// The result equals replay_kernel<BPR / LOGISTIC> bit for bit (tests/test_gpu_replay_dataflow.py),
// which in turn is the oracle's / the reference's single-thread result.
#pragma once

namespace {

struct __align__(16) RdfTask {  // 32 B
    int32_t user, item, neg;  // neg: BPR only; -1 = the kept draw is the positive item itself
    int32_t eu, ei, en;       // versions the user / item / negative rows must have reached
    float weight, y;
};

// header words (device scratch, zeroed before the launch)
enum { RDF_H_TOTAL = 0, RDF_H_STALL = 1, RDF_H_PRODUCED = 2, RDF_H_DONE = 3, RDF_H_SCHED_US = 4 };

struct RdfScratch {
    int32_t* header;    // see RDF_H_*
    int32_t* cnt_user;  // scheduler: touches so far             [n_users]   (global fallback)
    int32_t* cnt_item;  //                                       [n_items]
    int32_t* ver_user;  // executors: published row versions     [n_users]
    int32_t* ver_item;  //                                       [n_items]
    int32_t* mark;      // scheduler (BPR): chunk marks            [2 * n_items] (global fallback)
    RdfTask* tasks;     // [n]
    const Tuple* tuples;  // [n] the interactions in visiting order (pack_kernel), user < 0: skipped
    const uint32_t* bitmap;  // optional exact membership bitmap of the positives CSR (else null)
    int32_t bitmap_words;
    int32_t cnt_in_smem;     // 1: the touch counters live in the scheduler CTA's shared memory
    int32_t bitmap_in_smem;  // 1: ... and a copy of the bitmap behind them
    uint64_t mod_magic;      // ceil(2^64 / n): draw % n without a division (Lemire 2019)
};

#define RDF_WARPS 8      // warps per CTA
#define RDF_RING 256     // candidate ring entries (scheduler)
#define RDF_AHEAD 224    // draws requested ahead of the one being judged (7 cp.async groups)
#define RDF_W 8          // draws judged per interaction per round
#define RDF_WAIT 4       // cp.async groups that may still be in flight when a round reads the ring:
                         // a round consumes <= 32 + W draws and reads < qbase + 32 + W, (RDF_WAIT + 1) * 40 <= RDF_AHEAD
#define RDF_PUBLISH 8    // the scheduler publishes its progress every RDF_PUBLISH chunks of 32
#define RDF_PF 16         // chunks of the packed list prefetched into L2 ahead of the scheduler

__device__ __forceinline__ int rdf_ld_acquire(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int rdf_ld_relaxed(const int32_t* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rdf_st_release(int32_t* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long rdf_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- scheduler: three warps of CTA 0 in a pipeline -------------------------------------------
// S (sample): streams the packed list, gives every BPR interaction its negative (rand_r replay).
// V (version): counts the earlier touches of every row an interaction touches.
// E (emit): writes the tasks and publishes the scheduler's progress to the executors.
// The stages hand chunks of 32 interactions to each other through two small shared-memory rings;
// each stage is sequential in itself (that is what makes the result the reference's), but the
// three run concurrently on different chunks.
#define RDF_PIPE 3  // chunks a ring holds
struct RdfChunkA {  // S -> V
    int32_t user[32], item[32], neg[32];
    float w[32], y[32];
};
struct RdfChunkB {  // V -> E
    int32_t user[32], item[32], neg[32], eu[32], ei[32], en[32];
    float w[32], y[32];
};
struct RdfPipe {
    RdfChunkA a[RDF_PIPE];
    RdfChunkB b[RDF_PIPE];
    int32_t ring[RDF_RING];          // candidate ring of stage S
    volatile int32_t a_prod, a_cons, b_prod, b_cons;  // chunk counters of the two rings
};

// `fence`: a block-scope fence after the flag has been seen.  Stage V passes false: a fence there would
// wait for its outstanding (prefetching) global loads; it relies on the SM executing the shared-memory
// accesses of its warps in issue order (volatile flags, __syncwarp between data and flag), as the
// other side of each hand-off still fences.
__device__ __forceinline__ void rdf_wait_ge(volatile int32_t* flag, int v, long long& waited, bool fence = true) {
    const long long t0 = clock64();
    if (threadIdx.x % 32 == 0)
        while (*flag < v) __nanosleep(64);
    if (fence) __threadfence_block();  // what the other stage wrote before the flag is read after it
    __syncwarp();
    waited += clock64() - t0;
}

template <int LOSS>
__device__ __forceinline__ void rdf_stage_sample(const FitArgs& a, const RdfScratch& s, RdfPipe* pp, uint32_t* bitmap_smem) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t n = a.n;
    int32_t* ring = pp->ring;
    long long waited = 0;
    const long long t_begin = clock64();
    // membership bitmap: a copy in shared memory when the host found room for it
    const uint32_t* bitmap = s.bitmap;
    if (LOSS == LOSS_BPR && s.bitmap && s.bitmap_in_smem) {
        const int total = a.pos.rows * s.bitmap_words;
        for (int i = lane; i < total; i += 32) bitmap_smem[i] = s.bitmap[i];
        __syncwarp();
        bitmap = bitmap_smem;
    }
    // draws judged per interaction per round: the whole window when a test is a shared-memory load,
    // two when it is a random DRAM / L2 access, one when it is a binary search
    const int wuse = (LOSS == LOSS_BPR && s.bitmap) ? (s.bitmap_in_smem ? RDF_W : 2) : 1;
    // LCG jump-ahead (rand_r's state update, T:76): lane l holds (A^(l+1), C_(l+1)) with
    // state after l+1 draws = A^(l+1) * state + C_(l+1)
    uint32_t ja = 1103515245u, jc = 12345u;
    for (int i = 0; i < lane; i++) {
        jc = jc * 1103515245u + 12345u;
        ja = ja * 1103515245u;
    }
    auto jump = [&](uint32_t st, int k) -> uint32_t {  // state after k more draws (0 <= k <= 32), k warp-uniform
        const int src = k > 0 ? k - 1 : 0;
        const uint32_t A = __shfl_sync(LFM_FULL, ja, src), C = __shfl_sync(LFM_FULL, jc, src);
        return k > 0 ? A * st + C : st;
    };
    // Candidate ring (BPR): which item a draw names does not depend on who consumes it, so the
    // item_ids loads of the next RDF_AHEAD draws are always in flight (cp.async into shared memory)
    // and only the membership test of a round is a dependent load.
    uint32_t fbase = a.seed;     // rand_r state before draw number `filled`
    uint32_t filled = 0, qbase = 0;
    auto fetch = [&](int count) {  // request the candidates of draws filled .. filled + count - 1 (count <= 64): one group
        const uint32_t f32 = jump(fbase, count > 32 ? 32 : 0);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (lane + 32 * h < count) {
                const uint32_t st = ja * (h ? f32 : fbase) + jc;     // state after lane + 32 h + 1 draws
                const uint32_t r = lfm_temper(st) >> 1;              // rand_r's value (T:77-81)
                const uint32_t idx = (uint32_t)__umul64hi(s.mod_magic * (uint64_t)r, (uint64_t)n);  // r % n
                rp_cp_async4(&ring[(filled + lane + 32 * h) & (RDF_RING - 1)], a.item_ids + idx);
            }
        }
        rp_commit();
        fbase = jump(count > 32 ? f32 : fbase, count > 32 ? count - 32 : count);
        filled += (uint32_t)count;
    };
    if (LOSS == LOSS_BPR && n > 0)
        for (int g = 0; g < RDF_AHEAD / 32; g++) fetch(32);
    unsigned long long c_neg = 0, c_rej = 0;
    long long rounds = 0;

    // tuple pipeline: the list was packed in visiting order by pack_kernel, so the scheduler streams
    // it (one 16 B tuple per lane per chunk), three chunks ahead in registers and RDF_PF chunks ahead in L2
    const int4* tup = (const int4*)s.tuples;
    auto load_tuple = [&](int64_t t) -> int4 { return t < n ? __ldcg(tup + t) : make_int4(-1, 0, 0, 0); };
    auto prefetch = [&](int64_t t) {
        if (t < n && (lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(tup + t));
    };
#pragma unroll 1
    for (int c = 0; c < RDF_PF; c++) prefetch((int64_t)c * 32 + lane);
    int4 tp1 = load_tuple(lane), tp2 = load_tuple(32 + lane), tp3 = load_tuple(64 + lane);
    int chunk = 0;
    for (int64_t t0 = 0; t0 < n; t0 += 32, chunk++) {
        const int user = tp1.x, item = tp1.y;
        const int wbits = tp1.z, ybits = tp1.w;
        tp1 = tp2;
        tp2 = tp3;
        tp3 = load_tuple(t0 + 96 + lane);
        prefetch(t0 + (int64_t)RDF_PF * 32 + lane);
        const bool valid = user >= 0;  // beyond the list, or one of BPR's Y <= 0 interactions (T:1112-1113: pack_kernel marks them)
        const unsigned V = __ballot_sync(LFM_FULL, valid);
        const int nvalid = __popc(V);
        const int rank = __popc(V & lt);
        int neg = -1;
        if (LOSS == LOSS_BPR) {
            int ps = 0, pe = 0;
            const uint32_t* brow = nullptr;
            if (valid) {
                if (bitmap) {
                    brow = bitmap + (size_t)user * s.bitmap_words;
                } else {
                    ps = a.pos.indptr[user];
                    pe = a.pos.indptr[user + 1];
                }
            }
            int start = 0;     // valid interactions of this chunk already given their negative
            int attempts = 0;  // draws the interaction at rank == start has already rejected
            // One round judges `wuse` consecutive draws per interaction: lane r (counted from `start`)
            // would take draw r if nobody before it rejected anything, draw r + s after s rejections;
            // its membership bits for draws r .. r + wuse - 1 turn an incoming shift into an outgoing
            // one, and the lanes are resolved in order.  A chunk is done in one round unless more than
            // wuse - 1 rejections pile up before some lane.
            while (start < nvalid) {
                rounds++;
                const bool active = valid && rank >= start;
                const int k0 = active ? rank - start : 0;
                rp_wait<RDF_WAIT>();  // the groups holding draws qbase .. qbase + 31 + W have landed
                __syncwarp();
                unsigned M = 0;  // bit j: draw k0 + j names one of my positives (j < wuse)
                if (active) {
                    if (brow) {
                        uint32_t word[RDF_W];
                        int bit[RDF_W];
#pragma unroll
                        for (int j = 0; j < RDF_W; j++) {  // all loads of the window in flight together
                            const int cand = ring[(qbase + (uint32_t)(k0 + j)) & (RDF_RING - 1)];
                            bit[j] = cand & 31;
                            word[j] = j < wuse ? brow[cand >> 5] : 0u;
                        }
#pragma unroll
                        for (int j = 0; j < RDF_W; j++) M |= ((word[j] >> bit[j]) & 1u) << j;
                    } else {
                        for (int j = 0; j < wuse; j++) {
                            const int cand = ring[(qbase + (uint32_t)(k0 + j)) & (RDF_RING - 1)];
                            M |= (lfm_bsearch(a.pos.indices, ps, pe, cand) ? 1u : 0u) << j;
                        }
                    }
                }
                // T:1123-1127: the loop gives up after no_examples draws and keeps the last one, so a
                // member draw is a rejection only while fewer than n - 1 draws were rejected before it
                int64_t lim64 = n - 1 - ((rank == start) ? attempts : 0);
                const int lim = lim64 > 64 ? 64 : (lim64 < 0 ? 0 : (int)lim64);
                // resolve the lanes in order: incoming shift -> my rejections -> outgoing shift
                const unsigned packed = M | ((unsigned)lim << 8);
                // Only a lane whose draw at the current shift is a positive changes the shift, so the
                // walk jumps from one such lane to the next: lane s keeps B_s = the lanes whose draw
                // k0 + s is a member, and the next event is the lowest bit of B_shift at or above `cur`.
                unsigned myB = 0u;
#pragma unroll
                for (int sb = 0; sb < RDF_W; sb++) {
                    const unsigned b = __ballot_sync(LFM_FULL, active && ((M >> sb) & 1u));
                    if (lane == sb) myB = b;
                }
                int sft = 0, my_in = 0, over_lane = -1, over_s = 0, cur = 0;
                while (cur < 32) {
                    const unsigned Bs = __shfl_sync(LFM_FULL, myB, sft) & ~((1u << cur) - 1u);
                    if (Bs == 0u) break;
                    const int j = __ffs(Bs) - 1;  // the next lane that rejects its first draw
                    if (lane >= cur && lane <= j) my_in = sft;
                    const unsigned pj = __shfl_sync(LFM_FULL, packed, j);
                    const int run = __ffs(~((pj & 0xffu) >> sft)) - 1;  // consecutive member draws from draw k0 + sft
                    const int lj = (int)(pj >> 8);
                    const int v = sft + (run < lj ? run : lj);
                    if (v > wuse - 1) { over_lane = j; over_s = sft; break; }
                    sft = v;
                    cur = j + 1;
                }
                if (over_lane < 0 && lane >= cur) my_in = sft;
                const bool kept = active && (over_lane < 0 || lane < over_lane);
                bool kept_mem = false;
                if (kept) {
                    const int run = __ffs(~(M >> my_in)) - 1;
                    const int out = my_in + (run < lim ? run : lim);  // my accepted draw is k0 + out
                    const int cand = ring[(qbase + (uint32_t)(k0 + out)) & (RDF_RING - 1)];
                    kept_mem = (M >> out) & 1u;
                    neg = (cand == item) ? -1 : cand;  // give-up case only: the draw may be the positive itself
                }
                int consumed;
                if (over_lane < 0) {
                    consumed = (nvalid - start) + sft;  // one kept draw per interaction + the rejected ones
                    c_rej += (unsigned long long)sft;
                    start = nvalid;
                } else {
                    const int f = __shfl_sync(LFM_FULL, rank, over_lane);
                    consumed = (f - start) + wuse;      // ... and the whole window of the lane that ran out
                    c_rej += (unsigned long long)wuse;  // over_s by the lanes before it, wuse - over_s by itself
                    if (lane == over_lane) attempts = (rank == start ? attempts : 0) + (wuse - over_s);
                    start = f;
                }
                c_rej += __popc(__ballot_sync(LFM_FULL, kept && kept_mem));  // replay_kernel counts a kept member draw too
                c_neg += (unsigned long long)consumed;
                qbase += (uint32_t)consumed;
                __syncwarp();  // every lane has read its candidates before the ring moves on
                fetch(consumed);
            }
        }
        // hand the chunk to stage V
        rdf_wait_ge(&pp->a_cons, chunk - (RDF_PIPE - 1), waited);
        RdfChunkA& o = pp->a[chunk % RDF_PIPE];
        o.user[lane] = user; o.item[lane] = item; o.neg[lane] = neg;
        o.w[lane] = __int_as_float(wbits); o.y[lane] = __int_as_float(ybits);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) pp->a_prod = chunk + 1;
    }
    if (lane == 0) {
        a.counters->negatives = c_neg;
        a.counters->rejected = c_rej;
        s.header[8] = (int32_t)((clock64() - t_begin - waited) >> 10);  // kilo-cycles of work
        s.header[9] = (int32_t)(waited >> 10);
        s.header[14] = (int32_t)rounds;
    }
}

template <int LOSS>
__device__ __forceinline__ void rdf_stage_version(const FitArgs& a, const RdfScratch& s, RdfPipe* pp, int32_t* cnt_smem) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nchunks = (int)((a.n + 31) / 32);
    long long waited = 0;
    const long long t_begin = clock64();
    int32_t* cnt_user = s.cnt_user;
    int32_t* cnt_item = s.cnt_item;
    // mark arrays of the BPR version step (item id -> last chunk that touched it as positive / negative)
    int32_t* mark_pos = s.mark;
    int32_t* mark_neg = s.mark + a.model.item.n;
    if (s.cnt_in_smem) {
        cnt_user = cnt_smem;
        cnt_item = cnt_smem + a.model.user.n;
        const int total = a.model.user.n + a.model.item.n + (LOSS == LOSS_BPR ? 2 * a.model.item.n : 0);
        for (int i = lane; i < total; i += 32) cnt_smem[i] = 0;
        if (LOSS == LOSS_BPR) {
            mark_pos = cnt_smem + a.model.user.n + a.model.item.n;
            mark_neg = mark_pos + a.model.item.n;
        }
        __syncwarp();
    }
    // A chunk's touch counters are requested right after the previous chunk has stored its own
    // (same SM: the loads see those stores), so their latency runs under the hand-off to stage E
    // instead of at the head of the next iteration.
    int user = -1, item = 0, neg = -1, eu = 0, ei = 0, en = 0;
    float w = 0.f, y = 0.f;
    auto take = [&](int chunk) {  // read chunk `chunk` from ring A (it must be there), free its slot, request its counters
        const RdfChunkA& in = pp->a[chunk % RDF_PIPE];
        user = in.user[lane]; item = in.item[lane]; neg = in.neg[lane];
        w = in.w[lane]; y = in.y[lane];
        __syncwarp();
        if (lane == 0) pp->a_cons = chunk + 1;
        eu = ei = en = 0;
        if (user >= 0) {
            eu = cnt_user[user];
            ei = cnt_item[item];
            if (LOSS == LOSS_BPR && neg >= 0) en = cnt_item[neg];
        }
    };
    bool have = false;  // chunk `chunk` was already taken at the end of the previous iteration
    for (int chunk = 0; chunk < nchunks; chunk++) {
        if (!have) {
            rdf_wait_ge(&pp->a_prod, chunk + 1, waited, false);
            take(chunk);
        }
        const int cur_user = user, cur_item = item, cur_neg = neg;
        const float cur_w = w, cur_y = y;
        const bool valid = user >= 0;
        const unsigned V = __ballot_sync(LFM_FULL, valid);
        // versions: + touches by the lanes below me; and am I the last lane of the chunk on each row
        bool lu = true, li = true, ln = true;
        {
            const unsigned gt = ~lt & ~(1u << lane);
            const unsigned mu = __match_any_sync(LFM_FULL, valid ? user : -1 - lane) & V;  // lanes on my user row
            eu += __popc(mu & lt);
            lu = (mu & gt) == 0;
            if (LOSS == LOSS_LOGISTIC) {
                const unsigned mi = __match_any_sync(LFM_FULL, valid ? item : -1 - lane) & V;
                ei += __popc(mi & lt);
                li = (mi & gt) == 0;
            }
        }
        if (LOSS == LOSS_BPR) {
            // an item row is touched as a positive or as a negative.  Same-kind touches: match.any.
            // Cross touches (my positive = someone's negative or the reverse) are found through two
            // mark arrays indexed by item id, and only the lanes involved in one are walked.
            const unsigned gt = ~lt & ~(1u << lane);
            const int cid = chunk + 1;
            const unsigned mp = __match_any_sync(LFM_FULL, valid ? item : -1 - lane) & V;
            const unsigned mn = __match_any_sync(LFM_FULL, (valid && neg >= 0) ? neg : -1 - lane) & V;
            ei += __popc(mp & lt);
            li = (mp & gt) == 0;
            en += __popc(mn & lt);
            ln = (mn & gt) == 0;
            if (valid) {
                mark_pos[item] = cid;
                if (neg >= 0) mark_neg[neg] = cid;
            }
            __syncwarp();
            const bool cross = valid && ((neg >= 0 && mark_pos[neg] == cid) || mark_neg[item] == cid);
            unsigned C = __ballot_sync(LFM_FULL, cross);
            while (C) {
                const int j = __ffs(C) - 1;
                C &= C - 1;
                const int pj = __shfl_sync(LFM_FULL, item, j), nj = __shfl_sync(LFM_FULL, neg, j);
                if (valid && j != lane) {
                    const bool x1 = nj == item;              // lane j's negative is my positive
                    const bool x2 = neg >= 0 && pj == neg;   // lane j's positive is my negative
                    if (j < lane) {
                        ei += x1; en += x2;
                    } else {
                        li = li && !x1; ln = ln && !x2;
                    }
                }
            }
        }
        __syncwarp();  // every lane has read the counters before they move
        if (valid) {
            // one store per row: the last lane on it writes the count after this chunk
            if (lu) cnt_user[user] = eu + 1;
            if (li) cnt_item[item] = ei + 1;
            if (LOSS == LOSS_BPR && neg >= 0 && ln) cnt_item[neg] = en + 1;
        }
        const int out_eu = eu, out_ei = ei, out_en = en;
        __syncwarp();  // the stores above are ordered before the next chunk's counter loads
        // the next chunk, if stage S has it ready: take it now
        have = false;
        if (LOSS == LOSS_BPR && chunk + 1 < nchunks) {  // (logistic: measured slower, its version step is two match.any)
            int ready = 0;
            if (lane == 0) ready = pp->a_prod >= chunk + 2;
            ready = __shfl_sync(LFM_FULL, ready, 0);
            if (ready) {
                take(chunk + 1);
                have = true;
            }
        }
        // hand the chunk to stage E
        rdf_wait_ge(&pp->b_cons, chunk - (RDF_PIPE - 1), waited, false);
        RdfChunkB& o = pp->b[chunk % RDF_PIPE];
        o.user[lane] = cur_user; o.item[lane] = cur_item; o.neg[lane] = cur_neg;
        o.eu[lane] = out_eu; o.ei[lane] = out_ei; o.en[lane] = out_en;
        o.w[lane] = cur_w; o.y[lane] = cur_y;
        __syncwarp();  // also: counters (same SM: shared memory / L1) written before the next chunk reads them
        if (lane == 0) pp->b_prod = chunk + 1;
    }
    if (lane == 0) {
        s.header[10] = (int32_t)((clock64() - t_begin - waited) >> 10);
        s.header[11] = (int32_t)(waited >> 10);
    }
}

__device__ __forceinline__ void rdf_stage_emit(const FitArgs& a, const RdfScratch& s, RdfPipe* pp) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nchunks = (int)((a.n + 31) / 32);
    long long waited = 0;
    const long long t_begin = clock64();
    const unsigned long long ns_begin = rdf_now_ns();
    int out_base = 0;
    for (int chunk = 0; chunk < nchunks; chunk++) {
        rdf_wait_ge(&pp->b_prod, chunk + 1, waited);
        const RdfChunkB& in = pp->b[chunk % RDF_PIPE];
        RdfTask tk;
        tk.user = in.user[lane]; tk.item = in.item[lane]; tk.neg = in.neg[lane];
        tk.eu = in.eu[lane]; tk.ei = in.ei[lane]; tk.en = in.en[lane];
        tk.weight = in.w[lane]; tk.y = in.y[lane];
        __syncwarp();
        if (lane == 0) pp->b_cons = chunk + 1;
        const bool valid = tk.user >= 0;
        const unsigned V = __ballot_sync(LFM_FULL, valid);
        if (valid) s.tasks[out_base + __popc(V & lt)] = tk;
        out_base += __popc(V);
        __syncwarp();
        if ((chunk % RDF_PUBLISH) == RDF_PUBLISH - 1 && lane == 0) rdf_st_release(s.header + RDF_H_PRODUCED, out_base);
    }
    __syncwarp();
    if (lane == 0) {
        a.scales->item_scale = 1.0;  // alpha == 0: the scales never leave 1
        a.scales->user_scale = 1.0;
        a.counters->positives = (unsigned long long)out_base;
        a.counters->updates = (unsigned long long)out_base;
        s.header[RDF_H_TOTAL] = out_base;
        s.header[RDF_H_SCHED_US] = (int32_t)((rdf_now_ns() - ns_begin) / 1000ull);
        s.header[12] = (int32_t)((clock64() - t_begin - waited) >> 10);
        s.header[13] = (int32_t)(waited >> 10);
        s.header[15] = nchunks;
        rdf_st_release(s.header + RDF_H_PRODUCED, out_base);
        rdf_st_release(s.header + RDF_H_DONE, 1);
    }
}

// ---- executors -----------------------------------------------------------------------------
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

__device__ __forceinline__ RdfTask rdf_load_task(const RdfTask* p) {  // L2: the scheduler is still writing the list
    const int4 x = __ldcg((const int4*)p), y = __ldcg((const int4*)p + 1);
    RdfTask t;
    t.user = x.x; t.item = x.y; t.neg = x.z; t.eu = x.w;
    t.ei = y.x; t.en = y.y; t.weight = __int_as_float(y.z); t.y = __int_as_float(y.w);
    return t;
}

template <int LOSS, int K, int AD>
__device__ __forceinline__ void rdf_execute(const FitArgs& a, const RdfScratch& s, float* smem, int e_id, int n_exec) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const DevModel& m = a.model;
    const int d = m.d;
    float* su = smem + (size_t)wib * 3 * (d + 1);
    float* sp = su + (d + 1);
    float* sn = sp + (d + 1);
    const float fw = (float)((double)1.0f * 1.0);  // f32(double(1.0f) * scale), scale == 1
    const double lr = (double)m.lr;
    int produced = 0;  // tasks known to exist

    for (int t = e_id;; t += n_exec) {
        // ---- the task exists? ----
        if (t >= produced) {
            int go = 1;
            if (lane == 0) {
                for (;;) {
                    produced = rdf_ld_acquire(s.header + RDF_H_PRODUCED);
                    if (t < produced) break;
                    if (rdf_ld_acquire(s.header + RDF_H_DONE)) {
                        produced = rdf_ld_acquire(s.header + RDF_H_PRODUCED);
                        go = t < produced;
                        break;
                    }
                    // the scheduler emits a task every few tens of ns: sleep roughly until mine is due
                    const int ahead = t - produced;
                    __nanosleep(ahead > 200 ? 4000 : 100 + 20 * ahead);
                }
            }
            __syncwarp();  // lane 0's acquire is ordered before every lane's task load
            go = __shfl_sync(LFM_FULL, go, 0);
            produced = __shfl_sync(LFM_FULL, produced, 0);
            if (!go) return;
        }
        const RdfTask tk = rdf_load_task(s.tasks + t);
        const bool same = LOSS == LOSS_BPR && tk.neg < 0;  // the kept draw is the positive item itself
        const int nwait = (LOSS == LOSS_BPR && !same) ? 3 : 2;
        // ---- its rows are as the earlier tasks left them? ----
        if (lane < nwait) {
            const int32_t* p = lane == 0 ? s.ver_user + tk.user : (lane == 1 ? s.ver_item + tk.item : s.ver_item + tk.neg);
            const int e = lane == 0 ? tk.eu : (lane == 1 ? tk.ei : tk.en);
            // a version that never arrives would be a scheduling bug: give up loudly, not forever
            for (unsigned spins = 0; rdf_ld_relaxed(p) != e; spins++)
                if (spins > (1u << 23)) {
                    atomicExch(s.header + RDF_H_STALL, 1);
                    break;
                }
            __threadfence();  // acquire: the row loads below come after the version that was seen
        }
        __syncwarp();
        RdfRow<K, AD> U, P, N;
        rdf_load_row<K, AD>(U, m.user, tk.user, d, lane);
        rdf_load_row<K, AD>(P, m.item, tk.item, d, lane);
        if (LOSS == LOSS_BPR) {
            if (same) N = P;
            else rdf_load_row<K, AD>(N, m.item, tk.neg, d, lane);
        }
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
            // warp_update (T:537-649): biases, then per component positive / negative / user.  With
            // negative == positive the two steps hit the same element one after the other.
            if (lane == 0) {
                step(&P.b, &P.bg, &P.bm, (double)1.0f, -loss, AD, lr, 0.0, m.rho, m.eps);
                if (same) step(&P.b, &P.bg, &P.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
            }
            if (lane == 1 && !same) step(&N.b, &N.bg, &N.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
            if (lane == 2) step(&U.b, &U.bg, &U.bm, (double)1.0f, loss, AD, lr, 0.0, m.rho, m.eps);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int j = lane + 32 * k;
                if (j < d) {
                    const float uc = su[j], pc = sp[j], nc = sn[j];
                    step(&P.w[k], &P.g[k], &P.m[k], (double)1.0f, (-loss) * (double)uc, AD, lr, 0.0, m.rho, m.eps);
                    if (same) step(&P.w[k], &P.g[k], &P.m[k], (double)1.0f, loss * (double)uc, AD, lr, 0.0, m.rho, m.eps);
                    else step(&N.w[k], &N.g[k], &N.m[k], (double)1.0f, loss * (double)uc, AD, lr, 0.0, m.rho, m.eps);
                    step(&U.w[k], &U.g[k], &U.m[k], (double)1.0f, loss * (double)(float)(nc - pc), AD, lr, 0.0,
                         m.rho, m.eps);
                }
            }
            rdf_store_row<K, AD>(P, m.item, tk.item, d, lane, 0);
            if (!same) rdf_store_row<K, AD>(N, m.item, tk.neg, d, lane, 1);
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
        __syncwarp();  // every lane's row stores are ordered before the releases below (cumulative)
        if (lane < nwait) {
            int32_t* p = lane == 0 ? s.ver_user + tk.user : (lane == 1 ? s.ver_item + tk.item : s.ver_item + tk.neg);
            const int e = lane == 0 ? tk.eu : (lane == 1 ? tk.ei : tk.en);
            rdf_st_release(p, e + 1);
        }
    }
}

// CTA 0: its first three warps are the scheduler's stages (the rest of that SM stays idle so that
// the epoch's sequential part owns the SM's issue slots); every other CTA: RDF_WARPS executors.
template <int LOSS, int K, int AD>
__global__ void __launch_bounds__(RDF_WARPS * 32) rdf_kernel(FitArgs a, RdfScratch s) {
    extern __shared__ __align__(16) float rdf_smem[];
    if (blockIdx.x == 0) {
        // scheduler state in shared memory: [touch counters + marks | bitmap copy | pipe]
        int32_t* cnt_smem = (int32_t*)rdf_smem;
        size_t off = s.cnt_in_smem ? (size_t)(a.model.user.n + a.model.item.n + (LOSS == LOSS_BPR ? 2 * a.model.item.n : 0)) : 0;
        uint32_t* bitmap_smem = (uint32_t*)(cnt_smem + off);
        if (LOSS == LOSS_BPR && s.bitmap_in_smem) off += (size_t)a.pos.rows * s.bitmap_words;
        off = (off + 3) & ~(size_t)3;
        RdfPipe* pp = (RdfPipe*)(cnt_smem + off);
        if (threadIdx.x == 0) { pp->a_prod = 0; pp->a_cons = 0; pp->b_prod = 0; pp->b_cons = 0; }
        __syncthreads();
        const int w = threadIdx.x >> 5;
        if (w == 0) rdf_stage_sample<LOSS>(a, s, pp, bitmap_smem);
        else if (w == 1) rdf_stage_version<LOSS>(a, s, pp, cnt_smem);
        else if (w == 2) rdf_stage_emit(a, s, pp);
        return;
    }
    // consecutive tasks go to different SMs: the runnable ones are always the earliest ones
    const int n_exec = ((int)gridDim.x - 1) * RDF_WARPS;
    const int e_id = (int)(threadIdx.x >> 5) * ((int)gridDim.x - 1) + ((int)blockIdx.x - 1);
    rdf_execute<LOSS, K, AD>(a, s, rdf_smem, e_id, n_exec);
}

}  // namespace

// Scratch the dataflow path needs for (loss, a); 0 when (loss, a) is outside its scope.
static size_t rdf_scratch_bytes(int loss, const FitArgs& a, int64_t bitmap_limit_bytes) {
    const DevModel& m = a.model;
    if ((loss != LOSS_BPR && loss != LOSS_LOGISTIC) || !a.itf.identity || !a.usf.identity || a.item_alpha != 0.0 ||
        a.user_alpha != 0.0 || m.d > 256 || m.d < 1 || a.n > 0x7ff00000LL || a.n < 1)  // task indices are int32, with headroom for the executors' stride
        return 0;
    if (loss == LOSS_BPR && !a.pos.indptr) return 0;
    size_t b = 256 + sizeof(int32_t) * (2 * ((size_t)m.user.n + (size_t)m.item.n + 64) + 2 * (size_t)m.item.n) +
               (sizeof(RdfTask) + sizeof(Tuple)) * (size_t)a.n + 1024;
    if (loss == LOSS_BPR) {
        const size_t words = ((size_t)a.pos.cols + 31) / 32;
        const size_t bm = sizeof(uint32_t) * words * (size_t)a.pos.rows;
        if ((int64_t)bm <= bitmap_limit_bytes) b += bm + 256;
    }
    return b;
}

static cudaEvent_t g_rdf_ev[2] = {nullptr, nullptr};
static double g_rdf_ms[2] = {0.0, 0.0};  // scheduler (its emit stage's lifetime), whole kernel of the last dataflow epoch
static int g_rdf_tasks = -1;

// Returns cudaErrorNotSupported when the epoch has to run in replay_kernel instead (out of scope);
// nothing has been modified then.
static cudaError_t lfm_try_launch_replay_dataflow(int loss, const FitArgs& a, cudaStream_t st) {
    if (!a.replay_scratch || a.replay_scratch_bytes == 0) return cudaErrorNotSupported;
    for (int i = 0; i < 2; i++)
        if (!g_rdf_ev[i]) {
            cudaError_t ee = cudaEventCreate(&g_rdf_ev[i]);
            if (ee != cudaSuccess) return ee;
        }
    const DevModel& m = a.model;
    const int d = m.d;
    const size_t nu = (size_t)m.user.n, ni = (size_t)m.item.n;
    int dev = 0, sms = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < 2) return cudaErrorNotSupported;
    unsigned char* base = (unsigned char*)a.replay_scratch;
    RdfScratch s;
    s.header = (int32_t*)base;
    s.cnt_user = (int32_t*)(base + 256);
    s.cnt_item = s.cnt_user + nu;
    s.ver_user = s.cnt_item + ni;
    s.ver_item = s.ver_user + nu;
    s.mark = s.ver_item + ni;
    const size_t zero_bytes = 256 + sizeof(int32_t) * (2 * (nu + ni + 64) + 2 * ni);
    size_t off = zero_bytes;
    off = (off + 255) & ~(size_t)255;
    s.tasks = (RdfTask*)(base + off);
    off += sizeof(RdfTask) * (size_t)a.n;
    off = (off + 255) & ~(size_t)255;
    Tuple* tuples = (Tuple*)(base + off);
    s.tuples = tuples;
    off += sizeof(Tuple) * (size_t)a.n;
    off = (off + 255) & ~(size_t)255;
    s.bitmap = nullptr;
    s.bitmap_words = 0;
    s.mod_magic = ~0ull / (uint64_t)a.n + 1ull;
    // scheduler state in shared memory: touch counters [nu + ni] and, for BPR, the two mark arrays [2 ni]
    const size_t cnt_bytes = sizeof(int32_t) * (nu + ni + (loss == LOSS_BPR ? 2 * ni : 0));
    s.cnt_in_smem = cnt_bytes <= 160 * 1024 ? 1 : 0;
    cudaError_t e = cudaMemsetAsync(base, 0, zero_bytes, st);
    if (e != cudaSuccess) return e;
    e = lfm_launch_pack(a, loss, tuples, 0u, st);  // the host shuffle order (a.shuffle), Y <= 0 marked for BPR
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
    size_t smem = sizeof(float) * 3 * (d + 1) * RDF_WARPS;
    size_t sched_smem = s.cnt_in_smem ? cnt_bytes : 0;
    s.bitmap_in_smem = 0;
    const size_t pipe_bytes = sizeof(RdfPipe) + 32;
    if (s.bitmap) {
        const size_t bm = sizeof(uint32_t) * (size_t)s.bitmap_words * (size_t)a.pos.rows;
        if (sched_smem + bm + pipe_bytes <= 227 * 1024) {  // the per-CTA shared-memory limit of sm_100
            s.bitmap_in_smem = 1;
            sched_smem += bm;
        }
    }
    sched_smem += pipe_bytes;
    if (sched_smem > smem) smem = sched_smem;
    void* args[2] = {(void*)&a, (void*)&s};
    cudaEventRecord(g_rdf_ev[0], st);
#define RDF_LAUNCH(L, KK, AA)                                                                              \
    do {                                                                                                   \
        auto kern = rdf_kernel<L, KK, AA>;                                                                 \
        if (smem > 48 * 1024) {                                                                            \
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
            if (e != cudaSuccess) return e;                                                                \
        }                                                                                                  \
        int per_sm = 0;                                                                                    \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RDF_WARPS * 32, smem);            \
        if (e != cudaSuccess) return e;                                                                    \
        if (per_sm < 1) return cudaErrorNotSupported;                                                      \
        /* one CTA per SM: ~1200 warps, far more than the graph is wide; CTA 0 = the scheduler */          \
        int64_t blocks = sms;                                                                              \
        if (const char* ev = getenv("LFM_RDF_CTAS")) { const int v = atoi(ev); if (v >= 2 && v <= sms) blocks = v; } \
        const int64_t need = 1 + (a.n + RDF_WARPS - 1) / RDF_WARPS;                                        \
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
    cudaEventRecord(g_rdf_ev[1], st);
    int32_t hdr[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyAsync(hdr, s.header, sizeof(hdr), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (hdr[RDF_H_STALL] != 0) return cudaErrorLaunchTimeout;  // a task waited for a version that never came
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_rdf_ev[0], g_rdf_ev[1]);
    g_rdf_ms[0] = hdr[RDF_H_SCHED_US] / 1000.0;
    g_rdf_ms[1] = ms;
    g_rdf_tasks = hdr[RDF_H_TOTAL];
    if (getenv("LFM_RDF_PROFILE"))
        fprintf(stderr, "[rdf] tasks %d kernel %.3f ms scheduler %.3f ms | stage kcycles work/wait: sample %d/%d version %d/%d emit %d/%d | "
                        "rounds %d chunks %d | smem cnt %d bitmap %d\n", hdr[RDF_H_TOTAL], ms, hdr[RDF_H_SCHED_US] / 1000.0,
                hdr[8], hdr[9], hdr[10], hdr[11], hdr[12], hdr[13], hdr[14], hdr[15], s.cnt_in_smem, s.bitmap_in_smem);
    return cudaSuccess;
}
