// lfm_predict.cu -- read-only scoring kernels: predict_lightfm, predict_ranks,
// calculate_auc_from_rank, in_positives test hook.
//
// Compiled with --fmad=false: scores follow the reference's arithmetic exactly
// (fp32 multiply then fp32 add, features in CSR order, components left to
// right; T:287-334), so `predict` is bit-identical to the reference's IEEE
// build and `predict_rank` never disagrees with it on a near-tie.
//
// Reference: predict_lightfm T:1185-1229, predict_ranks T:1232-1323,
// calculate_auc_from_rank T:1326-1376, __test_in_positives T:1380-1385.
#include <atomic>

#include "lfm_common.cuh"

namespace {

// Lanes own components (l, l+32, ...), features visited sequentially in CSR order.
// Result goes to `repr` (shared or global, stride `rs` between components).
__device__ __forceinline__ void gather_to(const DevCsr& f, const float* __restrict__ emb,
                                          const float* __restrict__ bias, int d, int row,
                                          float* repr, size_t rs, int lane) {
    int start = f.indptr[row], stop = f.indptr[row + 1];
    for (int j = lane; j <= d; j += 32) {
        float acc = 0.0f;
        for (int i = start; i < stop; i++) {
            int ft = f.indices[i];
            float fw = f.data[i];  // scale == 1.0 outside training: f32(double(w) * 1.0) == w
            float v = (j < d) ? emb[(size_t)ft * d + j] : bias[ft];
            acc = acc + fw * v;
        }
        repr[(size_t)j * rs] = acc;
    }
}

__global__ void predict_kernel(DevCsr itf, DevCsr usf, DevModel m, const int32_t* user_ids,
                               const int32_t* item_ids, float* out, int64_t n) {
    extern __shared__ float sm[];
    int d = m.d;
    int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* u = sm + (size_t)wib * 2 * (d + 1);
    float* v = u + (d + 1);
    int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = w; i < n; i += nw) {
        __syncwarp();
        gather_to(usf, m.user.w, m.user.b, d, user_ids[i], u, 1, lane);
        gather_to(itf, m.item.w, m.item.b, d, item_ids[i], v, 1, lane);
        __syncwarp();
        if (lane == 0) {
            float r = u[d] + v[d];
            for (int j = 0; j < d; j++) r = r + u[j] * v[j];
            out[i] = r;
        }
    }
}

// Item representations, transposed: repr_t[j * ld + item], j in [0, d]; ld = n_items rounded up
// to a multiple of 4 so that every row starts 16 B aligned (the rank kernel reads float4s).
__global__ void item_repr_kernel(DevCsr itf, DevModel m, float* repr_t, int n_items, int ld) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = w; it < n_items; it += nw)
        gather_to(itf, m.item.w, m.item.b, m.d, (int)it, repr_t + it, (size_t)ld, lane);
}

// ---- predict_ranks (T:1232-1323), tiled over users ---------------------------------------
// rank[t] = #{ items i : i not a train positive of the user, i != t, score(u,i) >= score(u,t) }.
// A CTA takes UT users with test interactions at a time, so every element of the (transposed)
// item table it streams is used for UT scores instead of one: scores are accumulated per
// (user, item) in the reference's order (fp32 multiply then add, components left to right), and
// are compared with the users' test scores as soon as they exist -- nothing is written back.
// Train positives are handled by subtraction: count over ALL items, then take away the train
// positives' contribution (the same comparison on the same recomputed scores), which equals
// skipping them (T:1303-1304).  Comparisons with NaN are false in both passes, as in the reference.
#define RANK_UT 8    // users per tile (= warps per CTA)
#define RANK_TCH 64  // test interactions per user handled per pass over the catalogue
#define RANK_IT 4    // consecutive items per thread (one float4 of the transposed item table per component)

__global__ void compact_users_kernel(DevCsr test, int32_t* list, int32_t* count) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < test.rows && test.indptr[u + 1] > test.indptr[u]) list[atomicAdd(count, 1)] = u;
}

// score of (user tile slot, item): the reference's order -- biases first, then components left
// to right, multiply and add rounded separately (file is built with --fmad=false).
// `ut` is the user tile transposed: ut[j * RANK_UT + slot], j in [0, d].
__device__ __forceinline__ float rank_score(const float* __restrict__ ut, int slot,
                                            const float* __restrict__ irt, int ld, int d, int item) {
    float r = ut[d * RANK_UT + slot] + irt[(size_t)d * ld + item];
    for (int j = 0; j < d; j++) r = r + ut[j * RANK_UT + slot] * irt[(size_t)j * ld + item];
    return r;
}

// Scores of RANK_UT users x RANK_IT consecutive items per thread, streamed over the transposed item
// table.  Per component: one float4 of item values (coalesced), two broadcast float4 loads of the
// eight users' values, 32 multiplies + 32 adds: the FP32 pipe is the limiter, not the LSU (the first
// version read the eight user values with eight scalar shared loads per 16 flops).
template <int UNR = 4>
__device__ __forceinline__ void tile_scores(const float* __restrict__ ut, const float* __restrict__ irt,
                                            int ld, int d, int i, float (&acc)[RANK_IT][RANK_UT]) {
    const float4 vb = *(const float4*)(irt + (size_t)d * ld + i);
    const float4 ub0 = *(const float4*)(ut + d * RANK_UT), ub1 = *(const float4*)(ut + d * RANK_UT + 4);
    const float ub[8] = {ub0.x, ub0.y, ub0.z, ub0.w, ub1.x, ub1.y, ub1.z, ub1.w};
    const float vbb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int k = 0; k < RANK_IT; k++)
#pragma unroll
        for (int u = 0; u < RANK_UT; u++) acc[k][u] = ub[u] + vbb[k];
#pragma unroll UNR
    for (int j = 0; j < d; j++) {
        const float4 v4 = *(const float4*)(irt + (size_t)j * ld + i);
        const float4 u0 = *(const float4*)(ut + j * RANK_UT), u1 = *(const float4*)(ut + j * RANK_UT + 4);
        const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < RANK_IT; k++)
#pragma unroll
            for (int u = 0; u < RANK_UT; u++) acc[k][u] = acc[k][u] + uu[u] * vv[k];
    }
}

// Build the transposed tile of user representations: warp w gathers user w's vector.
__device__ __forceinline__ void build_user_tile(const DevCsr& usf, const DevModel& m, const int* s_user,
                                                float* ut, int w, int lane) {
    const int d = m.d;
    if (s_user[w] >= 0) gather_to(usf, m.user.w, m.user.b, d, s_user[w], ut + w, RANK_UT, lane);
    else for (int j = lane; j <= d; j += 32) ut[j * RANK_UT + w] = 0.0f;
}

// RANK_G user tiles run side by side in one CTA (one group of RANK_UT warps each) and walk the
// item table in lockstep (a barrier per 1024-item step): the second and third group find the item
// values the first one pulled from L2 in the SM's L1, which cuts the L2 -> SM traffic per score by
// RANK_G (at 8 users per item value the stream was L2-bandwidth-bound: profiles/README.md).
#define RANK_GT (RANK_UT * 32)  // threads per group

template <int RANK_G>
__global__ void __launch_bounds__(RANK_G * RANK_GT, 1) predict_ranks_tiled_kernel(
    DevCsr usf, DevCsr test, DevCsr train, DevModel m, const float* __restrict__ irt, int ld,
    const int32_t* __restrict__ active, const int32_t* __restrict__ n_active_p, float* ranks) {
    extern __shared__ __align__(16) float sm_all[];
    const int d = m.d, n_items = test.cols;
    const int g = threadIdx.x / RANK_GT, tl = threadIdx.x % RANK_GT;  // group, thread within the group
    const int lane = tl & 31, w = tl >> 5;
    const int per_group = RANK_UT * (d + 1) + RANK_UT * RANK_TCH * 3 + RANK_UT * RANK_UT * RANK_TCH;  // 4-byte words
    float* sm = sm_all + (size_t)g * ((per_group + 3) & ~3);
    float* ut = sm;                                        // [d+1][UT] transposed user tile
    float* tp = ut + RANK_UT * (d + 1);                    // [UT][TCH] test scores
    int* tid = (int*)(tp + RANK_UT * RANK_TCH);            // [UT][TCH] test item ids
    int* cnt = tid + RANK_UT * RANK_TCH;                   // [UT warps][UT][TCH] partial counts
    int* sub = cnt + RANK_UT * RANK_UT * RANK_TCH;         // [UT][TCH] train-positive counts
    __shared__ int s_user[RANK_G][RANK_UT], s_ts[RANK_G][RANK_UT], s_T[RANK_G][RANK_UT], s_maxT[RANK_G];
    const int n_active = *n_active_p;

    for (int tile0 = blockIdx.x * RANK_G; tile0 * RANK_UT < n_active; tile0 += gridDim.x * RANK_G) {
        const int tile = tile0 + g;
        __syncthreads();
        if (tl < RANK_UT) {
            int k = tile * RANK_UT + tl;
            int u = k < n_active ? active[k] : -1;
            s_user[g][tl] = u;
            s_ts[g][tl] = u >= 0 ? test.indptr[u] : 0;
            s_T[g][tl] = u >= 0 ? test.indptr[u + 1] - test.indptr[u] : 0;
        }
        __syncthreads();
        build_user_tile(usf, m, s_user[g], ut, w, lane);
        if (tl == 0) {
            int mt = 0;
            for (int u = 0; u < RANK_UT; u++) mt = max(mt, s_T[g][u]);
            s_maxT[g] = mt;
        }
        __syncthreads();
        int maxT = 0;  // block-uniform: every group walks the same number of test chunks
        for (int gg = 0; gg < RANK_G; gg++) maxT = max(maxT, s_maxT[gg]);

        for (int c0 = 0; c0 < maxT; c0 += RANK_TCH) {
            // B0: this chunk's test scores (T:1283-1298) and cleared counters
            for (int x = tl; x < RANK_UT * RANK_TCH; x += RANK_GT) {
                int u = x / RANK_TCH, t = x % RANK_TCH;
                if (c0 + t < s_T[g][u]) {
                    int id = test.indices[s_ts[g][u] + c0 + t];
                    tid[x] = id;
                    tp[x] = rank_score(ut, u, irt, ld, d, id);
                }
                sub[x] = 0;
            }
            for (int x = tl; x < RANK_UT * RANK_UT * RANK_TCH; x += RANK_GT) cnt[x] = 0;
            __syncthreads();

            // B + C: stream the catalogue once.  The test item itself is part of the stream: its
            // score there is the very same sequence of operations as tp[], so it counts itself
            // exactly once unless the score is NaN; that 1 is taken off in E instead of testing
            // `item != test item` on every comparison (T:1305-1306).
            for (int i0 = 0; i0 < n_items; i0 += RANK_GT * RANK_IT) {
                const int i = i0 + tl * RANK_IT;
                float acc[RANK_IT][RANK_UT];
                if (i < ld) tile_scores<(RANK_G == 2 ? 8 : 4)>(ut, irt, ld, d, i, acc);
                const float ninf = __int_as_float(0xff800000);
#pragma unroll
                for (int k = 0; k < RANK_IT; k++)
                    if (i + k >= n_items) {
#pragma unroll
                        for (int u = 0; u < RANK_UT; u++) acc[k][u] = ninf;  // past the catalogue: never >= anything
                    }
#pragma unroll
                for (int u = 0; u < RANK_UT; u++) {
                    const int Tc = min(RANK_TCH, s_T[g][u] - c0);
                    for (int t = 0; t < Tc; t++) {
                        const float ref = tp[u * RANK_TCH + t];
                        int c = 0;
#pragma unroll
                        for (int k = 0; k < RANK_IT; k++) c += (acc[k][u] >= ref) ? 1 : 0;
                        c = __reduce_add_sync(LFM_FULL, c);
                        if (lane == 0) cnt[(w * RANK_UT + u) * RANK_TCH + t] += c;
                    }
                }
                __syncthreads();  // keep the groups on the same 1024-item step (shared L1 lines)
            }
            // D: the train positives' share of those counts (warp w <-> user w)
            {
                const int u = s_user[g][w];
                const int Tc = min(RANK_TCH, s_T[g][w] - c0);
                if (u >= 0 && Tc > 0 && u < train.rows) {
                    const int rs = train.indptr[u], re = train.indptr[u + 1];
                    for (int e0 = rs; e0 < re; e0 += 32) {
                        const int e = e0 + lane;
                        bool in = e < re;
                        int item = in ? train.indices[e] : -1;
                        if (in && e > rs && train.indices[e - 1] == item) in = false;  // duplicate entry
                        if (in && (item < 0 || item >= n_items)) in = false;
                        const float sc = in ? rank_score(ut, w, irt, ld, d, item) : 0.0f;
                        for (int t = 0; t < Tc; t++) {
                            const bool hit = in && item != tid[w * RANK_TCH + t] && sc >= tp[w * RANK_TCH + t];
                            const unsigned bal = __ballot_sync(LFM_FULL, hit);
                            if (lane == 0) sub[w * RANK_TCH + t] += __popc(bal);
                        }
                    }
                }
            }
            __syncthreads();
            // E: reduce over warps and accumulate into ranks (pre-zeroed by the caller, L:968-975)
            for (int x = tl; x < RANK_UT * RANK_TCH; x += RANK_GT) {
                int u = x / RANK_TCH, t = x % RANK_TCH;
                if (c0 + t < s_T[g][u]) {
                    int total = 0;
                    for (int ww = 0; ww < RANK_UT; ww++) total += cnt[(ww * RANK_UT + u) * RANK_TCH + t];
                    const float ref = tp[x];
                    if (ref == ref) total -= 1;  // the test item's own (non-NaN) score in the stream
                    ranks[s_ts[g][u] + c0 + t] += (float)(total - sub[x]);
                }
            }
            __syncthreads();
        }
    }
}

// T:1336-1376 after the per-row ascending sort: sequential fp32 accumulation per user.
__global__ void auc_kernel(DevCsr ranks, const int32_t* num_train_pos, const float* rank_data,
                           float* auc) {
    int user = blockIdx.x * blockDim.x + threadIdx.x;
    if (user >= ranks.rows) return;
    int rs = ranks.indptr[user], re = ranks.indptr[user + 1];
    int num_pos = re - rs;
    int num_neg = ranks.cols - ((re - rs) + num_train_pos[user]);
    if (num_pos == 0 || num_neg == ranks.cols) { auc[user] = 0.5f; return; }
    float acc = auc[user];
    for (int i = 0; i < num_pos; i++) {
        float rank = rank_data[rs + i];
        rank = rank - (float)i;
        if (rank < 0) rank = 0;
        acc = (float)((double)acc + (1.0 - (double)(rank / (float)num_neg)));
    }
    if (num_pos != 0) acc = acc / (float)num_pos;
    auc[user] = acc;
}

// ---- per-row ascending sort of the rank values (T:1352 qsort(flt_compare)), in place ------------
// Rows of up to 32 values: one warp per row, each lane finds its value's position by counting
// (stable: equal values keep their order).  Longer rows: one CTA per row, bitonic network in shared
// memory up to SORT_SMEM values, position counting from global memory beyond that.
#define SORT_SMEM 4096
__global__ void row_sort_small_kernel(DevCsr r, float* data) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= r.rows) return;
    const int rs = r.indptr[row], n = r.indptr[row + 1] - rs;
    if (n < 2 || n > 32) return;
    const float x = lane < n ? data[rs + lane] : 0.0f;
    int pos = 0;
    for (int j = 0; j < n; j++) {
        const float y = __shfl_sync(LFM_FULL, x, j);
        pos += (y < x || (y == x && j < lane)) ? 1 : 0;
    }
    __syncwarp();
    if (lane < n) data[rs + pos] = x;
}

__global__ void __launch_bounds__(256) row_sort_large_kernel(DevCsr r, float* data, float* tmp) {
    __shared__ float sh[SORT_SMEM];
    for (int row = blockIdx.x; row < r.rows; row += gridDim.x) {
        const int rs = r.indptr[row], n = r.indptr[row + 1] - rs;
        if (n <= 32) continue;
        __syncthreads();
        if (n <= SORT_SMEM) {
            int p2 = 64;
            while (p2 < n) p2 <<= 1;
            for (int i = threadIdx.x; i < p2; i += blockDim.x) sh[i] = i < n ? data[rs + i] : __int_as_float(0x7f800000);
            __syncthreads();
            for (int k = 2; k <= p2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                        const int l = i ^ j;
                        if (l > i) {
                            const float a = sh[i], b = sh[l];
                            const bool up = (i & k) == 0;
                            if ((a > b) == up) { sh[i] = b; sh[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (int i = threadIdx.x; i < n; i += blockDim.x) data[rs + i] = sh[i];
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const float x = data[rs + i];
                int pos = 0;
                for (int j = 0; j < n; j++) {
                    const float y = data[rs + j];
                    pos += (y < x || (y == x && j < i)) ? 1 : 0;
                }
                tmp[rs + pos] = x;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) data[rs + i] = tmp[rs + i];
        }
    }
}

// ---- fused evaluation epilogue (lightfm/evaluation.py:14-327 on device-resident ranks) ----------
// Per user: hits = #{rank < k} (precision@k = hits / k, recall@k = hits / #test), best = smallest
// rank (reciprocal rank = 1 / (best + 1)); AUC is the reference's calculate_auc_from_rank on the
// sorted row.  Only these per-user values cross PCIe, never the nnz_test ranks.
__global__ void rank_metrics_kernel(DevCsr test, const float* __restrict__ ranks, int k, int32_t* hits,
                                    float* best) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= test.rows) return;
    const int rs = test.indptr[row], re = test.indptr[row + 1];
    int h = 0;
    float b = __int_as_float(0x7f800000);
    for (int e = rs + lane; e < re; e += 32) {
        const float r = ranks[e];
        h += (r < (float)k) ? 1 : 0;
        b = fminf(b, r);
    }
    h = __reduce_add_sync(LFM_FULL, h);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(LFM_FULL, b, o));
    if (lane == 0) {
        if (hits) hits[row] = h;
        if (best) best[row] = (re > rs) ? b : -1.0f;
    }
}

__global__ void row_counts_kernel(DevCsr m, int32_t* out, int rows) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) out[r] = r < m.rows ? m.indptr[r + 1] - m.indptr[r] : 0;
}

// ---- top-k recommendation (doc/quickstart.rst:125-126 does np.argsort(-model.predict(u, all items))) ---
// Phase 1 reuses the rank kernel's tiling to write the score rows of a batch of users; phase 2
// selects per user: 4-pass 8-bit radix select of the k-th largest key, collection of the winners
// (ties at the threshold by ascending item id), bitonic sort of the k winners.
__device__ __forceinline__ uint32_t score_key(float s) {  // larger score -> larger key; NaN lowest
    if (s != s) return 0u;
    if (s == 0.0f) s = 0.0f;  // -0 and +0 compare equal: give them one key
    const uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(RANK_UT * 32) score_rows_kernel(
    DevCsr usf, DevModel m, const float* __restrict__ irt, int ld, int n_items,
    const int32_t* __restrict__ user_ids, int n_users, float* __restrict__ scores) {
    extern __shared__ __align__(16) float sm[];
    float* ut = sm;
    __shared__ int s_user[RANK_UT];
    const int d = m.d, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile * RANK_UT < n_users; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < RANK_UT) {
            const int k = tile * RANK_UT + threadIdx.x;
            s_user[threadIdx.x] = k < n_users ? user_ids[k] : -1;
        }
        __syncthreads();
        build_user_tile(usf, m, s_user, ut, w, lane);
        __syncthreads();
        for (int i0 = 0; i0 < ld; i0 += blockDim.x * RANK_IT) {
            const int i = i0 + threadIdx.x * RANK_IT;
            if (i >= ld) continue;
            float acc[RANK_IT][RANK_UT];
            tile_scores(ut, irt, ld, d, i, acc);
#pragma unroll
            for (int u = 0; u < RANK_UT; u++)
                if (tile * RANK_UT + u < n_users)
                    *(float4*)(scores + (size_t)(tile * RANK_UT + u) * ld + i) =
                        make_float4(acc[0][u], acc[1][u], acc[2][u], acc[3][u]);
        }
    }
}

// Mark the excluded (train) items of each user in its score row: NaN sorts last and is never returned.
__global__ void exclude_kernel(DevCsr train, const int32_t* __restrict__ user_ids, int n_users, int n_items,
                               int ld, float* scores) {
    const int lane = threadIdx.x & 31;
    const int k = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (k >= n_users) return;
    const int u = user_ids[k];
    if (u < 0 || u >= train.rows) return;
    for (int e = train.indptr[u] + lane; e < train.indptr[u + 1]; e += 32) {
        const int it = train.indices[e];
        if (it >= 0 && it < n_items) scores[(size_t)k * ld + it] = __int_as_float(0x7fc00000);
    }
}

#define TOPK_MAX 1024
__global__ void __launch_bounds__(256) topk_select_kernel(const float* __restrict__ scores, int ld, int n_items,
                                                          int n_users, int k, int32_t* out_items, float* out_scores) {
    __shared__ unsigned hist[256];
    __shared__ uint32_t s_prefix, s_need, s_count, s_ties;
    __shared__ uint32_t win_key[TOPK_MAX];
    __shared__ int32_t win_idx[TOPK_MAX];
    __shared__ uint32_t warp_cnt[8];
    for (int row = blockIdx.x; row < n_users; row += gridDim.x) {
        const float* sc = scores + (size_t)row * ld;
        __syncthreads();
        if (threadIdx.x == 0) { s_prefix = 0u; s_need = (uint32_t)k; }
        // ---- radix select: after the pass over byte `shift`, s_prefix holds the high bytes of the
        //      k-th largest key and s_need how many keys equal to the prefix-so-far are still wanted
        for (int shift = 24; shift >= 0; shift -= 8) {
            hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
                const uint32_t key = score_key(sc[i]);
                if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t need = s_need;
                int b = 255;
                for (; b > 0; b--) {
                    if (hist[b] >= need) break;
                    need -= hist[b];
                }
                s_prefix = prefix | ((uint32_t)b << shift);
                s_need = need;
            }
            __syncthreads();
        }
        const uint32_t thr = s_prefix;      // key of the k-th largest score
        const uint32_t need_ties = s_need;  // how many keys == thr belong to the top k
        if (threadIdx.x == 0) { s_count = 0u; s_ties = 0u; }
        // every slot starts as "no item" (fewer than k items can be scorable: n_items < k, NaN
        // scores, excluded items), and sorts behind every real winner
        for (int i = threadIdx.x; i < TOPK_MAX; i += blockDim.x) { win_key[i] = 0u; win_idx[i] = 0x7fffffff; }
        __syncthreads();
        // ---- winners strictly above the threshold (any order), then the ties in ascending item order
        for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
            const uint32_t key = score_key(sc[i]);
            if (key > thr) {
                const uint32_t p = atomicAdd(&s_count, 1u);
                win_key[p] = key;
                win_idx[p] = i;
            }
        }
        __syncthreads();
        const uint32_t n_above = s_count;
        for (int i0 = 0; i0 < n_items && s_ties < need_ties; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            const bool tie = i < n_items && score_key(sc[i]) == thr;
            // ordered position of this thread's tie inside the block: ballots + per-warp prefix
            const unsigned bal = __ballot_sync(LFM_FULL, tie);
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            if (lane == 0) warp_cnt[w] = __popc(bal);
            __syncthreads();
            uint32_t before = s_ties;
            for (int ww = 0; ww < w; ww++) before += warp_cnt[ww];
            const uint32_t mine = before + __popc(bal & ((1u << lane) - 1u));
            if (tie && mine < need_ties && n_above + mine < (uint32_t)k && thr != 0u) {
                win_key[n_above + mine] = thr;
                win_idx[n_above + mine] = i;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t tot = 0;
                for (int ww = 0; ww < 8; ww++) tot += warp_cnt[ww];
                s_ties += tot;
            }
            __syncthreads();
        }
        // ---- sort the k winners: key descending, item id ascending
        int p2 = 2;
        while (p2 < k) p2 <<= 1;
        __syncthreads();
        for (int kk = 2; kk <= p2; kk <<= 1)
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                    const int l = i ^ j;
                    if (l > i) {
                        const uint32_t ka = win_key[i], kb = win_key[l];
                        const int32_t ia = win_idx[i], ib = win_idx[l];
                        const bool a_first = ka > kb || (ka == kb && ia < ib);  // a belongs before b
                        const bool up = (i & kk) == 0;
                        if (a_first != up) { win_key[i] = kb; win_key[l] = ka; win_idx[i] = ib; win_idx[l] = ia; }
                    }
                }
                __syncthreads();
            }
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const int32_t it = win_idx[i];
            const bool ok = it != 0x7fffffff && win_key[i] != 0u;   // fewer than k scorable items: pad with -1
            out_items[(size_t)row * k + i] = ok ? it : -1;
            out_scores[(size_t)row * k + i] = ok ? sc[it] : __int_as_float(0x7fc00000);
        }
    }
}

// Small k (<= TOPK_SMALL): ONE pass over the row.  Every thread keeps the TOPK_SMALL best
// (key, item) pairs of its strided share in registers (an insertion happens only when an element
// beats the thread's current worst: ~k ln(n / 256 k) times per thread), then the CTA extracts the
// k global winners one by one: block-wide arg-max over the threads' current best, the winner pops
// it.  Order and ties as in the radix path: key descending, item id ascending.
#define TOPK_SMALL 16
__global__ void __launch_bounds__(256) topk_small_kernel(const float* __restrict__ scores, int ld, int n_items,
                                                         int n_users, int k, int32_t* out_items, float* out_scores) {
    __shared__ unsigned long long warp_best[8];
    __shared__ unsigned long long s_best;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int row = blockIdx.x; row < n_users; row += gridDim.x) {
        const float* sc = scores + (size_t)row * ld;
        uint32_t key[TOPK_SMALL];
        int32_t idx[TOPK_SMALL];
#pragma unroll
        for (int j = 0; j < TOPK_SMALL; j++) { key[j] = 0u; idx[j] = 0x7fffffff; }
        auto offer = [&](uint32_t kx, int32_t i) {
            if (kx > key[TOPK_SMALL - 1]) {  // ids arrive in ascending order: an equal key never displaces
                uint32_t ck = kx;
                int32_t ci = i;
#pragma unroll
                for (int j = 0; j < TOPK_SMALL; j++) {  // insert, keeping (key desc, id asc)
                    if (ck > key[j]) {
                        const uint32_t tk = key[j]; const int32_t ti = idx[j];
                        key[j] = ck; idx[j] = ci;
                        ck = tk; ci = ti;
                    }
                }
            }
        };
        // four consecutive items per thread per step (rows are padded to a multiple of 4 and 16 B-aligned),
        // two steps in flight
        const float4* sc4 = (const float4*)sc;
        const int n4 = ld >> 2;
        for (int q0 = threadIdx.x; q0 < n4; q0 += 2 * blockDim.x) {
            const int q1 = q0 + blockDim.x;
            const float4 a4 = __ldcs(sc4 + q0);
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q1 < n4) b4 = __ldcs(sc4 + q1);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (4 * q0 + e < n_items) offer(score_key(av[e]), 4 * q0 + e);
            if (q1 < n4) {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (4 * q1 + e < n_items) offer(score_key(bv[e]), 4 * q1 + e);
            }
        }
        for (int r = 0; r < k; r++) {
            // (key, smaller id first) as one 64-bit value; 0 = nothing left / not scorable
            unsigned long long v = key[0] ? (((unsigned long long)key[0] << 32) | (uint32_t)(0x7fffffff - idx[0])) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ov = __shfl_xor_sync(LFM_FULL, v, o);
                v = ov > v ? ov : v;
            }
            if (lane == 0) warp_best[w] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long b = warp_best[0];
                for (int ww = 1; ww < 8; ww++) b = warp_best[ww] > b ? warp_best[ww] : b;
                s_best = b;
            }
            __syncthreads();
            const unsigned long long best = s_best;
            const int32_t bi = best ? (int32_t)(0x7fffffff - (uint32_t)(best & 0xffffffffull)) : -1;
            if (best && key[0] == (uint32_t)(best >> 32) && idx[0] == bi) {  // mine: pop it
#pragma unroll
                for (int j = 0; j + 1 < TOPK_SMALL; j++) { key[j] = key[j + 1]; idx[j] = idx[j + 1]; }
                key[TOPK_SMALL - 1] = 0u; idx[TOPK_SMALL - 1] = 0x7fffffff;
            }
            if (threadIdx.x == 0) {  // fewer than k scorable items: pad with -1 / NaN
                out_items[(size_t)row * k + r] = bi;
                out_scores[(size_t)row * k + r] = bi >= 0 ? sc[bi] : __int_as_float(0x7fc00000);
            }
            __syncthreads();  // s_best / warp_best are rewritten in the next round
        }
    }
}

__global__ void in_positives_kernel(DevCsr mat, int row, int col, int32_t* out) {
    int lane = threadIdx.x;
    bool a = lfm_warp_member(mat.indices, mat.indptr[row], mat.indptr[row + 1], col, lane);
    bool b = lfm_bsearch(mat.indices, mat.indptr[row], mat.indptr[row + 1], col);
    if (lane == 0) *out = (a ? 1 : 0) | (b ? 2 : 0);  // both searches must agree (host checks)
}

__global__ void check_identity_kernel(DevCsr m, int32_t* flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < m.rows; i += stride)
        if (m.indptr[i] != (int)i || m.indices[i] != (int)i || m.data[i] != 1.0f) bad = true;
    if (bad) *flag = 0;
}

}  // namespace

cudaError_t lfm_launch_predict(const DevCsr& itf, const DevCsr& usf, const DevModel& m,
                               const int32_t* user_ids, const int32_t* item_ids, float* out,
                               int64_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    int warps = 8;
    size_t smem = (size_t)warps * 2 * (m.d + 1) * sizeof(float);
    int64_t blocks = (n + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    predict_kernel<<<(int)blocks, warps * 32, smem, st>>>(itf, usf, m, user_ids, item_ids, out, n);
    return cudaGetLastError();
}


// 1: one user tile per CTA (3 CTAs per SM); 3: three tiles per CTA walking the item table in lockstep
static std::atomic<int> g_rank_groups{1};
extern "C" int lfm_set_rank_groups(int groups) {
    int old = g_rank_groups.load();
    if (groups >= 1 && groups <= 3) g_rank_groups.store(groups);
    return old;
}

cudaError_t lfm_launch_predict_ranks(const DevCsr& itf, const DevCsr& usf, const DevCsr& test,
                                     const DevCsr& train, const DevModel& m, float* ranks,
                                     float* scratch, cudaStream_t st, int* launches) {
    // scratch layout: [ (d+1) * ld transposed item table | test.rows + 1 ints: active users, count ]
    int n_items = test.cols;
    if (test.rows == 0 || test.nnz == 0 || n_items == 0) return cudaSuccess;
    const int ld = (n_items + 3) & ~3;
    float* repr_t = scratch;
    int32_t* active = (int32_t*)(scratch + (size_t)ld * (m.d + 1));
    int32_t* count = active + test.rows;
    int64_t blocks = ((int64_t)n_items * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    item_repr_kernel<<<(int)blocks, 256, 0, st>>>(itf, m, repr_t, n_items, ld);
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    compact_users_kernel<<<(test.rows + 255) / 256, 256, 0, st>>>(test, active, count);
    const size_t per_group = (size_t)RANK_UT * (m.d + 1) + RANK_UT * RANK_TCH * 3 + RANK_UT * RANK_UT * RANK_TCH;
    const size_t group_bytes = sizeof(float) * ((per_group + 3) & ~(size_t)3);
    if (g_rank_groups.load() == 2) {
        const size_t smem = 2 * group_bytes;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(predict_ranks_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int grid = (test.rows + RANK_UT * 2 - 1) / (RANK_UT * 2);
        if (grid > 148) grid = 148;  // one persistent CTA (two user tiles, 128 registers per thread) per SM
        predict_ranks_tiled_kernel<2><<<grid, 2 * RANK_GT, smem, st>>>(usf, test, train, m, repr_t, ld, active, count, ranks);
    } else if (g_rank_groups.load() == 3) {
        const size_t smem = 3 * group_bytes;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(predict_ranks_tiled_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int grid = (test.rows + RANK_UT * 3 - 1) / (RANK_UT * 3);
        if (grid > 148) grid = 148;  // one persistent CTA (three user tiles) per SM
        predict_ranks_tiled_kernel<3><<<grid, 3 * RANK_GT, smem, st>>>(usf, test, train, m, repr_t, ld, active, count, ranks);
    } else {
        if (group_bytes > 48 * 1024)
            cudaFuncSetAttribute(predict_ranks_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_bytes);
        int grid = (test.rows + RANK_UT - 1) / RANK_UT;
        if (grid > 148 * 3) grid = 148 * 3;
        predict_ranks_tiled_kernel<1><<<grid, RANK_GT, group_bytes, st>>>(usf, test, train, m, repr_t, ld, active, count, ranks);
    }
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// exported for the host layer: number of floats predict_ranks needs in `scratch`
size_t lfm_ranks_scratch_floats(int n_items, int d, int test_rows) {
    return (size_t)((n_items + 3) & ~3) * (d + 1) + (size_t)test_rows + 4;
}

cudaError_t lfm_launch_row_sort(const DevCsr& rows, float* data, float* tmp, cudaStream_t st) {
    if (rows.rows == 0 || rows.nnz == 0) return cudaSuccess;
    row_sort_small_kernel<<<(int)(((int64_t)rows.rows * 32 + 255) / 256), 256, 0, st>>>(rows, data);
    int grid = rows.rows < 148 * 8 ? rows.rows : 148 * 8;
    row_sort_large_kernel<<<grid, 256, 0, st>>>(rows, data, tmp);
    return cudaGetLastError();
}

cudaError_t lfm_launch_auc(const DevCsr& ranks, const int32_t* num_train_pos, float* rank_data,
                           float* auc, float* tmp, cudaStream_t st) {
    if (ranks.rows == 0) return cudaSuccess;
    // per-row ascending sort, in place for the caller (T:1352)
    cudaError_t e = lfm_launch_row_sort(ranks, rank_data, tmp, st);
    if (e != cudaSuccess) return e;
    auc_kernel<<<(ranks.rows + 255) / 256, 256, 0, st>>>(ranks, num_train_pos, rank_data, auc);
    return cudaGetLastError();
}

cudaError_t lfm_launch_rank_metrics(const DevCsr& test, const float* ranks, int k, int32_t* hits, float* best,
                                    cudaStream_t st) {
    if (test.rows == 0) return cudaSuccess;
    rank_metrics_kernel<<<(int)(((int64_t)test.rows * 32 + 255) / 256), 256, 0, st>>>(test, ranks, k, hits, best);
    return cudaGetLastError();
}

cudaError_t lfm_launch_row_counts(const DevCsr& m, int32_t* out, int rows, cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    row_counts_kernel<<<(rows + 255) / 256, 256, 0, st>>>(m, out, rows);
    return cudaGetLastError();
}

// scratch: transposed item table [(d+1) * ld] followed by the score rows [n_users * ld]
cudaError_t lfm_launch_recommend(const DevCsr& itf, const DevCsr& usf, const DevCsr* exclude, const DevModel& m,
                                 int n_items, const int32_t* user_ids, int n_users, int k, int32_t* out_items,
                                 float* out_scores, float* scratch, cudaStream_t st, int* launches) {
    if (n_users == 0 || n_items == 0 || k == 0) return cudaSuccess;
    const int ld = (n_items + 3) & ~3;
    float* repr_t = scratch;
    float* scores = scratch + (size_t)ld * (m.d + 1);
    int64_t blocks = ((int64_t)n_items * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    item_repr_kernel<<<(int)blocks, 256, 0, st>>>(itf, m, repr_t, n_items, ld);
    // padding columns of the transposed table feed tile_scores: keep them finite
    const size_t smem = sizeof(float) * RANK_UT * (m.d + 1);
    int grid = (n_users + RANK_UT - 1) / RANK_UT;
    if (grid > 148 * 4) grid = 148 * 4;
    score_rows_kernel<<<grid, RANK_UT * 32, smem, st>>>(usf, m, repr_t, ld, n_items, user_ids, n_users, scores);
    if (exclude && exclude->nnz > 0)
        exclude_kernel<<<(int)(((int64_t)n_users * 32 + 255) / 256), 256, 0, st>>>(*exclude, user_ids, n_users, n_items, ld, scores);
    int g2 = n_users < 148 * 8 ? n_users : 148 * 8;
    if (k <= TOPK_SMALL) topk_small_kernel<<<g2, 256, 0, st>>>(scores, ld, n_items, n_users, k, out_items, out_scores);
    else topk_select_kernel<<<g2, 256, 0, st>>>(scores, ld, n_items, n_users, k, out_items, out_scores);
    if (launches) *launches += 4;
    return cudaGetLastError();
}

cudaError_t lfm_launch_in_positives(const DevCsr& mat, int32_t row, int32_t col, int32_t* out,
                                    cudaStream_t st) {
    in_positives_kernel<<<1, 32, 0, st>>>(mat, row, col, out);
    return cudaGetLastError();
}

cudaError_t lfm_launch_check_identity(const DevCsr& m, int32_t* flag, cudaStream_t st) {
    int blocks = (m.rows + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    check_identity_kernel<<<blocks, 256, 0, st>>>(m, flag);
    return cudaGetLastError();
}
