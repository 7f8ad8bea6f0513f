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
#include <cub/device/device_segmented_sort.cuh>

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

// Item representations, transposed: repr_t[j * I + item], j in [0, d].
__global__ void item_repr_kernel(DevCsr itf, DevModel m, float* repr_t, int n_items) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = w; it < n_items; it += nw)
        gather_to(itf, m.item.w, m.item.b, m.d, (int)it, repr_t + it, (size_t)n_items, lane);
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

__global__ void compact_users_kernel(DevCsr test, int32_t* list, int32_t* count) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < test.rows && test.indptr[u + 1] > test.indptr[u]) list[atomicAdd(count, 1)] = u;
}

__device__ __forceinline__ float rank_score(const float* __restrict__ u, const float* __restrict__ irt,
                                            int n_items, int d, int item) {
    float r = u[d] + irt[(size_t)d * n_items + item];
    for (int j = 0; j < d; j++) r = r + u[j] * irt[(size_t)j * n_items + item];
    return r;
}

__global__ void __launch_bounds__(RANK_UT * 32) predict_ranks_tiled_kernel(
    DevCsr usf, DevCsr test, DevCsr train, DevModel m, const float* __restrict__ irt,
    const int32_t* __restrict__ active, const int32_t* __restrict__ n_active_p, float* ranks) {
    extern __shared__ float sm[];
    const int d = m.d, n_items = test.cols;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* ur = sm;                                        // [UT][d+1]
    float* tp = ur + RANK_UT * (d + 1);                    // [UT][TCH] test scores
    int* tid = (int*)(tp + RANK_UT * RANK_TCH);            // [UT][TCH] test item ids
    int* cnt = tid + RANK_UT * RANK_TCH;                   // [UT warps][UT][TCH] partial counts
    int* sub = cnt + RANK_UT * RANK_UT * RANK_TCH;         // [UT][TCH] train-positive counts
    __shared__ int s_user[RANK_UT], s_ts[RANK_UT], s_T[RANK_UT];
    const int n_active = *n_active_p;

    for (int tile = blockIdx.x; tile * RANK_UT < n_active; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < RANK_UT) {
            int k = tile * RANK_UT + threadIdx.x;
            int u = k < n_active ? active[k] : -1;
            s_user[threadIdx.x] = u;
            s_ts[threadIdx.x] = u >= 0 ? test.indptr[u] : 0;
            s_T[threadIdx.x] = u >= 0 ? test.indptr[u + 1] - test.indptr[u] : 0;
        }
        __syncthreads();
        // A: warp w builds user w's representation (zeros for an empty slot)
        if (s_user[w] >= 0) gather_to(usf, m.user.w, m.user.b, d, s_user[w], ur + w * (d + 1), 1, lane);
        else for (int j = lane; j <= d; j += 32) ur[w * (d + 1) + j] = 0.0f;
        int maxT = 0;
        for (int u = 0; u < RANK_UT; u++) maxT = max(maxT, s_T[u]);
        __syncthreads();

        for (int c0 = 0; c0 < maxT; c0 += RANK_TCH) {
            // B0: this chunk's test scores (T:1283-1298) and cleared counters
            for (int x = threadIdx.x; x < RANK_UT * RANK_TCH; x += blockDim.x) {
                int u = x / RANK_TCH, t = x % RANK_TCH;
                if (c0 + t < s_T[u]) {
                    int id = test.indices[s_ts[u] + c0 + t];
                    tid[x] = id;
                    tp[x] = rank_score(ur + u * (d + 1), irt, n_items, d, id);
                }
                sub[x] = 0;
            }
            for (int x = threadIdx.x; x < RANK_UT * RANK_UT * RANK_TCH; x += blockDim.x) cnt[x] = 0;
            __syncthreads();

            // B + C: stream the catalogue once; UT scores per item element read
            for (int i0 = 0; i0 < n_items; i0 += blockDim.x) {
                const int i = i0 + threadIdx.x;
                const bool in = i < n_items;
                float acc[RANK_UT];
                {
                    const float vb = in ? irt[(size_t)d * n_items + i] : 0.0f;
#pragma unroll
                    for (int u = 0; u < RANK_UT; u++) acc[u] = ur[u * (d + 1) + d] + vb;
                }
                for (int j = 0; j < d; j++) {
                    const float v = in ? irt[(size_t)j * n_items + i] : 0.0f;
#pragma unroll
                    for (int u = 0; u < RANK_UT; u++) acc[u] = acc[u] + ur[u * (d + 1) + j] * v;
                }
#pragma unroll
                for (int u = 0; u < RANK_UT; u++) {
                    const int Tc = min(RANK_TCH, s_T[u] - c0);
                    for (int t = 0; t < Tc; t++) {
                        const bool hit = in && i != tid[u * RANK_TCH + t] && acc[u] >= tp[u * RANK_TCH + t];
                        const unsigned bal = __ballot_sync(LFM_FULL, hit);
                        if (lane == 0) cnt[(w * RANK_UT + u) * RANK_TCH + t] += __popc(bal);
                    }
                }
            }
            // D: the train positives' share of those counts (warp w <-> user w)
            {
                const int u = s_user[w];
                const int Tc = min(RANK_TCH, s_T[w] - c0);
                if (u >= 0 && Tc > 0 && u < train.rows) {
                    const int rs = train.indptr[u], re = train.indptr[u + 1];
                    for (int e0 = rs; e0 < re; e0 += 32) {
                        const int e = e0 + lane;
                        bool in = e < re;
                        int item = in ? train.indices[e] : -1;
                        if (in && e > rs && train.indices[e - 1] == item) in = false;  // duplicate entry
                        if (in && (item < 0 || item >= n_items)) in = false;
                        const float sc = in ? rank_score(ur + w * (d + 1), irt, n_items, d, item) : 0.0f;
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
            for (int x = threadIdx.x; x < RANK_UT * RANK_TCH; x += blockDim.x) {
                int u = x / RANK_TCH, t = x % RANK_TCH;
                if (c0 + t < s_T[u]) {
                    int total = 0;
                    for (int ww = 0; ww < RANK_UT; ww++) total += cnt[(ww * RANK_UT + u) * RANK_TCH + t];
                    ranks[s_ts[u] + c0 + t] += (float)(total - sub[x]);
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


cudaError_t lfm_launch_predict_ranks(const DevCsr& itf, const DevCsr& usf, const DevCsr& test,
                                     const DevCsr& train, const DevModel& m, float* ranks,
                                     float* scratch, cudaStream_t st, int* launches) {
    // scratch layout: [ (d+1) * n_items transposed item table | test.rows + 1 ints: active users, count ]
    int n_items = test.cols;
    if (test.rows == 0 || test.nnz == 0 || n_items == 0) return cudaSuccess;
    float* repr_t = scratch;
    int32_t* active = (int32_t*)(scratch + (size_t)n_items * (m.d + 1));
    int32_t* count = active + test.rows;
    int64_t blocks = ((int64_t)n_items * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    item_repr_kernel<<<(int)blocks, 256, 0, st>>>(itf, m, repr_t, n_items);
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    compact_users_kernel<<<(test.rows + 255) / 256, 256, 0, st>>>(test, active, count);
    size_t smem = sizeof(float) * (RANK_UT * (m.d + 1) + RANK_UT * RANK_TCH) +
                  sizeof(int) * (RANK_UT * RANK_TCH + RANK_UT * RANK_UT * RANK_TCH + RANK_UT * RANK_TCH);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(predict_ranks_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = (test.rows + RANK_UT - 1) / RANK_UT;
    if (grid > 148 * 4) grid = 148 * 4;
    predict_ranks_tiled_kernel<<<grid, RANK_UT * 32, smem, st>>>(usf, test, train, m, repr_t, active, count, ranks);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// exported for the host layer: number of floats predict_ranks needs in `scratch`
size_t lfm_ranks_scratch_floats(int n_items, int d, int test_rows) {
    return (size_t)n_items * (d + 1) + (size_t)test_rows + 4;
}

cudaError_t lfm_launch_auc(const DevCsr& ranks, const int32_t* num_train_pos, float* rank_data,
                           float* auc, cudaStream_t st) {
    if (ranks.rows == 0) return cudaSuccess;
    cudaError_t e;
    if (ranks.nnz > 0) {
        // per-row ascending sort, in place for the caller (T:1352): sort into a temp, copy back
        float* tmp = nullptr;
        e = cudaMallocAsync((void**)&tmp, sizeof(float) * ranks.nnz, st);
        if (e != cudaSuccess) return e;
        size_t bytes = 0;
        cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, rank_data, tmp, (int)ranks.nnz,
                                           ranks.rows, ranks.indptr, ranks.indptr + 1, st);
        void* ws = nullptr;
        e = cudaMallocAsync(&ws, bytes ? bytes : 16, st);
        if (e != cudaSuccess) { cudaFreeAsync(tmp, st); return e; }
        e = cub::DeviceSegmentedSort::SortKeys(ws, bytes, rank_data, tmp, (int)ranks.nnz,
                                               ranks.rows, ranks.indptr, ranks.indptr + 1, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(rank_data, tmp, sizeof(float) * ranks.nnz,
                                cudaMemcpyDeviceToDevice, st);
        cudaFreeAsync(ws, st);
        cudaFreeAsync(tmp, st);
        if (e != cudaSuccess) return e;
    }
    auc_kernel<<<(ranks.rows + 255) / 256, 256, 0, st>>>(ranks, num_train_pos, rank_data, auc);
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
