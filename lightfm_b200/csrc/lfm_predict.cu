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

// One CTA per user with test interactions (persistent loop).
//   A: user representation -> shared
//   B: score every item (thread per item, coalesced over the transposed table) -> scratch row;
//      train positives are then overwritten with NaN (NaN >= x is false, so they never count)
//   C: one warp per test item counts items with score >= its score (T:1317-1319)
__global__ void predict_ranks_kernel(DevCsr usf, DevCsr test, DevCsr train, DevModel m,
                                     const float* __restrict__ item_repr_t, float* scratch,
                                     float* ranks) {
    extern __shared__ float sm[];
    int d = m.d;
    int n_items = test.cols;
    float* u = sm;  // [d+1]
    float* row = scratch + (size_t)blockIdx.x * n_items;
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int user = blockIdx.x; user < test.rows; user += gridDim.x) {
        int ts = test.indptr[user], te = test.indptr[user + 1];
        if (te == ts) continue;
        __syncthreads();
        if (wib == 0) gather_to(usf, m.user.w, m.user.b, d, user, u, 1, lane);
        __syncthreads();
        for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
            float r = u[d] + item_repr_t[(size_t)d * n_items + it];
            for (int j = 0; j < d; j++) r = r + u[j] * item_repr_t[(size_t)j * n_items + it];
            row[it] = r;
        }
        __syncthreads();
        int trs = train.indptr[user], tre = train.indptr[user + 1];
        // Test predictions go to shared memory in chunks of 1024, recomputed from the
        // transposed table (same arithmetic as phase B), so a test item that is also a train
        // positive keeps its own score (T:1283-1298) while its row[] slot is masked.
        for (int base = ts; base < te; base += 1024) {
            int cnt = min(1024, te - base);
            float* tp = sm + (d + 1);  // [1024] predictions of this chunk
            __syncthreads();
            for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
                int id = test.indices[base + t];
                // recompute from the unmasked definition: same arithmetic as above
                float r = u[d] + item_repr_t[(size_t)d * n_items + id];
                for (int j = 0; j < d; j++) r = r + u[j] * item_repr_t[(size_t)j * n_items + id];
                tp[t] = r;
            }
            __syncthreads();
            if (base == ts) {
                for (int t = trs + threadIdx.x; t < tre; t += blockDim.x) {
                    int id = train.indices[t];
                    if (id >= 0 && id < n_items) row[id] = __int_as_float(0x7fc00000);
                }
                __syncthreads();
            }
            for (int t = wib; t < cnt; t += nwarp) {
                int id = test.indices[base + t];
                float p = tp[t];
                int c = 0;
                for (int it = lane; it < n_items; it += 32) c += (it != id && row[it] >= p) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(LFM_FULL, c, o);
                if (lane == 0) ranks[base + t] += (float)c;
            }
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

size_t lfm_item_repr_scratch_floats(const DevCsr& itf, const DevModel& m) {
    // transposed item table + one score row per resident CTA
    return (size_t)itf.rows * (m.d + 1);
}

static int ranks_grid(const DevCsr& test) {
    int g = 148 * 2;
    if (g > test.rows) g = test.rows;
    return g < 1 ? 1 : g;
}

cudaError_t lfm_launch_predict_ranks(const DevCsr& itf, const DevCsr& usf, const DevCsr& test,
                                     const DevCsr& train, const DevModel& m, float* ranks,
                                     float* scratch, cudaStream_t st, int* launches) {
    // scratch layout: [ (d+1) * n_items transposed item table | grid * n_items score rows ]
    int n_items = test.cols;
    if (test.rows == 0 || test.nnz == 0 || n_items == 0) return cudaSuccess;
    float* repr_t = scratch;
    float* rows = scratch + (size_t)n_items * (m.d + 1);
    int64_t blocks = ((int64_t)n_items * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    DevCsr itf_n = itf;
    item_repr_kernel<<<(int)blocks, 256, 0, st>>>(itf_n, m, repr_t, n_items);
    size_t smem = (size_t)(m.d + 1 + 1024) * sizeof(float);
    predict_ranks_kernel<<<ranks_grid(test), 512, smem, st>>>(usf, test, train, m, repr_t, rows, ranks);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

// exported for the host layer: number of floats predict_ranks needs in `scratch`
extern "C" size_t lfm_ranks_scratch_floats(int n_items, int d, int test_rows) {
    int g = 148 * 2;
    if (g > test_rows) g = test_rows;
    if (g < 1) g = 1;
    return (size_t)n_items * (d + 1) + (size_t)g * n_items;
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
