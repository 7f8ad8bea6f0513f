// lfm_common.cuh -- shared device helpers and internal launcher declarations.
//
// Internal to libfm_cuda.so.  Public C ABI: include/lfm_cuda.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lfm_cuda.h"

#define LFM_FULL 0xffffffffu
#define LFM_MAX_REG_SCALE 1000000.0f  // T:19
#define LFM_MAX_LOSS 10.0             // T:817

// ---- device-side views ------------------------------------------------------
struct DevCsr {
    const int32_t* indptr;
    const int32_t* indices;
    const float* data;
    int32_t rows, cols;
    int64_t nnz;
    int32_t identity;  // 1 when the matrix is exactly I (indptr=arange, indices=arange, data=1)
};

struct DevTable {  // one side (item or user) of the FitModel layout
    float* w;   // [n, d] embeddings
    float* g;   // [n, d] adagrad / adadelta accumulator
    float* m;   // [n, d] adadelta momentum (may be null for adagrad)
    float* b;   // [n]
    float* bg;  // [n]
    float* bm;  // [n]
    int32_t n;
};

struct DevModel {
    DevTable item, user;
    int32_t d;
    int32_t adadelta;
    float lr, rho, eps;
    int32_t max_sampled;
};

// Packed, already-shuffled training tuple (built by lfm_pack_tuples).
struct __align__(16) Tuple {
    int32_t user;  // < 0 : skip (Y <= 0 for warp/bpr)
    int32_t item;
    float weight;
    float y;
};

// Device-resident counters; mirrors the integer part of lfm_counters.
struct DevCounters {
    unsigned long long positives, negatives, updates, rejected;
};

// Lazy-regularisation scales (T:213-214) kept in device memory across launches.
struct DevScales {
    double item_scale, user_scale;
};

enum LossKind { LOSS_LOGISTIC = 0, LOSS_WARP = 1, LOSS_BPR = 2, LOSS_KOS = 3 };

struct FitArgs {
    DevCsr itf, usf, pos;          // item features, user features, positives lookup
    DevModel model;
    const int32_t* user_ids;       // raw COO arrays (device)
    const int32_t* item_ids;
    const float* y;
    const float* sample_weight;
    const int32_t* shuffle;        // host-provided order, or null -> device permutation
    int64_t n;                     // no_examples of this launch (an epoch segment)
    int64_t n_all;                 // no_examples of the whole epoch (BPR draws negatives from all of it)
    double item_alpha, user_alpha;
    int32_t k, nkos;               // k-OS parameters
    uint32_t seed;                 // rand_r seed (replay) / philox key (hogwild)
    const double* loss_table;      // [max_sampled + 1] log terms precomputed on the host
    const float* loss_table_f;     // the same terms rounded to float (hogwild kernels)
    int64_t t_offset;              // index of this launch's first tuple in the epoch (Philox counter base)
    int64_t row_offset;            // first interaction (upload order) this epoch visits; it visits n of them
    DevCounters* counters;
    DevScales* scales;
    // Optional exact membership bitmap of the positives CSR: bit (u, i) at
    // pos_bitmap[u * bitmap_words + (i >> 5)] >> (i & 31).  Built once per resident plan when
    // it fits (lfm_launch_build_bitmap); the kernels fall back to the sorted-row search when null.
    const uint32_t* pos_bitmap;
    int32_t bitmap_words;
    // Hot feature rows (lfm_hogwild.cu, feature path): rows shared by many entities (tags) whose
    // updates are aggregated per CTA in shared memory before they reach L2.  hot_slot_*[row] is
    // the row's slot (or -1); hot_rows[slot] = row | (user table ? 1 << 31 : 0).
    const int32_t* hot_slot_item;
    const int32_t* hot_slot_user;
    const int32_t* hot_rows;
    int32_t n_hot;
    // 1 when every Y and every sample weight equals 1.0f (checked on the device when the inputs
    // are staged): pack_kernel then skips two of its three random reads per interaction.
    int32_t unit_weights;
    // Replay mode, dataflow path (lfm_replay_dataflow.cuh): device scratch for the task list, the
    // row version counters and the optional membership bitmap; null / 0 -> sequential replay kernels.
    void* replay_scratch;
    size_t replay_scratch_bytes;
};

// ---- launchers (defined in the .cu files) ------------------------------------
cudaError_t lfm_launch_replay(int loss, const FitArgs& a, cudaStream_t st);
// Bytes of FitArgs::replay_scratch the dataflow replay path wants for (loss, a); 0: not applicable.
size_t lfm_replay_dataflow_scratch_bytes(int loss, const FitArgs& a, int64_t bitmap_limit_bytes);
cudaError_t lfm_launch_hogwild(int loss, const FitArgs& a, Tuple* tuples, cudaStream_t st,
                               int* launches, cudaEvent_t ev_train_begin, cudaEvent_t ev_train_end,
                               cudaStream_t pack_stream = nullptr, cudaEvent_t pack_after = nullptr,
                               cudaEvent_t pack_done = nullptr, bool prepacked = false);
// tuples[i] = {user, item, weight, y}[order[i]]; order = a.shuffle when given, else a Feistel permutation keyed by perm_key
cudaError_t lfm_launch_pack(const FitArgs& a, int loss, Tuple* tuples, uint32_t perm_key, cudaStream_t st);
cudaError_t lfm_launch_regularize(const DevModel& m, DevScales* scales, cudaStream_t st);
// Delta exchange of a replicated table (multi-GPU): up to four (pointer, count)
// segments -- rows [begin, begin+count) of w, g, b, bg -- addressed as one flat range.
struct DeltaSegs {
    float* p[4];
    int64_t n[4];
};
// mode 0: S = cur                      (snapshot before the local epoch)
// mode 1: D = cur - S ; S = D          (local delta, kept in S; D goes to the all-reduce)
// mode 2: cur = cur + D - S            (add what the OTHER ranks did: reduced sum minus own delta)
cudaError_t lfm_launch_delta(int mode, const DeltaSegs& segs, float* S, float* D, cudaStream_t st);
cudaError_t lfm_launch_predict(const DevCsr& itf, const DevCsr& usf, const DevModel& m,
                               const int32_t* user_ids, const int32_t* item_ids, float* out,
                               int64_t n, cudaStream_t st);
cudaError_t lfm_launch_predict_ranks(const DevCsr& itf, const DevCsr& usf, const DevCsr& test,
                                     const DevCsr& train, const DevModel& m, float* ranks,
                                     float* item_repr_scratch, cudaStream_t st, int* launches);
cudaError_t lfm_launch_auc(const DevCsr& ranks, const int32_t* num_train_pos, float* rank_data,
                           float* auc, float* tmp, cudaStream_t st);
cudaError_t lfm_launch_rank_metrics(const DevCsr& test, const float* ranks, int k, int32_t* hits, float* best,
                                    cudaStream_t st);
cudaError_t lfm_launch_row_counts(const DevCsr& m, int32_t* out, int rows, cudaStream_t st);
cudaError_t lfm_launch_recommend(const DevCsr& itf, const DevCsr& usf, const DevCsr* exclude, const DevModel& m,
                                 int n_items, const int32_t* user_ids, int n_users, int k, int32_t* out_items,
                                 float* out_scores, float* scratch, cudaStream_t st, int* launches);
cudaError_t lfm_launch_in_positives(const DevCsr& mat, int32_t row, int32_t col, int32_t* out,
                                    cudaStream_t st);
cudaError_t lfm_launch_check_identity(const DevCsr& m, int32_t* flag, cudaStream_t st);
// Per-feature-row touch counts of one epoch (positives' rows once per interaction, plus
// `neg_per_item` for every row of every item: uniform negatives), for hot-row selection.
cudaError_t lfm_launch_feature_counts(const FitArgs& a, int loss, float* cnt_item, float* cnt_user,
                                      float neg_per_item, cudaStream_t st);
cudaError_t lfm_launch_check_unit(const float* y, const float* w, int64_t n, int32_t* flag, cudaStream_t st);
cudaError_t lfm_launch_build_bitmap_coo(const int32_t* user_ids, const int32_t* item_ids, int64_t n,
                                        uint32_t* bitmap, int32_t rows, int32_t words_per_row, cudaStream_t st);
cudaError_t lfm_launch_build_bitmap(const DevCsr& pos, uint32_t* bitmap, int32_t words_per_row, cudaStream_t st);
// internal helpers shared between translation units (not part of the C ABI)
size_t lfm_ranks_scratch_floats(int n_items, int d, int test_rows);
int lfm_hogwild_supported(int loss, int d, int nkos);
// 1 when the hogwild launch of (loss, a) would take the slot kernels (identity features, adagrad,
// no L2, d in {16,32,64,128}); plans without a positives CSR can only run there.
int lfm_fast_path_eligible(int loss, const FitArgs& a, int64_t count);

#ifdef __CUDACC__
// ---- RNG ---------------------------------------------------------------------
// musl rand_r as restated by the reference (T:64-81); used by replay mode.
__device__ __forceinline__ uint32_t lfm_temper(uint32_t x) {
    x ^= x >> 11;
    x ^= (x << 7) & 0x9D2C5680u;
    x ^= (x << 15) & 0xEFC60000u;
    x ^= x >> 18;
    return x;
}
__device__ __forceinline__ int lfm_rand_r(uint32_t& seed) {
    seed = seed * 1103515245u + 12345u;
    return (int)(lfm_temper(seed) >> 1);
}

// Philox4x32-10 (Salmon et al. 2011), counter-based; used by hogwild mode.
struct Philox4 {
    uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 lfm_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o = {c0, c1, c2, c3};
    return o;
}
// Uniform integer in [0, n) from 32 random bits (multiply-shift; bias < n / 2^32).
__device__ __forceinline__ int lfm_bounded(uint32_t r, uint32_t n) {
    return (int)__umulhi(r, n);
}

// ---- memory helpers --------------------------------------------------------------
// Fire-and-forget reductions performed in L2 (no return value -> RED, not ATOM).
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
// The same reduction WITH the old value returned (ATOMG.ADD.F32x4): an Adagrad step that adds its
// squared gradient this way reads an accumulator that already contains every earlier update of
// that element, however many interactions are in flight.
__device__ __forceinline__ float4 atom_add_v4(float* addr, float a, float b, float c, float d) {
    float4 r;
    asm volatile("atom.global.add.v4.f32 {%0, %1, %2, %3}, [%4], {%5, %6, %7, %8};"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
    return r;
}
// L2-coherent vector load (ld.global.cg): tables are updated concurrently by other SMs.
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg((const float4*)p); }
// G >= 1 under adagrad (starts at 1, only grows): no denormals, so the bare approximation is safe
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- warp helpers --------------------------------------------------------------
__device__ __forceinline__ float lfm_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LFM_FULL, v, o);
    return v;
}

// Sorted-row membership (T:270-284), scalar binary search; every calling lane
// walks the same path, so the loads are broadcasts.
__device__ __forceinline__ bool lfm_bsearch(const int32_t* __restrict__ idx, int lo, int hi,
                                            int key) {
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        int v = __ldg(idx + mid);
        if (v == key) return true;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return false;
}

// Warp-cooperative membership: 32-ary narrowing, then one 32-wide probe.
// All 32 lanes must call it with identical arguments; returns the same value on all lanes.
__device__ __forceinline__ bool lfm_warp_member(const int32_t* __restrict__ idx, int lo, int hi,
                                                int key, int lane) {
    while (hi - lo > 32) {
        int len = hi - lo;
        // pivots split [lo,hi) into 33 nearly equal segments
        int p = lo + (int)(((long long)len * (lane + 1)) / 33);
        int v = __ldg(idx + p);
        unsigned le = __ballot_sync(LFM_FULL, v <= key);  // monotone: 1..1 0..0
        int c = __popc(le);
        int nlo = (c == 0) ? lo : __shfl_sync(LFM_FULL, p, c - 1);
        int nhi = (c == 32) ? hi : __shfl_sync(LFM_FULL, p, c) ;
        // key, if present, lies in [nlo, nhi]; pivot c (first with v>key) is excluded
        lo = nlo;
        hi = (c == 32) ? hi : nhi;
    }
    int v = (lo + lane < hi) ? __ldg(idx + lo + lane) : -1;
    return __any_sync(LFM_FULL, (lo + lane < hi) && v == key);
}
#endif  // __CUDACC__
