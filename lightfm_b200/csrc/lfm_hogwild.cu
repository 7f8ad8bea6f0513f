// lfm_hogwild.cu -- throughput mode (num_threads > 1): lock-free concurrent SGD over the whole GPU.
//
// The reference runs the SGD loop under an OpenMP prange with racy, lock-free
// read-modify-writes on the shared tables (Hogwild; T:825 and SURVEY 0).  Here the epoch is
//   1. pack_kernel: tuples[i] = {user, item, weight, y}[order[i]] (order = host shuffle, or a
//      Feistel bijection of [0, n) generated in registers);
//   2. one SGD kernel over the tuples: gather user / item rows (L2-coherent ld.global.cg),
//      dot product by warp shuffles, Philox negative draws, the WARP rank-sampling loop in
//      registers, Adagrad-scaled deltas scattered with red.global.add (no lost updates).
// Floating point is fp32 with FMA; results are statistically equivalent to the reference's
// multi-thread runs, not bit-equal (nor is the reference to itself).
//
// Two kernel families:
//   generic (this file)      lanes own components; any d <= 256, any feature CSR, adagrad /
//                            adadelta, L2 regularisation (log-domain lazy scale).
//   fast (lfm_hogwild_fast.cuh)  identity features, adagrad, alpha == 0, d in {16,32,64,128}:
//                            slot-per-interaction kernels with cp.async staging.
//
// Reference: fit_logistic T:694-781, fit_warp T:784-912, fit_warp_kos T:915-1071,
// fit_bpr T:1074-1182, update T:454-534, warp_update T:537-649.
#include <atomic>

#include "lfm_common.cuh"

namespace {

// ---- device-side permutation (used when no host shuffle is supplied) ---------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// 4-round Feistel network on 2*hb bits + cycle walking: a bijection of [0, n).
__device__ __forceinline__ int64_t feistel_perm(int64_t i, int64_t n, int hb, uint32_t key) {
    uint32_t mask = (1u << hb) - 1u;
    uint64_t x = (uint64_t)i;
    do {
        uint32_t l = (uint32_t)(x >> hb) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (int round = 0; round < 4; round++) {
            uint32_t f = mix32(r ^ (key + 0x9E3779B9u * (round + 1))) & mask;
            uint32_t t = l ^ f;
            l = r;
            r = t;
        }
        x = ((uint64_t)l << hb) | r;
    } while (x >= (uint64_t)n);
    return (int64_t)x;
}

__global__ void pack_kernel(const int32_t* __restrict__ user_ids, const int32_t* __restrict__ item_ids,
                            const float* __restrict__ y, const float* __restrict__ w,
                            const int32_t* __restrict__ shuffle, int64_t n, int hb, uint32_t key,
                            int skip_nonpositive, int unit_weights, int64_t row_offset, Tuple* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int64_t row = shuffle ? (int64_t)shuffle[i] : row_offset + feistel_perm(i, n, hb, key);
        Tuple t;
        t.user = user_ids[row];
        t.item = item_ids ? item_ids[row] : (int32_t)i;  // k-OS: the tuple's index in the epoch
        t.y = (y && !unit_weights) ? y[row] : 1.0f;
        t.weight = (w && !unit_weights) ? w[row] : 1.0f;
        if (skip_nonpositive && !(t.y > 0)) t.user = -1;
        out[i] = t;
    }
}

__global__ void check_unit_kernel(const float* __restrict__ y, const float* __restrict__ w, int64_t n,
                                  int32_t* flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) bad |= (y[i] != 1.0f) || (w != y && w[i] != 1.0f);
    if (bad) *flag = 0;
}

__global__ void build_bitmap_coo_kernel(const int32_t* __restrict__ user_ids, const int32_t* __restrict__ item_ids,
                                        int64_t n, uint32_t* bitmap, int words) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride) {
        const int item = item_ids[e];
        atomicOr(bitmap + (size_t)user_ids[e] * words + (item >> 5), 1u << (item & 31));
    }
}

// One thread per stored entry of the positives CSR: set its bit (row found by binary search in indptr).
__global__ void build_bitmap_kernel(DevCsr pos, uint32_t* bitmap, int words) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < pos.nnz; e += stride) {
        int lo = 0, hi = pos.rows;  // largest row with indptr[row] <= e
        while (hi - lo > 1) {
            int mid = lo + ((hi - lo) >> 1);
            if ((int64_t)__ldg(pos.indptr + mid) <= e) lo = mid; else hi = mid;
        }
        int item = __ldg(pos.indices + e);
        atomicOr(bitmap + (size_t)lo * words + (item >> 5), 1u << (item & 31));
    }
}

// ---- generic kernel ------------------------------------------------------------
template <int KPL>
struct Repr {
    float v[KPL];
    float b;
};

// Gather one entity's representation.  A lane owns KPL registers: with VW == 1 register k is
// component lane + 32*k; with VW == 4 (d % 4 == 0, 16 B-aligned rows) registers 4j..4j+3 are the
// float4 chunk lane + 32*j of the row, fetched and reduced as one 16 B access.
template <int KPL, int VW>
__device__ __forceinline__ void gather(const DevCsr& f, const DevTable& t, int d, int row,
                                       float scale, Repr<KPL>& r, int lane) {
#pragma unroll
    for (int k = 0; k < KPL; k++) r.v[k] = 0.0f;
    r.b = 0.0f;
    int start, stop;
    if (f.identity) { start = row; stop = row + 1; }
    else { start = __ldg(f.indptr + row); stop = __ldg(f.indptr + row + 1); }
    for (int i = start; i < stop; i++) {
        int ft = f.identity ? row : __ldg(f.indices + i);
        float fw = (f.identity ? 1.0f : __ldg(f.data + i)) * scale;
        const float* p = t.w + (size_t)ft * d;
        if (VW == 4) {
#pragma unroll
            for (int j = 0; j < KPL / 4; j++) {
                int c = (lane + 32 * j) * 4;
                if (c < d) {
                    float4 x = ldcg4(p + c);
                    r.v[4 * j + 0] = fmaf(fw, x.x, r.v[4 * j + 0]);
                    r.v[4 * j + 1] = fmaf(fw, x.y, r.v[4 * j + 1]);
                    r.v[4 * j + 2] = fmaf(fw, x.z, r.v[4 * j + 2]);
                    r.v[4 * j + 3] = fmaf(fw, x.w, r.v[4 * j + 3]);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KPL; k++) {
                int c = lane + 32 * k;
                if (c < d) r.v[k] = fmaf(fw, __ldcg(p + c), r.v[k]);
            }
        }
        r.b = fmaf(fw, __ldcg(t.b + ft), r.b);
    }
}

template <int KPL>
__device__ __forceinline__ float dot(const Repr<KPL>& a, const Repr<KPL>& b) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < KPL; k++) s = fmaf(a.v[k], b.v[k], s);  // lanes past d hold zeros
    return lfm_warp_sum(s) + a.b + b.b;
}

// One parameter step in hogwild arithmetic.  Adagrad: pure adds -> red.global.add.
// Adadelta: not expressible as adds -> racy plain RMW, as in the reference.
// Returns the local learning rate (for the lazy-regularisation average).
template <bool ADADELTA>
__device__ __forceinline__ float hstep(float* w, float* G, float* M, float fw, float grad,
                                       const DevModel& m, float alpha) {
    float llr;
    if (ADADELTA) {
        float g0 = __ldcg(G), m0 = __ldcg(M), w0 = __ldcg(w);
        float t = fw * grad;
        float g1 = m.rho * g0 + (1.0f - m.rho) * t * t;
        llr = sqrtf(m0 + m.eps) / sqrtf(g1 + m.eps);
        float upd = llr * grad * fw;
        float m1 = m.rho * m0 + (1.0f - m.rho) * upd * upd;
        float w1 = (w0 - upd) * (1.0f + alpha * llr);
        __stcg(G, g1);
        __stcg(M, m1);
        __stcg(w, w1);
    } else {
        // the accumulator as every earlier update left it (atomic add with return), then the step
        float gw = grad * fw;
        float g0 = atomicAdd(G, gw * gw);
        llr = m.lr * rsqrtf(g0);
        float delta = -llr * gw;
        if (alpha != 0.0f) {
            float w0 = __ldcg(w);
            delta += (w0 + delta) * (alpha * llr);
        }
        atomicAdd(w, delta);
    }
    return llr;
}

// Apply `grad[k]` (per owned component) and `bgrad` to every feature row of `row`.
template <int KPL, int VW, bool ADADELTA>
__device__ __forceinline__ float scatter(const DevCsr& f, DevTable& t, const DevModel& m, int row,
                                         const float (&grad)[KPL], float bgrad, float alpha,
                                         int lane, int& nnz) {
    float lrsum = 0.0f;
    int d = m.d;
    int start, stop;
    if (f.identity) { start = row; stop = row + 1; }
    else { start = __ldg(f.indptr + row); stop = __ldg(f.indptr + row + 1); }
    nnz = stop - start;
    for (int i = start; i < stop; i++) {
        int ft = f.identity ? row : __ldg(f.indices + i);
        float fw = f.identity ? 1.0f : __ldg(f.data + i);
        size_t o = (size_t)ft * d;
        if (VW == 4) {  // adagrad, alpha == 0 (guaranteed by the launcher): vector reductions
#pragma unroll
            for (int j = 0; j < KPL / 4; j++) {
                int c = (lane + 32 * j) * 4;
                if (c < d) {
                    const float gx = grad[4 * j] * fw, gy = grad[4 * j + 1] * fw, gz = grad[4 * j + 2] * fw,
                                gw = grad[4 * j + 3] * fw;
                    const float4 g0 = atom_add_v4(t.g + o + c, gx * gx, gy * gy, gz * gz, gw * gw);
                    red_add_v4(t.w + o + c, -m.lr * rsqrt_ftz(g0.x) * gx, -m.lr * rsqrt_ftz(g0.y) * gy,
                               -m.lr * rsqrt_ftz(g0.z) * gz, -m.lr * rsqrt_ftz(g0.w) * gw);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KPL; k++) {
                int c = lane + 32 * k;
                if (c < d)
                    lrsum += hstep<ADADELTA>(t.w + o + c, t.g + o + c, ADADELTA ? t.m + o + c : nullptr,
                                             fw, grad[k], m, alpha);
            }
        }
        if (lane == 0)
            lrsum += hstep<ADADELTA>(t.b + ft, t.bg + ft, ADADELTA ? t.bm + ft : nullptr, fw, bgrad,
                                     m, alpha);
    }
    return lrsum;
}


// ---- feature path: batched row traffic + CTA-level aggregation of hot rows ----------------
// With shared feature rows (tags) the epoch is bound by L2 serialising same-line reductions:
// a tag carried by 40% of the items receives ~0.7 row-updates per interaction, i.e. ~10^7
// reductions per epoch on each of its eight 128 B lines, all processed one after another by one
// L2 slice (C3: 367 ms per epoch, 0.24 of the HBM roofline).  The HOT variant keeps, per CTA,
// a shared-memory accumulator (delta of w, delta of G, bias deltas) for the n_hot most-touched
// feature rows; a warp adds its update under a per-slot spin lock with plain vector
// loads/stores (shared memory has no vector atomics), and every warp flushes one slot per
// interaction with a single red.global per chunk, so a hot line sees one reduction per
// ~K/warps interactions of a CTA instead of one per touch.  A lock that cannot be taken within
// a few tries falls back to the direct global reduction: never a wait, never a lost update.
// Forward reads use the global rows (missing at most the deltas pending in shared memory:
// the same order of staleness as the interactions in flight).
struct HotSmem {
    float4* acc;   // [n_hot][2 * d4 + 1]: w-delta chunks, G-delta chunks, {db, dbg, -, -}
    int* locks;    // [n_hot]
    int stride;    // float4 per slot = 2 * d4 + 1
    int d4;        // float4 chunks per row = d / 4
};

__device__ __forceinline__ bool hot_lock(int* lock, int lane) {
    int got = 0;
    if (lane == 0) {
#pragma unroll 1
        for (int tries = 0; tries < 16 && !got; tries++) got = atomicCAS(lock, 0, 1) == 0;
    }
    got = __shfl_sync(LFM_FULL, got, 0);
    if (got) __threadfence_block();  // acquire
    return got != 0;
}
__device__ __forceinline__ void hot_unlock(int* lock, int lane) {
    __threadfence_block();  // release: this warp's stores to the slot before the lock word
    __syncwarp();
    if (lane == 0) atomicExch(lock, 0);
}

// Drain one slot into the global tables (called with the slot locked, or after the final barrier).
template <int NCH>
__device__ __forceinline__ void hot_drain(const HotSmem& h, int slot, const FitArgs& a, int lane) {
    const int32_t tag = __ldg(a.hot_rows + slot);
    const DevTable& t = (tag < 0) ? a.model.user : a.model.item;
    const int row = tag & 0x7fffffff;
    float4* base = h.acc + (size_t)slot * h.stride;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NCH; j++) {
        const int c = lane + 32 * j;
        if (c < h.d4) {
            const float4 dw = base[c], dg = base[h.d4 + c];
            if (dg.x != 0.f || dg.y != 0.f || dg.z != 0.f || dg.w != 0.f) {
                base[c] = z;
                base[h.d4 + c] = z;
                red_add_v4(t.w + (size_t)row * a.model.d + c * 4, dw.x, dw.y, dw.z, dw.w);
                red_add_v4(t.g + (size_t)row * a.model.d + c * 4, dg.x, dg.y, dg.z, dg.w);
            }
        }
    }
    if (lane == 0) {
        const float4 db = base[2 * h.d4];
        if (db.y != 0.f) {
            base[2 * h.d4] = z;
            red_add(t.b + row, db.x);
            red_add(t.bg + row, db.y);
        }
    }
}

// A row's feature list as the lanes hold it after a gather (feature i of the row on lane i):
// reused by the scatter of the same row, which then needs no index loads of its own.
struct FeatRow {
    int ft;     // feature (embedding row) id
    float fw;   // feature weight
    int hs;     // hot slot of that embedding row, or -1
    int cnt;    // number of features (<= 32), or -1 when the row was too long to keep
};

#define FB 4   // feature rows in flight per batch (10 in flight made the kernel twice as large and slower:
               // profiles/README.md, C3 steps v2-v5)

// Gather in the float4 layout with the row loads of up to FB features in flight at a time
// (the generic gather walks the features one dependent round trip after the other).
template <int KPL>
__device__ __forceinline__ void gather_b(const DevCsr& f, const DevTable& t, const int32_t* __restrict__ hot_slot,
                                         int d, int row, Repr<KPL>& r, FeatRow& fr, int lane) {
    constexpr int NCH = KPL / 4;
#pragma unroll
    for (int k = 0; k < KPL; k++) r.v[k] = 0.0f;
    r.b = 0.0f;
    const int d4 = d >> 2;
    if (f.identity) {
        fr.ft = row; fr.fw = 1.0f; fr.cnt = 1;
        fr.hs = (hot_slot != nullptr && lane == 0) ? __ldg(hot_slot + row) : -1;
#pragma unroll
        for (int j = 0; j < NCH; j++) {
            const int c = lane + 32 * j;
            if (c < d4) {
                const float4 x = ldcg4(t.w + (size_t)row * d + c * 4);
                r.v[4 * j] = x.x; r.v[4 * j + 1] = x.y; r.v[4 * j + 2] = x.z; r.v[4 * j + 3] = x.w;
            }
        }
        r.b = __ldcg(t.b + row);
        return;
    }
    const int start = __ldg(f.indptr + row), stop = __ldg(f.indptr + row + 1);
    fr.cnt = (stop - start <= 32) ? stop - start : -1;
    fr.ft = 0; fr.fw = 0.0f; fr.hs = -1;
    float bsum = 0.0f;
    for (int base = start; base < stop; base += 32) {
        const int cnt = min(32, stop - base);
        const int my_ft = lane < cnt ? __ldg(f.indices + base + lane) : 0;
        const float my_fw = lane < cnt ? __ldg(f.data + base + lane) : 0.0f;
        if (lane < cnt) {
            bsum = fmaf(my_fw, __ldcg(t.b + my_ft), bsum);
            if (hot_slot != nullptr) fr.hs = __ldg(hot_slot + my_ft);
        }
        fr.ft = my_ft; fr.fw = my_fw;
#pragma unroll 1
        for (int i0 = 0; i0 < cnt; i0 += FB) {
            float4 x[FB][NCH];
            float w[FB];
#pragma unroll
            for (int k = 0; k < FB; k++) {
                const int i = i0 + k;
                const int ft = __shfl_sync(LFM_FULL, my_ft, i & 31);
                w[k] = __shfl_sync(LFM_FULL, my_fw, i & 31);
                if (i >= cnt) w[k] = 0.0f;
#pragma unroll
                for (int j = 0; j < NCH; j++) {
                    const int c = lane + 32 * j;
                    x[k][j] = (i < cnt && c < d4) ? ldcg4(t.w + (size_t)ft * d + c * 4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int k = 0; k < FB; k++) {
#pragma unroll
                for (int j = 0; j < NCH; j++) {
                    r.v[4 * j] = fmaf(w[k], x[k][j].x, r.v[4 * j]);
                    r.v[4 * j + 1] = fmaf(w[k], x[k][j].y, r.v[4 * j + 1]);
                    r.v[4 * j + 2] = fmaf(w[k], x[k][j].z, r.v[4 * j + 2]);
                    r.v[4 * j + 3] = fmaf(w[k], x[k][j].w, r.v[4 * j + 3]);
                }
            }
        }
    }
    r.b = lfm_warp_sum(bsum);
}


// ---- feature path, WARP: the negative candidates' feature lists are fetched ahead -------------
// The Philox draws of an interaction are known up front, so the item ids of the next candidates
// are known before the current one is scored.  A candidate costs three dependent loads
// (indptr pair -> indices/data -> embedding rows); stages one and two of the NEXT candidates run
// under the row fetch and scoring of the current one: the row bounds wait in registers, the
// (feature id, weight) lists stream into a per-warp shared-memory ring with cp.async.
#define FR_RING 3
#define FR_STAGED 0  // measured on C3: 301 ms per epoch with the staging vs 247 ms without (the extra
                     // instructions and the out-of-line gathers cost more than the hidden round trips)
__device__ __forceinline__ void fr_cp_async4(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fr_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fr_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void fr_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// stage two: stream the (feature id, weight) list of a row into ring slot `slot` (rows of <= 32 features)
__device__ __forceinline__ void fr_stage_list(const DevCsr& f, int start, int cnt, int2* ring, int slot, int lane) {
    if (cnt >= 0 && lane < cnt) {
        fr_cp_async4(&ring[slot * 32 + lane].x, f.indices + start + lane);
        fr_cp_async4(&ring[slot * 32 + lane].y, f.data + start + lane);
    }
    fr_commit();
}

// stage three: the rows of a staged list (the lane's entry of the list is read back from the ring)
template <int KPL>
__device__ __noinline__ void gather_staged(const DevTable& t, const int32_t* __restrict__ hot_slot, int d, int cnt,
                                           const int2* ring, int slot, Repr<KPL>& r, FeatRow& fr, int lane) {
    constexpr int NCH = KPL / 4;
#pragma unroll
    for (int k = 0; k < KPL; k++) r.v[k] = 0.0f;
    const int d4 = d >> 2;
    const int2 e = ring[slot * 32 + lane];
    const int my_ft = lane < cnt ? e.x : 0;
    const float my_fw = lane < cnt ? __int_as_float(e.y) : 0.0f;
    fr.ft = my_ft; fr.fw = my_fw; fr.cnt = cnt;
    fr.hs = (hot_slot != nullptr && lane < cnt) ? __ldg(hot_slot + my_ft) : -1;
    float bsum = lane < cnt ? my_fw * __ldcg(t.b + my_ft) : 0.0f;
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += FB) {
        float4 x[FB][NCH];
        float w[FB];
#pragma unroll
        for (int k = 0; k < FB; k++) {
            const int i = i0 + k;
            const int ft = __shfl_sync(LFM_FULL, my_ft, i & 31);
            w[k] = __shfl_sync(LFM_FULL, my_fw, i & 31);
            if (i >= cnt) w[k] = 0.0f;
#pragma unroll
            for (int j = 0; j < NCH; j++) {
                const int c = lane + 32 * j;
                x[k][j] = (i < cnt && c < d4) ? ldcg4(t.w + (size_t)ft * d + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < FB; k++) {
#pragma unroll
            for (int j = 0; j < NCH; j++) {
                r.v[4 * j] = fmaf(w[k], x[k][j].x, r.v[4 * j]);
                r.v[4 * j + 1] = fmaf(w[k], x[k][j].y, r.v[4 * j + 1]);
                r.v[4 * j + 2] = fmaf(w[k], x[k][j].z, r.v[4 * j + 2]);
                r.v[4 * j + 3] = fmaf(w[k], x[k][j].w, r.v[4 * j + 3]);
            }
        }
    }
    r.b = lfm_warp_sum(bsum);
}

// One feature row's share of an Adagrad update (alpha == 0), accumulator chunk already in g0.
// Hot rows go to the CTA's shared-memory accumulator under the slot lock, everything else (and
// a hot row whose lock is busy) straight to L2.  One float4 chunk per lane (d <= 128); arguments
// by value so that they stay in registers across the call.
__device__ __forceinline__ void apply_row(float* tw, float* tg, float* tb, float* tbg, float4* hacc, int* hlocks,
                                       int hstride, int d, float lr, int ft, float fw, int hs, float4 g0,
                                       float4 grad, float bgrad, int lane) {
    const int d4 = d >> 2;
    const int c = lane;
    const float gx = grad.x * fw, gy = grad.y * fw, gz = grad.z * fw, gw = grad.w * fw;
    if (hs >= 0) {
        int* lock = hlocks + hs;
        if (hot_lock(lock, lane)) {
            float4* sb = hacc + (size_t)hs * hstride;
            if (c < d4) {
                float4 aw = sb[c], ag = sb[d4 + c];
                aw.x -= lr * rsqrt_ftz(g0.x) * gx; aw.y -= lr * rsqrt_ftz(g0.y) * gy;
                aw.z -= lr * rsqrt_ftz(g0.z) * gz; aw.w -= lr * rsqrt_ftz(g0.w) * gw;
                ag.x = fmaf(gx, gx, ag.x); ag.y = fmaf(gy, gy, ag.y);
                ag.z = fmaf(gz, gz, ag.z); ag.w = fmaf(gw, gw, ag.w);
                sb[c] = aw;
                sb[d4 + c] = ag;
            }
            if (lane == 0) {
                const float g = bgrad * fw;
                const float bg0 = __ldcg(tbg + ft);
                float4 ab = sb[2 * d4];
                ab.x -= lr * rsqrt_ftz(bg0) * g;
                ab.y = fmaf(g, g, ab.y);
                sb[2 * d4] = ab;
            }
            hot_unlock(lock, lane);
            return;
        }
    }
    if (c < d4) {
        const size_t o = (size_t)ft * d + c * 4;
        red_add_v4(tw + o, -lr * rsqrt_ftz(g0.x) * gx, -lr * rsqrt_ftz(g0.y) * gy, -lr * rsqrt_ftz(g0.z) * gz,
                   -lr * rsqrt_ftz(g0.w) * gw);
        if (hs >= 0) red_add_v4(tg + o, gx * gx, gy * gy, gz * gz, gw * gw);  // (direct rows: added by the caller's atomic)
    }
    if (hs >= 0 && lane == 0) {  // slot busy: its bias goes the direct way as well
        const float g = bgrad * fw;
        const float bg0 = atomicAdd(tbg + ft, g * g);
        red_add(tb + ft, -lr * rsqrt_ftz(bg0) * g);
    }
}

// Adagrad scatter (alpha == 0) in the float4 layout: the accumulator rows of up to FB features
// are fetched together; `fr` (from the gather of the same row) saves the index loads.
template <int KPL>
__device__ __forceinline__ void scatter_b(const DevCsr& f, DevTable& t, const int32_t* __restrict__ hot_slot,
                                          const HotSmem& h, const DevModel& m, int row, const FeatRow& fr,
                                          const float (&grad)[KPL], float bgrad, int lane) {
    static_assert(KPL == 4, "the hot-row path holds one float4 chunk per lane (d <= 128)");
    constexpr int NCH = KPL / 4;
    const int d = m.d, d4 = d >> 2;
    const float lr = m.lr;
    int start = 0, stop = fr.cnt;
    const bool reuse = fr.cnt >= 0;
    if (!reuse) { start = __ldg(f.indptr + row); stop = __ldg(f.indptr + row + 1); }
    for (int base = start; base < stop; base += 32) {
        const int cnt = min(32, stop - base);
        int my_ft = fr.ft, my_hs = fr.hs;
        float my_fw = fr.fw;
        if (!reuse) {
            my_ft = lane < cnt ? __ldg(f.indices + base + lane) : 0;
            my_fw = lane < cnt ? __ldg(f.data + base + lane) : 0.0f;
            my_hs = (hot_slot != nullptr && lane < cnt) ? __ldg(hot_slot + my_ft) : -1;
        }
        if (lane >= cnt) my_hs = -1;
        // biases of the rows without a slot: one feature per lane, all in flight together
        if (lane < cnt && my_hs < 0) {
            const float g = bgrad * my_fw;
            const float g0 = atomicAdd(t.bg + my_ft, g * g);
            red_add(t.b + my_ft, -lr * rsqrt_ftz(g0) * g);
        }
#pragma unroll 1
        for (int i0 = 0; i0 < cnt; i0 += FB) {
            float4 g0[FB][NCH];
#pragma unroll
            for (int k = 0; k < FB; k++) {
                const int i = i0 + k;
                const int ft = __shfl_sync(LFM_FULL, my_ft, i & 31);
#pragma unroll
                for (int j = 0; j < NCH; j++) {
                    const int c = lane + 32 * j;
                    g0[k][j] = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (i < cnt && c < d4) {
                        const int hs = __shfl_sync(LFM_FULL, my_hs, i & 31);
                        float* G = t.g + (size_t)ft * d + c * 4;
                        if (hs >= 0) {
                            g0[k][j] = ldcg4(G);     // slot row: the accumulator delta goes to shared memory
                        } else {                     // direct row: add g^2 now and keep what was there before
                            const float fw = __shfl_sync(LFM_FULL, my_fw, i & 31);
                            const float gx = grad[4 * j] * fw, gy = grad[4 * j + 1] * fw, gz = grad[4 * j + 2] * fw,
                                        gw = grad[4 * j + 3] * fw;
                            g0[k][j] = atom_add_v4(G, gx * gx, gy * gy, gz * gz, gw * gw);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < FB; k++) {
                const int i = i0 + k;
                if (i >= cnt) break;  // warp-uniform
                const int ft = __shfl_sync(LFM_FULL, my_ft, i & 31);
                const float fw = __shfl_sync(LFM_FULL, my_fw, i & 31);
                const int hs = __shfl_sync(LFM_FULL, my_hs, i & 31);
                apply_row(t.w, t.g, t.b, t.bg, h.acc, h.locks, h.stride, d, lr, ft, fw, hs, g0[k][0],
                          make_float4(grad[0], grad[1], grad[2], grad[3]), bgrad, lane);
            }
        }
    }
}

struct RegState {  // per-warp view of the lazy-regularisation scales (log domain)
    double base_i, base_u;  // last value read from global
    double loc_i, loc_u;    // local contribution not yet flushed
    int pending;
};

template <int LOSS, int KPL, int VW, bool ADADELTA, bool REG, bool HOT = false>
__global__ void __launch_bounds__(HOT ? 512 : 256) hogwild_kernel(FitArgs a, const Tuple* __restrict__ tuples) {
    static_assert(!HOT || (VW == 4 && !ADADELTA && !REG), "the hot-row path is adagrad, alpha == 0, float4 layout");
    extern __shared__ __align__(16) unsigned char hot_raw[];
    HotSmem hsm = {nullptr, nullptr, 0, 0};
    if constexpr (HOT) {
        hsm.d4 = a.model.d >> 2;
        hsm.stride = 2 * hsm.d4 + 1;
        hsm.acc = (float4*)hot_raw;
        hsm.locks = (int*)(hsm.acc + (size_t)a.n_hot * hsm.stride);
        const int words = a.n_hot * hsm.stride * 4 + a.n_hot;
        // (the per-warp candidate rings follow, 8 B aligned: see fr_ring below)
        for (int i = threadIdx.x; i < words; i += blockDim.x) ((int*)hot_raw)[i] = 0;
        __syncthreads();
    }
    int flush_iter = 0;
    const int lane = threadIdx.x & 31;
    int2* fr_ring = nullptr;  // [FR_RING][32] (feature id, weight bits) per warp: staged candidate lists
    if constexpr (HOT) {
        const size_t off = ((size_t)a.n_hot * (hsm.stride * 16 + 4) + 15) & ~(size_t)15;
        fr_ring = (int2*)(hot_raw + off) + (size_t)(threadIdx.x >> 5) * FR_RING * 32;
    }
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    DevModel& m = a.model;
    const int d = m.d;
    const float alpha_i = (float)a.item_alpha, alpha_u = (float)a.user_alpha;
    const int n_items = a.itf.rows;
    unsigned long long c_pos = 0, c_neg = 0, c_upd = 0, c_rej = 0;

    RegState rs;
    rs.base_i = rs.base_u = rs.loc_i = rs.loc_u = 0.0;
    rs.pending = 0;
    if (REG) {
        rs.base_i = __ldcg(&a.scales->item_scale);  // log-scale in hogwild mode
        rs.base_u = __ldcg(&a.scales->user_scale);
    }

    FeatRow fr_u = {0, 0.f, -1, -1}, fr_p = fr_u, fr_q = fr_u;  // feature lists kept from the gathers (HOT)
#define GATHER(F, T, SLOTS, ROW, SCALE, R, FR)                               \
    do {                                                                     \
        if constexpr (HOT) gather_b<KPL>(F, T, SLOTS, d, ROW, R, FR, lane);  \
        else gather<KPL, VW>(F, T, d, ROW, SCALE, R, lane);                  \
    } while (0)
#define SCATTER(F, T, SLOTS, ROW, FR, GRAD, BGRAD, ALPHA, NNZ)                                        \
    do {                                                                                              \
        if constexpr (HOT) { scatter_b<KPL>(F, T, SLOTS, hsm, m, ROW, FR, GRAD, BGRAD, lane); NNZ = 0; } \
        else lrsum += scatter<KPL, VW, ADADELTA>(F, T, m, ROW, GRAD, BGRAD, ALPHA, lane, NNZ);        \
    } while (0)
    for (int64_t tl = warp; tl < a.n; tl += nwarps) {
        if constexpr (HOT) {
            // every warp drains one slot per interaction, round robin over the CTA's slots
            const int wpb = blockDim.x >> 5;
            const int slot = (int)(((unsigned)flush_iter * (unsigned)wpb + (threadIdx.x >> 5)) % (unsigned)a.n_hot);
            flush_iter++;
            if (hot_lock(hsm.locks + slot, lane)) {
                hot_drain<KPL / 4>(hsm, slot, a, lane);
                hot_unlock(hsm.locks + slot, lane);
            }
        }
        Tuple tp = tuples[tl];
        if (tp.user < 0) continue;
        const int64_t t = tl + a.t_offset;  // index in the epoch: segments must not replay one stream
        float item_scale = 1.0f, user_scale = 1.0f;
        if (REG) {
            // clamped well above the ln(1e6) rescale trigger and below float overflow (e^88)
            item_scale = (float)exp(fmin(rs.base_i + rs.loc_i, 60.0));
            user_scale = (float)exp(fmin(rs.base_u + rs.loc_u, 60.0));
        }
        const int user = tp.user;
        Repr<KPL> u, p, q;
        GATHER(a.usf, m.user, a.hot_slot_user, user, user_scale, u, fr_u);
        float lrsum = 0.0f;
        int nnz_total = 0;
        bool updated = false;

        if (LOSS == LOSS_LOGISTIC) {
            GATHER(a.itf, m.item, a.hot_slot_item, tp.item, item_scale, p, fr_p);
            float pred = 1.0f / (1.0f + __expf(-dot<KPL>(u, p)));
            float loss = tp.weight * (pred - (tp.y > 0 ? 1.0f : 0.0f));
            float gi[KPL], gu[KPL];
#pragma unroll
            for (int k = 0; k < KPL; k++) { gi[k] = loss * u.v[k]; gu[k] = loss * p.v[k]; }
            int n1, n2;
            SCATTER(a.itf, m.item, a.hot_slot_item, tp.item, fr_p, gi, loss, alpha_i, n1);
            SCATTER(a.usf, m.user, a.hot_slot_user, user, fr_u, gu, loss, alpha_u, n2);
            nnz_total = n1 + n2;
            updated = true;
            c_pos++; c_upd++;
        } else {
            int pos_id = tp.item;
            float pp = 0.0f;
            const int ps = __ldg(a.pos.indptr + user), pe = __ldg(a.pos.indptr + user + 1);
            uint32_t ctr = 0;  // philox block counter for this tuple
            Philox4 rnd = lfm_philox((uint32_t)t, (uint32_t)(t >> 32), ctr++, 0u, a.seed, 0x4c464d31u);
            int rpos = 0;
            auto next_u32 = [&]() -> uint32_t {
                if (rpos == 4) {
                    rnd = lfm_philox((uint32_t)t, (uint32_t)(t >> 32), ctr++, 0u, a.seed, 0x4c464d31u);
                    rpos = 0;
                }
                uint32_t r = rpos == 0 ? rnd.x : rpos == 1 ? rnd.y : rpos == 2 ? rnd.z : rnd.w;
                rpos++;
                return r;
            };

            if (LOSS == LOSS_KOS) {
                if (pe == ps) continue;
                int no_pos = min(a.nkos, pe - ps);  // host guarantees nkos <= 32 in this mode
                int my_idx = 0;
                float my_val = 0.0f;
                for (int j = 0; j < no_pos; j++) {
                    int sid = __ldg(a.pos.indices + ps + lfm_bounded(next_u32(), (uint32_t)(pe - ps)));
                    GATHER(a.itf, m.item, a.hot_slot_item, sid, item_scale, p, fr_p);
                    float s = dot<KPL>(u, p);
                    if (lane == j) { my_idx = sid; my_val = s; }
                }
                // position in the stable descending order == what qsort(reverse_pair_compare) yields
                int rank = 0;
                for (int j = 0; j < no_pos; j++) {
                    float vj = __shfl_sync(LFM_FULL, my_val, j);
                    rank += (vj > my_val || (vj == my_val && j < lane)) ? 1 : 0;
                }
                int sel = min(a.k, no_pos) - 1;
                unsigned hit = __ballot_sync(LFM_FULL, lane < no_pos && rank == sel);
                int src = __ffs(hit) - 1;
                if (src < 0) src = 0;  // NaN scores: fall back to the first sample
                pos_id = __shfl_sync(LFM_FULL, my_idx, src);
                pp = __shfl_sync(LFM_FULL, my_val, src);
                GATHER(a.itf, m.item, a.hot_slot_item, pos_id, item_scale, p, fr_p);
                c_pos++;
            } else if constexpr (HOT && LOSS == LOSS_WARP && FR_STAGED) {
                // positive item handled together with the first candidates below
                c_pos++;
            } else {
                GATHER(a.itf, m.item, a.hot_slot_item, pos_id, item_scale, p, fr_p);
                pp = dot<KPL>(u, p);
                c_pos++;
            }

            int neg_id = -1;
            float loss = 0.0f;
            if constexpr (HOT && LOSS == LOSS_WARP && FR_STAGED) {
                // ---- rank sampling with the candidates' feature lists fetched ahead ----
                const DevCsr& f = a.itf;
                auto bounds = [&](int item, int& st, int& cn) {  // stage one (identity: nothing to fetch)
                    if (f.identity) { st = item; cn = 1; return; }
                    st = __ldg(f.indptr + item);
                    cn = __ldg(f.indptr + item + 1) - st;
                    if (cn > 32) cn = -1;  // long rows take the unstaged gather
                };
                auto stage_list = [&](int item, int st, int cn, int slot) {
                    if (f.identity) {
                        if (lane == 0) fr_ring[slot * 32] = make_int2(item, __float_as_int(1.0f));
                        fr_commit();
                    } else {
                        fr_stage_list(f, st, cn, fr_ring, slot, lane);
                    }
                };
                int c0 = lfm_bounded(next_u32(), (uint32_t)n_items), c1 = lfm_bounded(next_u32(), (uint32_t)n_items);
                int ps_st, ps_cn, c0_st, c0_cn, c1_st, c1_cn;
                bounds(pos_id, ps_st, ps_cn);
                bounds(c0, c0_st, c0_cn);
                bounds(c1, c1_st, c1_cn);
                __syncwarp();
                stage_list(pos_id, ps_st, ps_cn, 2);   // the positive borrows slot 2 (free until candidate 2 is staged)
                stage_list(c0, c0_st, c0_cn, 0);
                fr_wait1();
                __syncwarp();
                if (ps_cn >= 0) gather_staged<KPL>(m.item, a.hot_slot_item, d, ps_cn, fr_ring, 2, p, fr_p, lane);
                else gather_b<KPL>(a.itf, m.item, a.hot_slot_item, d, pos_id, p, fr_p, lane);
                pp = dot<KPL>(u, p);
                int sampled = 0;
                int cur = c0, cur_cn = c0_cn, nxt = c1, nxt_st = c1_st, nxt_cn = c1_cn;
                while (sampled < m.max_sampled) {
                    sampled++;
                    const int slot = (sampled - 1) % FR_RING;
                    // stages one / two of the next candidates, under this candidate's rows
                    const int nn = lfm_bounded(next_u32(), (uint32_t)n_items);
                    int nn_st, nn_cn;
                    bounds(nn, nn_st, nn_cn);
                    __syncwarp();
                    stage_list(nxt, nxt_st, nxt_cn, sampled % FR_RING);
                    fr_wait1();
                    __syncwarp();
                    const int cand = cur;
                    if (cur_cn >= 0) gather_staged<KPL>(m.item, a.hot_slot_item, d, cur_cn, fr_ring, slot, q, fr_q, lane);
                    else gather_b<KPL>(a.itf, m.item, a.hot_slot_item, d, cand, q, fr_q, lane);
                    const float np = dot<KPL>(u, q);
                    c_neg++;
                    cur = nxt; cur_cn = nxt_cn;
                    nxt = nn; nxt_st = nn_st; nxt_cn = nn_cn;
                    if (np > pp - 1.0f) {
                        if (lfm_warp_member(a.pos.indices, ps, pe, cand, lane)) { c_rej++; continue; }
                        loss = fminf(tp.weight * __ldg(a.loss_table_f + sampled), (float)LFM_MAX_LOSS);
                        neg_id = cand;
                        updated = true;
                        break;
                    }
                }
                fr_wait0();  // nothing of this interaction's staging may land in the ring later
                __syncwarp();
            } else if (LOSS == LOSS_BPR) {
                int tries = 0;
                do {  // T:1123-1127: popularity-weighted draw from the interaction list
                    int64_t j = (int64_t)(((unsigned long long)next_u32() * (unsigned long long)a.n_all) >> 32);
                    neg_id = __ldg(a.item_ids + j);
                    c_neg++;
                    tries++;
                    if (!lfm_warp_member(a.pos.indices, ps, pe, neg_id, lane)) break;
                    c_rej++;
                } while (tries < 256);
                GATHER(a.itf, m.item, a.hot_slot_item, neg_id, item_scale, q, fr_q);
                float np = dot<KPL>(u, q);
                loss = tp.weight * (1.0f - 1.0f / (1.0f + __expf(-(pp - np))));
                updated = true;
            } else {
                int sampled = 0;
                while (sampled < m.max_sampled) {
                    sampled++;
                    int cand = lfm_bounded(next_u32(), (uint32_t)n_items);
                    // with a resident plan's positives bitmap the candidate's membership word leaves
                    // before its gather: a violating candidate then needs no search afterwards
                    uint32_t wm = 0u;
                    if (a.pos_bitmap) wm = __ldg(a.pos_bitmap + (size_t)user * a.bitmap_words + (cand >> 5));
                    GATHER(a.itf, m.item, a.hot_slot_item, cand, item_scale, q, fr_q);
                    float np = dot<KPL>(u, q);
                    c_neg++;
                    if (np > pp - 1.0f) {
                        const bool member = a.pos_bitmap ? ((wm >> (cand & 31)) & 1u) != 0u
                                                         : lfm_warp_member(a.pos.indices, ps, pe, cand, lane);
                        if (member) { c_rej++; continue; }
                        float l = __ldg(a.loss_table_f + sampled);
                        loss = (LOSS == LOSS_KOS) ? l : tp.weight * l;
                        if (loss > (float)LFM_MAX_LOSS) loss = (float)LFM_MAX_LOSS;
                        neg_id = cand;
                        updated = true;
                        break;
                    }
                }
            }
            if (updated) {
                float gp[KPL], gn[KPL], gu[KPL];
#pragma unroll
                for (int k = 0; k < KPL; k++) {
                    gp[k] = -loss * u.v[k];
                    gn[k] = loss * u.v[k];
                    gu[k] = loss * (q.v[k] - p.v[k]);
                }
                int n1, n2, n3;
                SCATTER(a.itf, m.item, a.hot_slot_item, pos_id, fr_p, gp, -loss, alpha_i, n1);
                SCATTER(a.itf, m.item, a.hot_slot_item, neg_id, fr_q, gn, loss, alpha_i, n2);
                SCATTER(a.usf, m.user, a.hot_slot_user, user, fr_u, gu, loss, alpha_u, n3);
                nnz_total = n1 + n2 + n3;
                c_upd++;
            }
        }

        if (REG && updated) {
            // T:528-534 / T:640-649 in the log domain; flushed to the global scale every 16 tuples.
            float avg = lfm_warp_sum(lrsum) / (float)((d + 1) * nnz_total);
            rs.loc_i += log1p(a.item_alpha * (double)avg);
            rs.loc_u += log1p(a.user_alpha * (double)avg);
            if (++rs.pending == 16) {
                if (lane == 0) {
                    atomicAdd(&a.scales->item_scale, rs.loc_i);
                    atomicAdd(&a.scales->user_scale, rs.loc_u);
                }
                rs.base_i = __ldcg(&a.scales->item_scale);
                rs.base_u = __ldcg(&a.scales->user_scale);
                rs.loc_i = rs.loc_u = 0.0;
                rs.pending = 0;
            }
        }
    }
#undef GATHER
#undef SCATTER
    if constexpr (HOT) {  // no more updates after the barrier: drain everything that is still pending
        __syncthreads();
        for (int slot = threadIdx.x >> 5; slot < a.n_hot; slot += blockDim.x >> 5)
            hot_drain<KPL / 4>(hsm, slot, a, lane);
    }
    if (REG && rs.pending && lane == 0) {
        atomicAdd(&a.scales->item_scale, rs.loc_i);
        atomicAdd(&a.scales->user_scale, rs.loc_u);
    }
    if (lane == 0) {
        atomicAdd(&a.counters->positives, c_pos);
        atomicAdd(&a.counters->negatives, c_neg);
        atomicAdd(&a.counters->updates, c_upd);
        atomicAdd(&a.counters->rejected, c_rej);
    }
}

__global__ void delta_kernel(int mode, DeltaSegs segs, float* __restrict__ S, float* __restrict__ D) {
    const int64_t total = segs.n[0] + segs.n[1] + segs.n[2] + segs.n[3];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int64_t j = i;
        int s = 0;
        while (j >= segs.n[s]) { j -= segs.n[s]; s++; }
        float* cur = segs.p[s] + j;
        if (mode == 0) S[i] = *cur;
        else if (mode == 1) { const float dlt = *cur - S[i]; D[i] = dlt; S[i] = dlt; }
        else *cur = *cur + (D[i] - S[i]);
    }
}

// Hogwild-mode regularize: scales live in the log domain.
__global__ void regularize_log_kernel(DevModel m, DevScales* scales) {
    float is = (float)exp(scales->item_scale), us = (float)exp(scales->user_scale);
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t ni = (size_t)m.item.n * m.d, nu = (size_t)m.user.n * m.d;
    for (size_t i = tid; i < ni; i += stride) m.item.w[i] = m.item.w[i] / is;
    for (size_t i = tid; i < (size_t)m.item.n; i += stride) m.item.b[i] = m.item.b[i] / is;
    for (size_t i = tid; i < nu; i += stride) m.user.w[i] = m.user.w[i] / us;
    for (size_t i = tid; i < (size_t)m.user.n; i += stride) m.user.b[i] = m.user.b[i] / us;
}
__global__ void reset_scales_kernel(DevScales* s) { s->item_scale = 0.0; s->user_scale = 0.0; }

}  // namespace
// Hogwild relies on collisions being rare: the interactions in flight at any moment must be a
// small fraction of the epoch, or every update is computed from a state that is a whole epoch
// stale (the reference's OpenMP loop has num_threads <= ~100 in flight).  The grid is therefore
// capped at max(64, n / divisor) concurrent interactions; a full B200 wave (~4.7k warps) is
// reached from ~600k interactions per launch.
static std::atomic<int> g_inflight_divisor{128};
static std::atomic<int> g_hot_enabled{1};
extern "C" int lfm_set_hot_rows(int enabled) { return g_hot_enabled.exchange(enabled ? 1 : 0); }
extern "C" int lfm_set_inflight_divisor(int divisor) {
    int old = g_inflight_divisor.load();
    if (divisor >= 1) g_inflight_divisor.store(divisor);
    return old;
}
static int64_t lfm_inflight_cap(int64_t count) {
    int64_t cap = count / g_inflight_divisor.load();
    return cap < 64 ? 64 : cap;
}
namespace {

template <int LOSS, int KPL>
cudaError_t launch_generic(const FitArgs& a, const Tuple* tuples, int64_t begin, int64_t count,
                           cudaStream_t st) {
    FitArgs b = a;
    b.n = count;
    b.t_offset = begin;
    const Tuple* tp = tuples + begin;
    bool reg = (a.item_alpha != 0.0 || a.user_alpha != 0.0);
    int block = 256;
    int64_t warps_needed = count;
    int64_t blocks = (warps_needed * 32 + block - 1) / block;
    int64_t cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    int64_t fl = (lfm_inflight_cap(count) + 7) / 8;
    if (blocks > fl) blocks = fl;
    if (blocks < 1) blocks = 1;
    const DevModel& m = a.model;
    const bool vec = KPL % 4 == 0 && !m.adadelta && !reg && m.d % 4 == 0 &&
                     (((uintptr_t)m.item.w | (uintptr_t)m.item.g | (uintptr_t)m.user.w | (uintptr_t)m.user.g) & 15) == 0;
    if (m.adadelta) {
        if (reg) hogwild_kernel<LOSS, KPL, 1, true, true><<<(int)blocks, block, 0, st>>>(b, tp);
        else hogwild_kernel<LOSS, KPL, 1, true, false><<<(int)blocks, block, 0, st>>>(b, tp);
    } else if (reg) {
        hogwild_kernel<LOSS, KPL, 1, false, true><<<(int)blocks, block, 0, st>>>(b, tp);
    } else if (vec) {
        if constexpr (KPL == 4) {
            if (a.n_hot > 0 && g_hot_enabled.load()) {
                // hot-row variant: 512-thread CTAs, one per SM (the accumulators take most of the
                // shared memory), persistent warps
                auto kern = hogwild_kernel<LOSS, 4, 4, false, false, true>;
                const size_t smem = (((size_t)a.n_hot * ((2 * (m.d >> 2) + 1) * sizeof(float4) + sizeof(int)) + 15) & ~(size_t)15) +
                                    (size_t)16 * FR_RING * 32 * sizeof(int2);  // accumulators + locks, then 16 warps' candidate rings
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                int per_sm = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 512, smem);
                if (per_sm >= 1) {
                    int dev = 0, sms = 148;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                    int64_t hb = (int64_t)sms * per_sm, need = (count + 15) / 16, capb = (lfm_inflight_cap(count) + 15) / 16;
                    if (hb > need) hb = need;
                    if (hb > capb) hb = capb;
                    if (hb < 1) hb = 1;
                    kern<<<(int)hb, 512, smem, st>>>(b, tp);
                    return cudaGetLastError();
                }
            }
        }
        if constexpr (KPL % 4 == 0) hogwild_kernel<LOSS, KPL, 4, false, false><<<(int)blocks, block, 0, st>>>(b, tp);
    } else {
        hogwild_kernel<LOSS, KPL, 1, false, false><<<(int)blocks, block, 0, st>>>(b, tp);
    }
    return cudaGetLastError();
}

template <int LOSS>
cudaError_t launch_generic_kpl(const FitArgs& a, const Tuple* tuples, int64_t begin, int64_t count,
                               cudaStream_t st) {
    int d = a.model.d;
    const bool reg = (a.item_alpha != 0.0 || a.user_alpha != 0.0);
    // Vector layout (one float4 chunk per lane per 128 components) when the update is a pure add
    // and the rows are private to an entity (identity features).  With shared feature rows the
    // epoch is bound by same-address reductions serialising in L2 on the hot rows, where
    // red.v4 measured slower than scalar reds (C3: 414 ms vs 367 ms per epoch).
    if (d % 4 == 0 && !a.model.adadelta && !reg && a.itf.identity && a.usf.identity) {
        if (d <= 128) return launch_generic<LOSS, 4>(a, tuples, begin, count, st);
        if (d <= 256) return launch_generic<LOSS, 8>(a, tuples, begin, count, st);
    }
    if (d <= 32) return launch_generic<LOSS, 1>(a, tuples, begin, count, st);
    if (d <= 64) return launch_generic<LOSS, 2>(a, tuples, begin, count, st);
    if (d <= 128) return launch_generic<LOSS, 4>(a, tuples, begin, count, st);
    if (d <= 256) return launch_generic<LOSS, 8>(a, tuples, begin, count, st);
    return cudaErrorInvalidValue;
}

}  // namespace

#include "lfm_hogwild_fast.cuh"

namespace {
__global__ void feature_count_kernel(FitArgs a, int kos, float* cnt_item, float* cnt_user) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < a.n; e += stride) {
        const int user = a.user_ids[e];
        // k-OS draws its positive from the user's row of the CSR: the same multiset of items
        const int item = kos ? a.pos.indices[e] : a.item_ids[e];
        if (cnt_item && !a.itf.identity)
            for (int i = a.itf.indptr[item]; i < a.itf.indptr[item + 1]; i++) atomicAdd(cnt_item + a.itf.indices[i], 1.0f);
        if (cnt_user && !a.usf.identity)
            for (int i = a.usf.indptr[user]; i < a.usf.indptr[user + 1]; i++) atomicAdd(cnt_user + a.usf.indices[i], 1.0f);
    }
}
__global__ void feature_count_items_kernel(DevCsr itf, float per_item, float* cnt_item) {
    int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= itf.rows) return;
    for (int i = itf.indptr[item]; i < itf.indptr[item + 1]; i++) atomicAdd(cnt_item + itf.indices[i], per_item);
}
}  // namespace

cudaError_t lfm_launch_feature_counts(const FitArgs& a, int loss, float* cnt_item, float* cnt_user,
                                      float neg_per_item, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    if (cnt_item) e = cudaMemsetAsync(cnt_item, 0, sizeof(float) * (size_t)a.model.item.n, st);
    if (e == cudaSuccess && cnt_user) e = cudaMemsetAsync(cnt_user, 0, sizeof(float) * (size_t)a.model.user.n, st);
    if (e != cudaSuccess || a.n == 0) return e;
    int64_t blocks = (a.n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    feature_count_kernel<<<(int)blocks, 256, 0, st>>>(a, loss == LOSS_KOS ? 1 : 0, cnt_item, cnt_user);
    if (cnt_item && !a.itf.identity && loss != LOSS_LOGISTIC && neg_per_item > 0)
        feature_count_items_kernel<<<(a.itf.rows + 255) / 256, 256, 0, st>>>(a.itf, neg_per_item, cnt_item);
    return cudaGetLastError();
}

cudaError_t lfm_launch_check_unit(const float* y, const float* w, int64_t n, int32_t* flag, cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    check_unit_kernel<<<(int)blocks, 256, 0, st>>>(y, w, n, flag);
    return cudaGetLastError();
}

cudaError_t lfm_launch_build_bitmap_coo(const int32_t* user_ids, const int32_t* item_ids, int64_t n,
                                        uint32_t* bitmap, int32_t rows, int32_t words_per_row, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * (size_t)rows * words_per_row, st);
    if (e != cudaSuccess || n == 0) return e;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    build_bitmap_coo_kernel<<<(int)blocks, 256, 0, st>>>(user_ids, item_ids, n, bitmap, words_per_row);
    return cudaGetLastError();
}

cudaError_t lfm_launch_build_bitmap(const DevCsr& pos, uint32_t* bitmap, int32_t words_per_row, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * (size_t)pos.rows * words_per_row, st);
    if (e != cudaSuccess || pos.nnz == 0) return e;
    int64_t blocks = (pos.nnz + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    build_bitmap_kernel<<<(int)blocks, 256, 0, st>>>(pos, bitmap, words_per_row);
    return cudaGetLastError();
}

cudaError_t lfm_launch_delta(int mode, const DeltaSegs& segs, float* S, float* D, cudaStream_t st) {
    const int64_t total = segs.n[0] + segs.n[1] + segs.n[2] + segs.n[3];
    if (total == 0) return cudaSuccess;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    delta_kernel<<<(int)blocks, 256, 0, st>>>(mode, segs, S, D);
    return cudaGetLastError();
}

// Host-visible helper: is (loss, model, features) eligible for the hogwild path at all?
int lfm_hogwild_supported(int loss, int d, int nkos) {
    if (d < 1 || d > 256) return 0;
    if (loss == LOSS_KOS && nkos > 32) return 0;
    return 1;
}

cudaError_t lfm_launch_pack(const FitArgs& a, int loss, Tuple* tuples, uint32_t perm_key,
                            cudaStream_t st) {
    if (a.n == 0) return cudaSuccess;
    int bits = 1;
    while (((int64_t)1 << bits) < a.n) bits++;
    int hb = (bits + 1) / 2;
    if (hb < 1) hb = 1;
    int64_t blocks = (a.n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    int skip = (loss == LOSS_WARP || loss == LOSS_BPR) ? 1 : 0;
    pack_kernel<<<(int)blocks, 256, 0, st>>>(a.user_ids, loss == LOSS_KOS ? nullptr : a.item_ids,
                                             loss == LOSS_KOS ? nullptr : a.y,
                                             loss == LOSS_KOS ? nullptr : a.sample_weight, a.shuffle,
                                             a.n, hb, perm_key, skip, a.unit_weights, a.row_offset, tuples);
    return cudaGetLastError();
}

// Runs one epoch in hogwild mode.  With L2 regularisation the epoch is cut into
// segments so the host-visible rescale check (T:901-904) happens between launches.
cudaError_t lfm_launch_hogwild(int loss, const FitArgs& a, Tuple* tuples, cudaStream_t st,
                               int* launches, cudaEvent_t ev_train_begin, cudaEvent_t ev_train_end,
                               cudaStream_t pack_stream, cudaEvent_t pack_after, cudaEvent_t pack_done,
                               bool prepacked) {
    cudaError_t e;
    if (prepacked) {
        // `tuples` was packed for this epoch's seed on the side stream while the previous epoch trained
        // (lfm_plan_epoch_next): wait for that kernel, launch nothing
        e = cudaStreamWaitEvent(st, pack_done, 0);
        if (launches) (*launches)--;  // (counted by the epoch that launched it)
    } else if (pack_stream) {
        // pack_kernel only reads the interaction list: it runs on a side stream under whatever the main
        // stream is still doing (the model state's H2D copies), after `pack_after`
        e = cudaStreamWaitEvent(pack_stream, pack_after, 0);
        if (e != cudaSuccess) return e;
        e = lfm_launch_pack(a, loss, tuples, a.seed ^ 0x5bd1e995u, pack_stream);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(pack_done, pack_stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(st, pack_done, 0);
    } else {
        e = lfm_launch_pack(a, loss, tuples, a.seed ^ 0x5bd1e995u, st);
    }
    if (e != cudaSuccess) return e;
    if (launches) (*launches)++;
    bool reg = (a.item_alpha != 0.0 || a.user_alpha != 0.0);
    if (reg) { reset_scales_kernel<<<1, 1, 0, st>>>(a.scales); if (launches) (*launches)++; }
    if (ev_train_begin) cudaEventRecord(ev_train_begin, st);
    if (a.n == 0) {
        if (ev_train_end) cudaEventRecord(ev_train_end, st);
        return cudaSuccess;
    }
    // With L2 the lazy scale grows by at most log1p(alpha * llr) per update (llr <= lr under adagrad,
    // whose accumulators start at 1): keep a segment short enough that the log-scale cannot pass
    // ln(1e6) before the check between launches (the reference rescales right after the update
    // that crosses 1e6, T:901-904).
    int64_t seg = a.n;
    if (reg) {
        const double amax = a.item_alpha > a.user_alpha ? a.item_alpha : a.user_alpha;
        const double step = amax * (a.model.adadelta ? 1.0 : (double)a.model.lr);
        double lim = step > 0 ? 13.8 / step : 2097152.0;
        if (lim > 2097152.0) lim = 2097152.0;
        if (lim < 16384.0) lim = 16384.0;
        seg = (int64_t)lim;
    }
    for (int64_t begin = 0; begin < a.n; begin += seg) {
        int64_t count = (a.n - begin < seg) ? (a.n - begin) : seg;
        bool done = false;
        e = lfm_try_launch_fast(loss, a, tuples, begin, count, st, &done);
        if (e != cudaSuccess) return e;
        if (!done) {
            // bitmap-only plans carry no positives CSR for the generic kernels to search
            if (loss != LOSS_LOGISTIC && a.pos.indptr == nullptr) return cudaErrorInvalidValue;
            switch (loss) {
                case LOSS_LOGISTIC: e = launch_generic_kpl<LOSS_LOGISTIC>(a, tuples, begin, count, st); break;
                case LOSS_WARP: e = launch_generic_kpl<LOSS_WARP>(a, tuples, begin, count, st); break;
                case LOSS_BPR: e = launch_generic_kpl<LOSS_BPR>(a, tuples, begin, count, st); break;
                case LOSS_KOS: e = launch_generic_kpl<LOSS_KOS>(a, tuples, begin, count, st); break;
                default: return cudaErrorInvalidValue;
            }
            if (e != cudaSuccess) return e;
        }
        if (launches) (*launches)++;
        if (reg && begin + seg < a.n) {
            // mid-epoch rescale (T:901-904): check the scales between segments
            DevScales h;
            e = cudaMemcpyAsync(&h, a.scales, sizeof(h), cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            if (h.item_scale > 13.815510557964274 || h.user_scale > 13.815510557964274) {  // ln(1e6)
                regularize_log_kernel<<<148 * 8, 256, 0, st>>>(a.model, a.scales);
                reset_scales_kernel<<<1, 1, 0, st>>>(a.scales);
                if (launches) (*launches) += 2;
            }
        }
    }
    if (ev_train_end) cudaEventRecord(ev_train_end, st);
    if (reg) {
        regularize_log_kernel<<<148 * 8, 256, 0, st>>>(a.model, a.scales);
        if (launches) (*launches)++;
    }
    return cudaGetLastError();
}
