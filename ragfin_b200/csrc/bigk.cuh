// Large-k path (k above the fused-list limit, up to Milvus' 16384): exact from the start.
//   score_all_kernel   canonical fp64 score of every row for one query -> fp32 scores [n]
//   radix_hist_kernel  + radix_pick_kernel: 8 passes of 8 bits over the 64-bit keys
//                      (ordered(score) << 32 | ~row) find the k-th largest key exactly (keys are unique)
//   compact_kernel     keys >= threshold -> exactly k keys
//   rank_sort_kernel   rank of each key by counting (O(k^2), tiled through shared memory) -> ordered output
// Used by graph_cons.py:279-style calls (limit = 1000) on corpora of more than 256 rows.
#pragma once
#include "common.cuh"

namespace rfk {

template <int DT>
__global__ void __launch_bounds__(256) score_all_kernel(const void* __restrict__ data, int64_t n_rows, int ld,
                                                        const float* __restrict__ qhat, float* __restrict__ scores,
                                                        const uint32_t* __restrict__ allow) {
    const int lane = threadIdx.x & 31;
    const typename Store<DT>::T* base = reinterpret_cast<const typename Store<DT>::T*>(data);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += stride) {
        if (allow != nullptr && !row_allowed(allow, r)) {   // filtered out: below every real score
            if (lane == 0) scores[r] = -INFINITY;
            continue;
        }
        const double s = canonical_dot_row<DT>(base + (size_t)r * ld, qhat, ld, lane);
        if (lane == 0) scores[r] = (float)s + 0.0f;
    }
}

struct RadixState {
    u64 prefix;        // bits of the k-th key decided so far (high to low)
    int64_t remaining; // rank still to resolve inside the current prefix
    unsigned int hist[256];
    unsigned int out_count;
};

// histogram of byte `shift/8` over the keys whose higher bytes equal the prefix
__global__ void __launch_bounds__(256) radix_hist_kernel(const float* __restrict__ scores, int64_t n, int shift,
                                                         RadixState* st) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u64 prefix = st->prefix;
    const u64 himask = shift >= 56 ? 0ull : ~0ull << (shift + 8);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const u64 key = make_key(scores[i], (uint32_t)i);
        if ((key & himask) == (prefix & himask)) atomicAdd(&h[(key >> shift) & 255], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

// pick the byte value that contains the wanted rank, descend
__global__ void radix_pick_kernel(int shift, RadixState* st) {
    if (threadIdx.x != 0) return;
    int64_t rem = st->remaining;
    int b = 255;
    for (; b > 0; --b) {
        const int64_t c = st->hist[b];
        if (rem <= c) break;
        rem -= c;
    }
    st->prefix |= (u64)b << shift;
    st->remaining = rem;
    for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

__global__ void __launch_bounds__(256) compact_kernel(const float* __restrict__ scores, int64_t n, RadixState* st,
                                                      u64* __restrict__ keys, int cap) {
    const u64 thr = st->prefix;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const u64 key = make_key(scores[i], (uint32_t)i);
        if (key >= thr) {
            const unsigned pos = atomicAdd(&st->out_count, 1u);
            if ((int)pos < cap) keys[pos] = key;
        }
    }
}

// out position of key i = number of keys greater than it
__global__ void __launch_bounds__(256) rank_sort_kernel(const u64* __restrict__ keys, int m, int k, int64_t id_base,
                                                        int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    __shared__ u64 tile[1024];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 mine = i < m ? keys[i] : 0ull;
    int rank = 0;
    for (int t0 = 0; t0 < m; t0 += 1024) {
        for (int j = threadIdx.x; j < 1024; j += blockDim.x) tile[j] = t0 + j < m ? keys[t0 + j] : 0ull;
        __syncthreads();
        const int lim = m - t0 < 1024 ? m - t0 : 1024;
        for (int j = 0; j < lim; ++j) rank += tile[j] > mine;
        __syncthreads();
    }
    if (i < m && rank < k) {
        out_ids[rank] = id_base + (int64_t)key_row(mine);
        out_scores[rank] = key_score(mine);
    }
    if (i >= m && i < k) {   // slots past the number of rows
        out_ids[i] = -1;
        out_scores[i] = -INFINITY;
    }
}

// ---------------------------------------------------------------------------------
// Batched large-k pipeline (k > 256, up to 16384; graph_cons.py:279 asks limit = 1000): no host synchronisation per query.
//   1. the tensor-core sweep dumps approximate scores of up to 16 queries at once: approx[q][row]   (gemm.cuh MODE 1)
//   2. bk_hist_kernel x 4   radix select of T = the keff-th largest approximate score of every query (allowed rows only);
//                           pass p derives its prefix from the histogram of pass p - 1, so no pick kernels in between
//   3. bk_compact_kernel    rows with approx >= T - 2 eps  ->  candidate rows (every row of the exact top-k is among them:
//                           DESIGN.md 2.4, with the whole approximate score vector at hand instead of an append buffer)
//   4. bk_rescore_kernel    canonical fp64 score of every candidate  ->  exact keys
//   5. bk_rank_sort_kernel  the keff largest exact keys in order
// A query whose candidates exceed the buffer (thousands of rows within 2 eps of the k-th score) is flagged and answered by
// the exact one-query path above.
// ---------------------------------------------------------------------------------
struct BkState {                  // per query
    uint32_t prefix[4];           // ordered-score bits decided after pass p (high to low)
    int32_t remaining[4];         // rank still to resolve inside prefix[p]
    uint32_t hist[4][256];
    uint32_t n_cand;              // rows compacted (may exceed the buffer: overflow)
    uint32_t overflow;
    uint32_t pad[2];
};

// (prefix, remaining) after pass p from the state of pass p - 1 and hist[p]: the bin, from the top, where the cumulative count
// reaches `remaining`.  Executed by one warp; every block computes the same values.
__device__ __forceinline__ void bk_resolve(const BkState* st, int p, int want0, uint32_t* prefix_out, int* remaining_out, int lane) {
    uint32_t prefix = 0u;
    int want = want0;
    for (int pp = 0; pp <= p; ++pp) {
        if (pp < p) { prefix = st->prefix[pp]; want = st->remaining[pp]; continue; }   // stored by block 0 of the pass that computed it
        const int shift = 24 - 8 * pp;
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { c[b] = st->hist[pp][255 - 8 * lane - b]; sum += c[b]; }
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(kFull, incl, off);
            if (lane >= off) incl += o;
        }
        const uint32_t before = incl - sum;
        uint32_t np = 0u;
        int nw = 0;
        if ((int)before < want && (int)incl >= want) {
            uint32_t run = before;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if ((int)run < want && (int)(run + c[b]) >= want) { np = prefix | ((uint32_t)(255 - 8 * lane - b) << shift); nw = want - (int)run; }
                run += c[b];
            }
        }
        const unsigned who = __ballot_sync(kFull, nw > 0);
        if (who) { const int src = __ffs(who) - 1; np = __shfl_sync(kFull, np, src); nw = __shfl_sync(kFull, nw, src); }
        prefix = np; want = nw;
    }
    *prefix_out = prefix;
    *remaining_out = want;
}

// pass p (0..3): histogram of byte (3 - p) of the ordered scores whose higher bytes equal the prefix decided so far
__global__ void __launch_bounds__(256) bk_hist_kernel(const float* __restrict__ approx, int64_t n, int pass, int keff,
                                                      BkState* __restrict__ states, const uint32_t* __restrict__ allow) {
    __shared__ uint32_t h[256];
    __shared__ uint32_t s_prefix;
    const int q = blockIdx.y, lane = threadIdx.x & 31;
    BkState* st = states + q;
    h[threadIdx.x] = 0u;
    if (pass > 0 && threadIdx.x < 32) {
        uint32_t prefix; int rem;
        bk_resolve(st, pass - 1, keff, &prefix, &rem, lane);
        if (lane == 0) {
            s_prefix = prefix;
            if (blockIdx.x == 0) { st->prefix[pass - 1] = prefix; st->remaining[pass - 1] = rem; }
        }
    }
    __syncthreads();
    const uint32_t prefix = pass > 0 ? s_prefix : 0u;
    const int shift = 24 - 8 * pass;
    const float* sc = approx + (size_t)q * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (allow != nullptr && !row_allowed(allow, i)) continue;
        const uint32_t o = float_to_ordered(sc[i] + 0.0f);
        if (pass == 0 || (o >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&h[(o >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[pass][threadIdx.x], h[threadIdx.x]);
}

// rows whose approximate score reaches T - 2 eps -> cand_rows[q][0 .. n_cand)
__global__ void __launch_bounds__(256) bk_compact_kernel(const float* __restrict__ approx, int64_t n, int keff, float eps_const,
                                                         const float* __restrict__ eps_q, BkState* __restrict__ states,
                                                         uint32_t* __restrict__ cand_rows, int cmax, const uint32_t* __restrict__ allow) {
    __shared__ float s_cut;
    const int q = blockIdx.y, lane = threadIdx.x & 31;
    BkState* st = states + q;
    if (threadIdx.x < 32) {
        uint32_t t_ord; int rem;
        bk_resolve(st, 3, keff, &t_ord, &rem, lane);
        if (lane == 0) {
            const float e = eps_const + (eps_q ? eps_q[q] : 0.0f);
            s_cut = keff > 0 ? __fsub_rd(__fsub_rd(ordered_to_float(t_ord), __fmul_ru(2.0f, e)), 2.384185791015625e-07f) : INFINITY;
        }
    }
    __syncthreads();
    const float cut = s_cut;
    const float* sc = approx + (size_t)q * n;
    uint32_t* out = cand_rows + (size_t)q * cmax;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (allow != nullptr && !row_allowed(allow, i)) continue;
        if (sc[i] >= cut) {
            const uint32_t pos = atomicAdd(&st->n_cand, 1u);
            if (pos < (uint32_t)cmax) out[pos] = (uint32_t)i;
            else st->overflow = 1u;
        }
    }
}

// one warp per candidate: canonical fp64 score -> exact key (0 past the candidate count)
template <int DT>
__global__ void __launch_bounds__(256) bk_rescore_kernel(const void* __restrict__ data, int ld, const float* __restrict__ qhat,
                                                         const BkState* __restrict__ states, const uint32_t* __restrict__ cand_rows,
                                                         int cmax, u64* __restrict__ keys) {
    const int q = blockIdx.y, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= cmax) return;
    const uint32_t nc = states[q].n_cand < (uint32_t)cmax ? states[q].n_cand : (uint32_t)cmax;
    u64 key = 0ull;
    if ((uint32_t)c < nc) {
        const uint32_t row = cand_rows[(size_t)q * cmax + c];
        const double s = canonical_dot_row<DT>(reinterpret_cast<const typename Store<DT>::T*>(data) + (size_t)row * ld, qhat + (size_t)q * ld, ld, lane);
        key = make_key((float)s + 0.0f, row);
    }
    if (lane == 0) keys[(size_t)q * cmax + c] = key;
}

// out position of key i = number of keys greater than it (grid.y = query); flagged queries are left to the fallback
__global__ void __launch_bounds__(256) bk_rank_sort_kernel(const u64* __restrict__ keys_all, const BkState* __restrict__ states, int cmax, int k,
                                                           int keff, int64_t id_base, int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    __shared__ u64 tile[1024];
    const int q = blockIdx.y;
    const int m = (int)(states[q].n_cand < (uint32_t)cmax ? states[q].n_cand : (uint32_t)cmax);
    const u64* keys = keys_all + (size_t)q * cmax;
    int64_t* oid = out_ids + (size_t)q * k;
    float* osc = out_scores + (size_t)q * k;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 mine = i < m ? keys[i] : 0ull;
    int rank = 0;
    for (int t0 = 0; t0 < m; t0 += 1024) {
        for (int j = threadIdx.x; j < 1024; j += blockDim.x) tile[j] = t0 + j < m ? keys[t0 + j] : 0ull;
        __syncthreads();
        const int lim = m - t0 < 1024 ? m - t0 : 1024;
        for (int j = 0; j < lim; ++j) rank += tile[j] > mine;
        __syncthreads();
    }
    if (i < m && rank < keff) {
        oid[rank] = id_base + (int64_t)key_row(mine);
        osc[rank] = key_score(mine);
    }
    if (i >= keff && i < k) {   // slots past the number of rows that exist
        oid[i] = -1;
        osc[i] = -INFINITY;
    }
}

}  // namespace rfk
