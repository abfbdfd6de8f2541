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

}  // namespace rfk
