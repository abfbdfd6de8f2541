// K3s (EXPERIMENTAL, opt-in gemm variant 5): the <= 16-query sweep of gemm_rows.cuh with the bound pass INSIDE it.
//
// STATUS: compiles for sm_100a but HAS NOT RUN ON A GPU YET - written after the round's GPU budget was spent.  Never
// selected automatically; scripts/seeded_check.py is the first thing to run (parity against the oracle, then timing
// against variant 3), in a child process under a timeout (the kernel holds a grid-wide barrier).
//
// Why: a batch-1 call is  prep -> bound sweep (MODE 2, ~20 us) -> bound_select (~6 us) -> sweep -> finalize -> 2 gated
// tier-2 launches.  The two bound-pass launches are pure latency: ~1 % of a 10M-row call but ~8 % of the 0.34 ms step
// on the 1.25M-row shard of the 8-GPU split.  Here the sweep seeds itself:
//   phase A   every CTA first scores `sample_tiles` (1-4) tiles at the head of its first slice and writes, per query, the
//             maximum over each 128-row half tile to bmax[q][block] (block = (CTA, tile, half): 2 * sample_tiles * grid
//             blocks of DISTINCT rows, evenly spread over the corpus because the slices are)
//   barrier   one release-add on a grid-wide counter per CTA, then a spin until all CTAs have arrived (cooperative
//             launch: all CTAs are co-resident; the counter only ever grows, the host passes the value to wait for)
//   select    every CTA computes, redundantly, thr[q] = (k-th largest block maximum) - 2 eps - 2^-22 for its <= 16
//             queries (one warp per query, values staged in shared memory), exactly bound_select_kernel's rule: k
//             distinct rows reach that value, so it bounds the k-th best approximate score from below
//   phase B   the normal sweep from the first tile of the slice (the sample tiles are scored again: they were scored
//             before the threshold existed), appending rows with score >= thr[q]
// The TMA producer and the MMA issuer do not know about the phases: they stream [sample tiles] + [all tiles]; the
// epilogue's back-pressure (two accumulators, four ring stages) holds them at the barrier.
// Host side: k <= 16 (the warp select walks at most k distinct values), cnt[] zeroed by a memset before the launch.
#pragma once
#include "gemm_rows.cuh"

namespace rfk {

constexpr int kSeedMaxK = 16;
constexpr int kSeedMaxSampleTiles = 4;

struct SeededArgs {
    RowsArgs r;                 // r.thr is unused: the thresholds are computed in the kernel
    int sample_tiles;           // tiles each CTA samples at the head of its first slice (1..kSeedMaxSampleTiles)
    int rank;                   // k
    float eps_const;            // accumulation error bound (eps_gemm_const)
    const float* eps_q;         // [nq] query rounding error bound
    float* bmax;                // [nq][nblk] block maxima, nblk = gridDim.x * 2 * sample_tiles
    uint32_t* gbar;             // grid-wide arrival counter (monotonic across launches)
    uint32_t gbar_target;       // value of *gbar once every CTA of THIS launch has arrived
};

__host__ __device__ constexpr size_t seeded_extra_smem(int grid, int sample_tiles) {
    return 2 * 4 * kRN * sizeof(float) /*wmax*/ + kRN * sizeof(float) /*thresholds*/ +
           (size_t)4 * grid * 2 * sample_tiles * sizeof(uint32_t) /*one select buffer per epilogue warp*/;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void epi_bar_sync() {   // the 128 epilogue threads only (named barrier 1)
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// want-th largest (1-based, duplicates counted) of the n ordered-uint values in v (shared memory), one warp.
// Every value is > 0 (float_to_ordered never returns 0 for a non-NaN float).  Walks at most `want` distinct values.
__device__ __forceinline__ uint32_t warp_kth_largest(const uint32_t* v, int n, int want, int lane) {
    uint32_t prev = 0u;
    bool first = true;
    int remaining = want;
    for (int it = 0; it < want; ++it) {
        uint32_t best = 0u;
        for (int i = lane; i < n; i += kWarp) {
            const uint32_t x = v[i];
            if ((first || x < prev) && x > best) best = x;
        }
        best = __reduce_max_sync(kFull, best);
        if (best == 0u) break;                       // fewer than `want` values
        int c = 0;
        for (int i = lane; i < n; i += kWarp) c += v[i] == best ? 1 : 0;
        c = __reduce_add_sync(kFull, c);
        if (c >= remaining) return best;
        remaining -= c;
        prev = best;
        first = false;
    }
    return 0x007FFFFFu;                              // -inf: collect everything (overflows to the exact tier; never wrong)
}

template <int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_rows_seeded_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmB, const SeededArgs s) {
    const RowsArgs& a = s.r;
    extern __shared__ uint8_t ssm_raw[];
    const uint32_t raw = smem_u32(ssm_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* ssm = ssm_raw + (base - raw);
    const int stages = a.stages, nkb = a.num_kblocks;
    const uint32_t smQ = base;                                         // [nkb][2 KB]
    const uint32_t smB = base + (uint32_t)nkb * kRQBytes;              // [stages][32 KB]
    uint8_t* tail = ssm + (size_t)nkb * kRQBytes + (size_t)stages * kBBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int st) { return bar0 + 8u * st; };
    auto empty_bar = [&](int st) { return bar0 + 8u * (kRMaxStages + st); };
    const uint32_t qfull_bar = bar0 + 8u * (2 * kRMaxStages);
    auto tfull_bar = [&](int st) { return bar0 + 8u * (2 * kRMaxStages + 1 + st); };
    auto tempty_bar = [&](int st) { return bar0 + 8u * (2 * kRMaxStages + 3 + st); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRMaxStages + 5);
    float* wmax = reinterpret_cast<float*>(tail + 256);                // [2 halves][4 warps][16 queries]
    float* thr_sh = wmax + 2 * 4 * kRN;                                // [16]
    uint32_t* selbuf = reinterpret_cast<uint32_t*>(thr_sh + kRN);      // [4 warps][nblk]
    const int nblk = (int)gridDim.x * 2 * s.sample_tiles;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int st = 0; st < stages; ++st) { mbar_init(full_bar(st), 1); mbar_init(empty_bar(st), 1); }
        mbar_init(qfull_bar, 1);
        for (int st = 0; st < 2; ++st) { mbar_init(tfull_bar(st), 1); mbar_init(tempty_bar(st), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    auto slice_tiles = [&](int sl, long long& r0, long long& r1) -> int {
        r0 = (long long)sl * a.rows_per_slice;
        r1 = r0 + a.rows_per_slice;
        if (r1 > a.n_rows) r1 = a.n_rows;
        return r1 > r0 ? (int)((r1 - r0 + kGN - 1) / kGN) : 0;
    };
    // sample tiles of this CTA: the first `na` tiles of its first slice (the host guarantees gridDim.x <= S)
    long long f0, f1;
    const int first_tiles = slice_tiles((int)blockIdx.x, f0, f1);
    const int na = first_tiles < s.sample_tiles ? first_tiles : s.sample_tiles;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer: [sample tiles] + [every tile of every slice] =====
            mbar_expect_tx(qfull_bar, (uint32_t)nkb * kRQBytes);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smQ + (uint32_t)kb * kRQBytes, &tmQ, kb * a.k_elems, 0, qfull_bar);
            int stage = 0;
            uint32_t phase = 0;
            auto produce_tile = [&](long long row0) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), kBBytes);
                    tma_load_2d(smB + (uint32_t)stage * kBBytes, &tmB, kb * a.k_elems, (int)row0, full_bar(stage));
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            };
            for (int t = 0; t < na; ++t) produce_tile(f0 + (long long)t * kGN);
            for (int sl = blockIdx.x; sl < a.S; sl += gridDim.x) {
                long long r0, r1;
                const int ntiles = slice_tiles(sl, r0, r1);
                for (int t = 0; t < ntiles; ++t) produce_tile(r0 + (long long)t * kGN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer: same tile sequence =====
            mbar_wait(qfull_bar, 0u);
            tc_fence_after();
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            auto mma_tile = [&]() {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t qd = make_smem_desc(smQ + (uint32_t)kb * kRQBytes);
#pragma unroll
                    for (int k4 = 0; k4 < kGKBytes / 32; ++k4) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint64_t cd = make_smem_desc(smB + (uint32_t)stage * kBBytes + (uint32_t)h * (kBBytes / 2));
                            const uint32_t d_tmem = tmem_base + (uint32_t)acc * kRAccCols + (uint32_t)h * (kRSplit * kRN) + (uint32_t)k4 * kRN;
                            tc_mma<KIND>(d_tmem, cd + 2u * k4, qd + 2u * k4, a.idesc, (uint32_t)(kb != 0));
                        }
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                tc_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            };
            for (int t = 0; t < na; ++t) mma_tile();
            for (int sl = blockIdx.x; sl < a.S; sl += gridDim.x) {
                long long r0, r1;
                const int ntiles = slice_tiles(sl, r0, r1);
                for (int t = 0; t < ntiles; ++t) mma_tile();
            }
        }
    } else {   // ===== epilogue: thread <-> TMEM lane <-> corpus row of a half tile =====
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        // this thread's scores of the 16 queries against corpus row (half h, lane m) of the accumulator `acc`
        auto load_scores = [&](int h, float (&sc)[kRN]) {
            uint32_t v[kRSplit][kRN];
#pragma unroll
            for (int p = 0; p < kRSplit; ++p)
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * kRAccCols + (uint32_t)h * (kRSplit * kRN) + (uint32_t)p * kRN, v[p]);
#pragma unroll
            for (int j = 0; j < kRN; ++j)
                sc[j] = (__uint_as_float(v[0][j]) + __uint_as_float(v[1][j])) + (__uint_as_float(v[2][j]) + __uint_as_float(v[3][j]));
        };
        // ---- phase A: block maxima of the sample tiles ----
        for (int t = 0; t < s.sample_tiles; ++t) {
            const int blk0 = ((int)blockIdx.x * s.sample_tiles + t) * 2;
            if (t < na) {   // block-uniform
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const long long trow = f0 + (long long)t * kGN;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sc[kRN];
                    load_scores(h, sc);
                    const bool valid = trow + h * 128 + m < f1;   // rows past the slice / corpus (zero-filled by TMA) do not count
                    float mine = -INFINITY;
#pragma unroll
                    for (int j = 0; j < kRN; ++j) {
                        float x = valid ? sc[j] : -INFINITY;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) x = fmaxf(x, __shfl_xor_sync(kFull, x, off));
                        if (lane == j) mine = x;
                    }
                    if (lane < kRN) wmax[(h * 4 + quarter) * kRN + lane] = mine;
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                epi_bar_sync();
                if (m < 2 * kRN) {
                    const int h = m / kRN, j = m % kRN;
                    const float x = fmaxf(fmaxf(wmax[(h * 4 + 0) * kRN + j], wmax[(h * 4 + 1) * kRN + j]),
                                          fmaxf(wmax[(h * 4 + 2) * kRN + j], wmax[(h * 4 + 3) * kRN + j]));
                    if (j < a.nq) s.bmax[(size_t)j * nblk + blk0 + h] = x;
                }
                epi_bar_sync();   // wmax is rewritten by the next sample tile
            } else if (m < 2 * kRN) {   // this CTA's first slice is shorter than the sample: empty blocks
                const int h = m / kRN, j = m % kRN;
                if (j < a.nq) s.bmax[(size_t)j * nblk + blk0 + h] = -INFINITY;
            }
        }
        // ---- grid-wide barrier: every CTA's block maxima are written ----
        __threadfence();
        epi_bar_sync();
        if (m == 0) {
            atomicAdd(s.gbar, 1u);
            while ((int32_t)(ld_acquire_gpu(s.gbar) - s.gbar_target) < 0) __nanosleep(64);
        }
        epi_bar_sync();
        (void)ld_acquire_gpu(s.gbar);   // every thread acquires for its own loads
        // ---- thresholds: warp `quarter` serves queries quarter, quarter + 4, ... ----
        uint32_t* mybuf = selbuf + (size_t)quarter * nblk;
        for (int j = quarter; j < a.nq; j += 4) {
            for (int i = lane; i < nblk; i += kWarp) mybuf[i] = float_to_ordered(__ldcg(s.bmax + (size_t)j * nblk + i) + 0.0f);
            __syncwarp();
            const uint32_t kth = warp_kth_largest(mybuf, nblk, s.rank, lane);
            if (lane == 0) {
                const float e = s.eps_const + (s.eps_q ? __ldg(s.eps_q + j) : 0.0f);
                thr_sh[j] = __fsub_rd(__fsub_rd(ordered_to_float(kth), __fmul_ru(2.0f, e)), 2.384185791015625e-07f);
            }
            __syncwarp();
        }
        epi_bar_sync();
        float thr[kRN];
#pragma unroll
        for (int j = 0; j < kRN; ++j) thr[j] = j < a.nq ? thr_sh[j] : INFINITY;   // padding queries never append
        // ---- phase B: the sweep (gemm_rows_kernel's epilogue) ----
        for (int sl = blockIdx.x; sl < a.S; sl += gridDim.x) {
            long long r0, r1;
            const int ntiles = slice_tiles(sl, r0, r1);
            for (int t = 0; t < ntiles; ++t) {
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const long long trow = r0 + (long long)t * kGN;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sc[kRN];
                    load_scores(h, sc);
                    const long long row = trow + h * 128 + m;
                    if (row < r1) {
#pragma unroll
                        for (int j = 0; j < kRN; ++j) {
                            if (sc[j] >= thr[j]) {
                                const uint32_t pos = atomicAdd(a.cnt + j, 1u);
                                if (pos < (uint32_t)a.cap) a.cand[(size_t)j * a.cap + pos] = make_key(sc[j] + 0.0f, (uint32_t)row);
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace rfk
