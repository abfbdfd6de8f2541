// K3b: A-stationary variant of the tcgen05 path for 16-bit storage with at most 768 columns (the headline
// shape).  The 128 x K query tile lives in TENSOR MEMORY for a whole sweep (K/2 32-bit columns: 384 of the 512
// for K = 768; the epilogue threads write it once per work item with tcgen05.st), the MMA takes A from TMEM
// (tcgen05.mma ... [d], [a_tmem], b_desc), and shared memory is a deep ring of small corpus slabs
// (64 rows x 128 B = 8 KB).  Consequences:
//   * no query bytes are re-streamed from L2 per corpus tile (a third of the operand traffic of K3);
//   * the 12 k-block boxes of a 64-row tile are requested back to back, so every 1 536-byte corpus row is
//     fetched in one burst (DRAM-friendly) and ~190 KB per SM are in flight - this is what the HBM-bound
//     mid-batch regime needs;
//   * two 64-column accumulators (TMEM columns 384..511) double-buffer MMA and epilogue.
// Everything else (work items, clusters with TMA multicast, per-query lists, gtau) is as in gemm.cuh.
#pragma once
#include "gemm.cuh"

namespace rfk {

constexpr int kSN = 64;                       // corpus rows per tile (UMMA N)
constexpr int kSlabBytes = kSN * kGKBytes;    // 8 KB: one k-block of one tile
constexpr int kAColsMax = 384;                // TMEM columns reserved for the query tile (K <= 768 16-bit elements)
constexpr int kMaxSlots = 24;

__host__ __device__ constexpr int astat_slots(int kp) {
    // shared memory: slots * 8 KB + lists (128 * kp * 8 B) + barriers + 1 KB alignment slack <= 227 KB
    return kp <= 32 ? 24 : kp <= 64 ? 20 : 12;
}
__host__ __device__ constexpr size_t astat_smem_bytes(int kp) {
    return 1024 + (size_t)astat_slots(kp) * kSlabBytes + (size_t)kGM * kp * sizeof(u64) + 512;
}
__host__ __device__ constexpr uint32_t make_idesc_n(int ab_format, int n) {
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(kGM >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

struct AstatArgs {
    uint32_t idesc;
    int num_kblocks;        // ceil(ld / 64)
    int ld;                 // elements per stored row (multiple of 8)
    int nq;
    long long n_rows;
    int QT, S;
    long long rows_per_slice;   // multiple of kSN
    int slots;
    int kp;
    const uint16_t* q16;    // [nq][ld] queries rounded to the storage type
    u64* cand;              // [nq][S][kp]
    uint32_t* gtau;
    const uint32_t* allow;  // scalar filter bitmask over rows, or null
    float* dump;
};

template <bool DUMP, int C>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_astat_kernel(const __grid_constant__ CUtensorMap tmB, const AstatArgs a) {
    extern __shared__ uint8_t gsm_raw[];
    const uint32_t raw = smem_u32(gsm_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gsm = gsm_raw + (base - raw);
    const int slots = a.slots, kp = a.kp;
    const uint32_t smB = base;                                                       // [slots][8 KB]
    u64* lists = reinterpret_cast<u64*>(gsm + (size_t)slots * kSlabBytes);           // [kp][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + (size_t)slots * kSlabBytes + (size_t)kGM * kp * sizeof(u64));
    const uint32_t bar0 = smem_u32(bars);
    // barriers: full[0..24) empty[24..48) tmem_full[48..50) tmem_empty[50..52) a_ready[52]; tmem base at 53
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxSlots + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kMaxSlots + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxSlots + 2 + s); };
    const uint32_t aready_bar = bar0 + 8u * (2 * kMaxSlots + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxSlots + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = C > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << C) - 1u);
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const int n_groups = (a.QT + C - 1) / C;
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), C); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 128); }
        mbar_init(aready_bar, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d0 = tmem_base + kAColsMax;   // accumulators: columns 384..447 and 448..511

    const int n_items = n_groups * a.S;
    const int nkb = a.num_kblocks;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer: corpus slabs only =====
            constexpr int kSubRows = kSN / C;
            int slot = 0;
            uint32_t phase = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int sl = item / n_groups;
                const long long r0 = (long long)sl * a.rows_per_slice;
                long long r1 = r0 + a.rows_per_slice;
                if (r1 > a.n_rows) r1 = a.n_rows;
                const int ntiles = r1 > r0 ? (int)((r1 - r0 + kSN - 1) / kSN) : 0;
                for (int t = 0; t < ntiles; ++t) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(empty_bar(slot), phase ^ 1u);
                        mbar_expect_tx(full_bar(slot), kSlabBytes);
                        const uint32_t bdst = smB + (uint32_t)slot * kSlabBytes + crank * (uint32_t)(kSubRows * kGKBytes);
                        const int brow = (int)(r0 + (long long)t * kSN) + (int)crank * kSubRows;
                        if (C > 1) tma_load_2d_mc(bdst, &tmB, kb * 64, brow, full_bar(slot), cmask);
                        else tma_load_2d(bdst, &tmB, kb * 64, brow, full_bar(slot));
                        if (++slot == slots) { slot = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer: A from tensor memory =====
            int slot = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_phase = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int sl = item / n_groups;
                const long long r0 = (long long)sl * a.rows_per_slice;
                long long r1 = r0 + a.rows_per_slice;
                if (r1 > a.n_rows) r1 = a.n_rows;
                const int ntiles = r1 > r0 ? (int)((r1 - r0 + kSN - 1) / kSN) : 0;
                mbar_wait(aready_bar, a_phase);      // this item's query tile is in TMEM
                a_phase ^= 1u;
                tc_fence_after();
                for (int t = 0; t < ntiles; ++t) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_d0 + (uint32_t)acc * kSN;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(full_bar(slot), phase);
                        tc_fence_after();
                        const uint64_t bd = make_smem_desc(smB + (uint32_t)slot * kSlabBytes);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)   // 16 elements of K per instruction = 8 TMEM columns of A
                            tc_mma_ts(d_tmem, tmem_base + (uint32_t)(kb * 32 + k4 * 8), bd + 2u * k4, a.idesc,
                                      (uint32_t)((kb | k4) != 0));
                        if (C > 1) tc_commit_mc(empty_bar(slot), cmask);
                        else tc_commit(empty_bar(slot));
                        if (++slot == slots) { slot = 0; phase ^= 1u; }
                    }
                    tc_commit(tfull_bar(acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else {   // ===== epilogue warps: load the query tile into TMEM, then filter accumulators =====
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int qt = (item % n_groups) * C + (int)crank, sl = item / n_groups;
            const long long r0 = (long long)sl * a.rows_per_slice;
            long long r1 = r0 + a.rows_per_slice;
            if (r1 > a.n_rows) r1 = a.n_rows;
            const int ntiles = r1 > r0 ? (int)((r1 - r0 + kSN - 1) / kSN) : 0;
            const int q = qt * kGM + m;
            const int qc = q < a.nq ? q : a.nq - 1;
            // (every MMA of the previous item has completed: this thread saw its last tfull)
            {
                const uint4* qrow = reinterpret_cast<const uint4*>(a.q16 + (size_t)qc * a.ld);
                const int nvec = a.ld / 8;            // 16-byte vectors in a query row
                const int ncol_chunks = nkb * 32 / 32;   // 32 TMEM columns (64 elements) per k-block
                for (int ch = 0; ch < ncol_chunks; ++ch) {
                    uint32_t r[32];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int vi = ch * 8 + v;
                        uint4 w = make_uint4(0, 0, 0, 0);
                        if (q < a.nq && vi < nvec) w = __ldg(qrow + vi);
                        r[4 * v + 0] = w.x; r[4 * v + 1] = w.y; r[4 * v + 2] = w.z; r[4 * v + 3] = w.w;
                    }
                    tmem_st32(tmem_base + lane_base + (uint32_t)ch * 32, r);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(aready_bar);
            }
            CandState st;
            st.cnt = 0;
            st.minpos = 0;
            st.tau_key = 0;
            st.tau_s = q < a.nq ? -INFINITY : INFINITY;
            st.gptr = (!DUMP && q < a.nq) ? a.gtau + q : nullptr;
            st.allow = a.allow;
            uint32_t g_next = DUMP ? 0u : __ldcg(a.gtau + qc);
            for (int t = 0; t < ntiles; ++t) {
                if (!DUMP && (t & 3) == 0) {   // refresh the published bound every 256 rows
                    const uint32_t g = g_next;
                    if (g && q < a.nq) st.tau_s = fmaxf(st.tau_s, ordered_to_float(g));
                }
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                if (!DUMP && (t & 3) == 0) g_next = __ldcg(a.gtau + qc);
                const long long trow = r0 + (long long)t * kSN;
                const uint32_t taddr = tmem_d0 + lane_base + (uint32_t)acc * kSN;
                const int valid = r1 - trow < kSN ? (int)(r1 - trow) : kSN;
                uint32_t vb[2][32];
                tmem_ld32_async(taddr, vb[0]);
                tmem_ld32_async(taddr + 32, vb[1]);
#pragma unroll
                for (int c = 0; c < kSN / 32; ++c) {
                    uint32_t(&v)[32] = vb[c];
                    tmem_wait32(v);
                    if (DUMP) {
                        if (q < a.nq) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j < valid) a.dump[(size_t)q * a.n_rows + trow + c * 32 + j] = __uint_as_float(v[j]);
                        }
                        continue;
                    }
                    if (valid < c * 32 + 32) {   // last tile of a slice only: rows past the corpus must never qualify
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j >= valid) v[j] = 0x7FC00000u;   // NaN: fails every >= test, ignored by fmaxf
                    }
                    // maxima of the four groups of 8 columns, then of the chunk: one compare in the common case
                    float gmx[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float t = fmaxf(__uint_as_float(v[g * 8]), __uint_as_float(v[g * 8 + 1]));
#pragma unroll
                        for (int j = 2; j < 8; ++j) t = fmaxf(t, __uint_as_float(v[g * 8 + j]));
                        gmx[g] = t;
                    }
                    const float mx = fmaxf(fmaxf(gmx[0], gmx[1]), fmaxf(gmx[2], gmx[3]));
                    if (mx >= st.tau_s) {          // rare per thread; every v[j] stays in its register
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (gmx[g] >= st.tau_s) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float sc = __uint_as_float(v[g * 8 + j]);
                                    if (sc >= st.tau_s) cand_insert(st, lists, m, kp, sc, (uint32_t)(trow + c * 32 + g * 8 + j));
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (!DUMP && q < a.nq) {
                u64* out = a.cand + ((size_t)q * a.S + sl) * kp;
                for (int e = 0; e < kp; ++e) out[e] = e < st.cnt ? lists[(size_t)e * kGM + m] : 0ull;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace rfk
