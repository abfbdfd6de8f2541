// K2t: the small-batch HBM-bound scan, fed by the TMA engine.
//
// Same arithmetic, thresholds and candidate layout as scan_topk_kernel (kernels.cuh) - fp32 queries in registers,
// fp32 FMA, transposed butterfly, per-warp sorted top-K' lists - but the corpus does not travel through the
// register file's load path: one producer thread streams the CTA's contiguous row range into a 4-stage shared-
// memory ring with `cp.async.bulk` (one bulk copy of stage_rows * row_bytes per stage, completion on an
// mbarrier), eight consumer warps read rows back with conflict-free LDS.128.  192 KB are in flight per SM
// instead of the 96 KB the LDG version can hold in registers; scripts/tma_stream_probe.cu measured 7.3-7.4 TB/s
// for this request pattern with a consumer that frees slots at once.
//
// MEASURED (10M x 768 bf16, B200): 2.235 ms = 6.87 TB/s at batch 1 - no better than the LDG kernel (2.22 ms), and
// 2.37 ms at batch 2 (LDG: 2.24).  ncu: consumers wait on `full` 38 % of samples while the producer waits on
// `empty`: the CTA-wide slot hand-off (eight warps per 48 KB slot) idles the slots.  Releasing slots right after
// the FMAs did not help; polling with one lane per warp was slower.  Kept as an opt-in variant
// (ragfin_set_scan_variant(h, 2)), bit-exact under test; the LDG kernel stays the default.
//
//   cand[q * cand_q_stride + blockIdx.x * kp + i]   sorted descending, 0 = empty   (identical to scan_topk_kernel)
#pragma once
#include "gemm.cuh"      // mbarrier / smem helpers
#include "kernels.cuh"

namespace rfk {

constexpr int kTsConsumers = 8;                          // consumer warps
constexpr int kTsThreads = kWarp * (1 + kTsConsumers);   // warp 0 = producer
constexpr int kTsStages = 4;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// arrive that cannot be scheduled before `dep` is available (a register dependency, the value itself is unused)
__device__ __forceinline__ void mbar_arrive_after(uint32_t bar, float dep) {
    asm volatile("{\n\t.reg .f32 t;\n\tmov.f32 t, %1;\n\tmbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar), "f"(dep) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

// rows per stage: a multiple of (consumer warps * rows per warp iteration) holding at most ~48 KB
__host__ __device__ inline int ts_stage_rows(int steps, size_t row_bytes) {
    const int unit = kTsConsumers * scan_rows(steps);
    int units = (int)((48u << 10) / (row_bytes * unit));
    if (units < 1) units = 1;
    return units * unit;
}
__host__ __device__ inline size_t ts_smem_bytes(int steps, size_t row_bytes, int nq, int kp) {
    return 128 + (size_t)kTsStages * ts_stage_rows(steps, row_bytes) * row_bytes + (size_t)nq * kTsConsumers * kp * sizeof(u64) +
           2 * kTsStages * sizeof(uint64_t);
}

template <int DT, int NQ, int STEPS>
__global__ void __launch_bounds__(kTsThreads, 1)
scan_tma_kernel(const void* __restrict__ data, int64_t n_rows, int ld, const float* __restrict__ qhat, int nq_valid,
                int kp, u64* __restrict__ cand, int64_t cand_q_stride, const uint32_t* __restrict__ allow) {
    typedef Store<DT> S;
    constexpr int V = S::kVec;
    constexpr int R = scan_rows(STEPS);
    constexpr int M = R * NQ;
    constexpr int SH = 5 - ilog2(M);
    static_assert(M <= 32, "too many partial sums per lane");
    extern __shared__ __align__(128) uint8_t ts_raw[];
    const size_t row_bytes = (size_t)ld * sizeof(typename S::T);
    const int stage_rows = ts_stage_rows(STEPS, row_bytes);
    const uint32_t stage_bytes = (uint32_t)(stage_rows * row_bytes);
    const uint32_t ring = (smem_u32(ts_raw) + 127u) & ~127u;
    uint8_t* after = ts_raw + (ring - smem_u32(ts_raw)) + (size_t)kTsStages * stage_bytes;
    u64* lists = reinterpret_cast<u64*>(after);                                   // [NQ][kTsConsumers][kp]
    uint64_t* bars = reinterpret_cast<uint64_t*>(lists + (size_t)NQ * kTsConsumers * kp);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kTsStages + s); };

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NQ * kTsConsumers * kp; i += kTsThreads) lists[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTsStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kTsConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's contiguous range of stage-sized chunks
    const int64_t chunks = (n_rows + stage_rows - 1) / stage_rows;
    const int64_t per = (chunks + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = per * blockIdx.x;
    int64_t c1 = c0 + per;
    if (c1 > chunks) c1 = chunks;
    const char* base = reinterpret_cast<const char*>(data);

    if (warp == 0) {
        if (lane == 0) {   // ===== producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t c = c0; c < c1; ++c) {
                const int64_t row0 = c * stage_rows;
                const int64_t rows = n_rows - row0 < stage_rows ? n_rows - row0 : stage_rows;
                const uint32_t bytes = (uint32_t)(rows * row_bytes);
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_expect_tx(full_bar(stage), bytes);
                bulk_load_1d(ring + (uint32_t)stage * stage_bytes, base + (size_t)row0 * row_bytes, bytes, full_bar(stage));
                if (++stage == kTsStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {   // ===== consumers =====
        const int cw = warp - 1;
        const int nvec = ld / V;
        float qreg[NQ][STEPS * V];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int j = 0; j < STEPS; ++j) {
                const int vi = j * kWarp + lane;
#pragma unroll
                for (int e = 0; e < V; ++e) qreg[q][j * V + e] = vi < nvec ? qhat[(size_t)q * ld + vi * V + e] : 0.0f;
            }
        float tau[NQ];   // padding queries (zero rows beyond nq_valid) never collect
#pragma unroll
        for (int q = 0; q < NQ; ++q) tau[q] = q < nq_valid ? -INFINITY : INFINITY;

        int stage = 0;
        uint32_t phase = 0;
        for (int64_t c = c0; c < c1; ++c) {
            const int64_t crow0 = c * stage_rows;
            mbar_wait(full_bar(stage), phase);
            const uint32_t sbase = ring + (uint32_t)stage * stage_bytes;
            for (int r0 = cw * R; r0 < stage_rows; r0 += kTsConsumers * R) {
                const int64_t row0 = crow0 + r0;
                if (row0 >= n_rows) break;   // warp-uniform: the tail of the last chunk was not copied
                uint4 d[R][STEPS];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t rp = sbase + (uint32_t)((r0 + r) * row_bytes);
#pragma unroll
                    for (int j = 0; j < STEPS; ++j) {
                        const int vi = j * kWarp + lane;
                        d[r][j] = vi < nvec ? lds128(rp + (uint32_t)vi * 16u) : make_uint4(0, 0, 0, 0);
                    }
                }
                float acc[M];
#pragma unroll
                for (int i = 0; i < M; ++i) acc[i] = 0.0f;
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int j = 0; j < STEPS; ++j) {
                        float f[V];
                        S::unpack(d[r][j], f);
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
#pragma unroll
                            for (int e = 0; e < V; ++e) acc[r * NQ + q] = fmaf(f[e], qreg[q][j * V + e], acc[r * NQ + q]);
                    }
                // every byte this warp needs from the slot has been consumed by the FMAs above: release it before the
                // reduction / threshold work so that the refill overlaps them (the slot's idle time bounds bandwidth)
                if (r0 + kTsConsumers * R >= stage_rows) {
                    // The arrive names a value that depends on every LDS of this iteration.  An LDS is one warp-wide
                    // instruction whose destination registers become ready together, so lane 0's dependency covers
                    // the bytes all 32 lanes read.
                    float dep = acc[0];
#pragma unroll
                    for (int i = 1; i < M; ++i) dep += acc[i];
                    if (lane == 0) mbar_arrive_after(empty_bar(stage), dep);
                }
                TransposeReduce<M, 16>::run(acc, lane);
                const float s = acc[0];
                const int idx = lane >> SH;
                const int myq = idx % NQ;
                float mytau = tau[0];
#pragma unroll
                for (int q = 1; q < NQ; ++q) mytau = myq == q ? tau[q] : mytau;
                const bool leader = (lane & ((1 << SH) - 1)) == 0;
                const bool hit = leader && (row0 + idx / NQ < n_rows) && (s >= mytau);   // rows past the corpus hold stale bytes
                unsigned mask = __ballot_sync(kFull, hit);
                while (mask) {
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float bs = __shfl_sync(kFull, s, l) + 0.0f;
                    const int bidx = l >> SH;
                    const int q = bidx % NQ;
                    if (allow != nullptr && !row_allowed(allow, row0 + bidx / NQ)) continue;   // scalar filter (warp-uniform)
                    const u64 key = make_key(bs, (uint32_t)(row0 + bidx / NQ));
                    u64* list = lists + (size_t)(q * kTsConsumers + cw) * kp;
                    if (key > list[kp - 1]) {
                        warp_list_insert(list, kp, key, lane);
                        const u64 last = list[kp - 1];
                        const float nt = last ? key_score(last) : -INFINITY;
#pragma unroll
                        for (int qq = 0; qq < NQ; ++qq) tau[qq] = q == qq ? nt : tau[qq];
                    }
                }
            }
            if (++stage == kTsStages) { stage = 0; phase ^= 1u; }
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
        u64* region = lists + (size_t)q * kTsConsumers * kp;
        block_bitonic_sort_desc(region, kTsConsumers * kp, threadIdx.x, kTsThreads);
        u64* out = cand + (size_t)q * cand_q_stride + (size_t)blockIdx.x * kp;
        for (int i = threadIdx.x; i < kp; i += kTsThreads) out[i] = region[i];
    }
}

}  // namespace rfk
