// Probe: how fast can TMA stream a [N x 768] bf16 corpus into shared memory, by request pattern?
// No MMA, no epilogue: one producer thread per CTA fills a ring of stages, one consumer thread frees them.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_stream_probe tma_stream_probe.cu
//
//  P0  row-major, k-blocked: stage = one 2D box 256 rows x 128 B (what gemm_topk_kernel does today)
//  P1  row-major, stage = 32 full rows as 12 boxes of 32 rows x 128 B (k innermost in issue order)
//  P2  row-major, stage = 32 full rows as ONE 3D box (64, 12, 32): 48 KB contiguous in HBM
//  P3  8-row-interleaved layout [N/8][12][8][64], k-blocked: stage = 4D box (64, 8, 1, 32): 1 KB pieces, 12 KB apart
//  P4  8-row-interleaved layout, stage = 4D box (64, 8, 12, 4): 48 KB contiguous
//  P5  P0 plus a 16 KB query-tile box per stage (L2-resident A operand re-streamed, as today)
//  P6  P3 with 64-row stages (64,8,1,8) = 8 KB, 12 boxes per stage iterating k (a full-K 64-row tile, 96 KB/stage)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"((uint64_t)m), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk1(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"((uint64_t)m), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"((uint64_t)m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

struct Args { long long n_rows; int stages; int stage_bytes; const char* base; int pieces; };

template <int P>
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmA, const Args a) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + (base - smem_u32(raw)) + (size_t)a.stages * a.stage_bytes);
    const uint32_t bar0 = smem_u32(bars);
    const int S = a.stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (S + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // rows per stage-group ("tile") and number of stage fills per tile
    constexpr int TILE_ROWS = (P == 0 || P == 3 || P == 5) ? 256 : (P == 6 ? 64 : 32);   // P7: rows per stage = stage_bytes / 1536
    const int tile_rows = P == 7 ? a.stage_bytes / 1536 : TILE_ROWS;
    constexpr int FILLS = (P == 0 || P == 3 || P == 5) ? 12 : 1;
    const long long n_tiles = a.n_rows / tile_rows;
    const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = per * blockIdx.x;
    long long t1 = t0 + per;
    if (t1 > n_tiles) t1 = n_tiles;
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (long long t = t0; t < t1; ++t) {
            const int row0 = (int)(t * tile_rows);
            for (int f = 0; f < FILLS; ++f) {
                mbar_wait(bar0 + 8 * (S + stage), phase ^ 1u);
                const uint32_t full = bar0 + 8 * stage;
                const uint32_t dst = base + (uint32_t)stage * a.stage_bytes;
                mbar_expect_tx(full, a.stage_bytes);
                if (P == 0) tma2(dst, &tm, f * 64, row0, full);
                if (P == 5) { tma2(dst, &tm, f * 64, row0, full); tma2(dst + 32768, &tmA, f * 64, 0, full); }
                if (P == 1) for (int kb = 0; kb < 12; ++kb) tma2(dst + kb * 4096, &tm, kb * 64, row0, full);
                if (P == 2) tma3(dst, &tm, 0, 0, row0, full);
                if (P == 7) { const int pb = a.stage_bytes / a.pieces; for (int i = 0; i < a.pieces; ++i) bulk1(dst + i * pb, a.base + (size_t)row0 * 1536 + (size_t)i * pb, pb, full); }
                if (P == 3) tma4(dst, &tm, 0, 0, f, row0 / 8, full);
                if (P == 4) tma4(dst, &tm, 0, 0, 0, row0 / 8, full);
                if (P == 6) for (int kb = 0; kb < 12; ++kb) tma4(dst + kb * 8192, &tm, 0, 0, kb, row0 / 8, full);
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (long long t = t0; t < t1; ++t)
            for (int f = 0; f < FILLS; ++f) {
                mbar_wait(bar0 + 8 * stage, phase);
                mbar_arrive(bar0 + 8 * (S + stage));
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
    }
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap mk(enc_fn enc, void* p, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapL2promotion prom) {
    CUtensorMap m;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (rank %d)\n", (int)r, rank); exit(1); }
    return m;
}

static const char* g_base = nullptr;
static int g_pieces = 1;
template <int P>
static void run(const char* name, const CUtensorMap& tm, const CUtensorMap& tmA, long long n, int stages, int stage_bytes, int grid) {
    Args a{n, stages, stage_bytes, g_base, g_pieces};
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 256;
    CK(cudaFuncSetAttribute(probe<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0));
        probe<P><<<grid, 64, smem>>>(tm, tmA, a);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    printf("%-58s stages %d x %3d KB grid %3d : %7.3f ms  %7.1f GB/s\n", name, stages, stage_bytes / 1024, grid, best, (double)n * 1536 / best / 1e6);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 10000000LL / 256 * 256;
    void* p; CK(cudaMalloc(&p, (size_t)n * 1536)); CK(cudaMemset(p, 0, (size_t)n * 1536));
    g_base = (const char*)p;
    void* qa; CK(cudaMalloc(&qa, 128 * 1536)); CK(cudaMemset(qa, 0, 128 * 1536));
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    enc_fn enc = (enc_fn)fp;
    const CUtensorMapL2promotion proms[2] = {CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_L2_PROMOTION_NONE};
    for (int pi = 0; pi < 2; ++pi) {
        const CUtensorMapL2promotion pr = proms[pi];
        printf("--- L2 promotion %s\n", pi == 0 ? "256B" : "none");
        cuuint64_t d2[2] = {768, (cuuint64_t)n}, s2[1] = {1536};
        cuuint32_t b256[2] = {64, 256}, b32[2] = {64, 32}, b128[2] = {64, 128};
        cuuint64_t dA[2] = {768, 128};
        CUtensorMap tA = mk(enc, qa, 2, dA, s2, b128, pr);
        CUtensorMap t0 = mk(enc, p, 2, d2, s2, b256, pr);
        CUtensorMap t1 = mk(enc, p, 2, d2, s2, b32, pr);
        cuuint64_t d3[3] = {64, 12, (cuuint64_t)n}, s3[2] = {128, 1536};
        cuuint32_t b3[3] = {64, 12, 32};
        CUtensorMap t2 = mk(enc, p, 3, d3, s3, b3, pr);
        cuuint64_t d4[4] = {64, 8, 12, (cuuint64_t)(n / 8)}, s4[3] = {128, 1024, 12288};
        cuuint32_t b4a[4] = {64, 8, 1, 32}, b4b[4] = {64, 8, 12, 4}, b4c[4] = {64, 8, 1, 8};
        CUtensorMap t3 = mk(enc, p, 4, d4, s4, b4a, pr);
        CUtensorMap t4 = mk(enc, p, 4, d4, s4, b4b, pr);
        CUtensorMap t6 = mk(enc, p, 4, d4, s4, b4c, pr);
        for (int grid : {148, 296}) {
            const int st32 = grid == 148 ? 6 : 3, st48 = grid == 148 ? 4 : 2;
            run<0>("P0 row-major k-blocked 256x128B", t0, tA, n, grid == 148 ? 4 : 3, 32768, grid);
            run<0>("P0 row-major k-blocked 256x128B", t0, tA, n, st32, 32768, grid);
            run<5>("P5 = P0 + 16 KB query tile per stage", t0, tA, n, grid == 148 ? 4 : 2, 49152, grid);
            run<1>("P1 row-major 32 full rows, 12 boxes", t1, tA, n, st48, 49152, grid);
            run<2>("P2 row-major 32 full rows, one 3D box", t2, tA, n, st48, 49152, grid);
            run<3>("P3 8-row-interleaved k-blocked (1 KB pieces)", t3, tA, n, grid == 148 ? 4 : 3, 32768, grid);
            run<3>("P3 8-row-interleaved k-blocked (1 KB pieces)", t3, tA, n, st32, 32768, grid);
            run<4>("P4 8-row-interleaved contiguous 48 KB", t4, tA, n, st48, 49152, grid);
            if (pi == 0) {
                g_pieces = 1; run<7>("P7 1D cp.async.bulk, one 48 KB copy per stage", t0, tA, n, st48, 49152, grid);
                g_pieces = 4; run<7>("P7 1D cp.async.bulk, 4 x 12 KB copies per stage", t0, tA, n, st48, 49152, grid);
                g_pieces = 12; run<7>("P7 1D cp.async.bulk, 12 x 4 KB copies per stage", t0, tA, n, st48, 49152, grid);
                g_pieces = 1; run<7>("P7 1D cp.async.bulk, 24 KB stages", t0, tA, n, grid == 148 ? 8 : 4, 24576, grid);
                g_pieces = 1; run<7>("P7 1D cp.async.bulk, 12 KB stages", t0, tA, n, grid == 148 ? 16 : 8, 12288, grid);
            }
            if (grid == 148) run<6>("P6 8-row-interleaved 64-row full-K stage (12 x 8 KB)", t6, tA, n, 2, 98304, grid);
        }
    }
    return 0;
}
