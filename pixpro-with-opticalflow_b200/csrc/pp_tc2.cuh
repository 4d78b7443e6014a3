// pp_tc2.cuh — TMA-fed, warp-specialised tcgen05 3xTF32 batched GEMM (second generation of pp_tc.cuh).
//
//   C[b][m][n] = sum_k A[b][m][k] * B[b][n][k]          fp32 accuracy: a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
//
// Why a second kernel: ncu on pp_tc.cuh's tc_gemm_kernel (profiles/r01_s3_tc_ncu_summary.txt) showed the tensor pipe
// 32 % active with the L1 / shared-memory data pipe at 58 % and 0 bytes of TMA traffic — every operand went
// global -> registers -> hi/lo split -> STS.128 -> UMMA, with a CTA-wide __syncthreads per 16-wide K chunk.  Shared
// memory carries BOTH the staging stores and the tensor core's operand reads (128 B/clk/SM in total), so staging by
// threads caps the pipe near 60 % before any latency.  Here:
//   * operands arrive PRE-SPLIT: an elementwise prologue (which the PPM needs anyway: normalisation, relu^gamma,
//     gS + gS^T) writes the hi and lo planes once, K-major ([batch][rows][K] fp32), to global memory;
//   * one elected thread (warp 0) streams them with TMA (cp.async.bulk.tensor.3d, 128-byte swizzle, zero fill past the
//     matrix edges = no tail code) into a ring of STAGES stages, one elected thread (warp 1) issues the tcgen05.mma
//     instructions straight from the swizzled tiles and recycles a stage with tcgen05.commit -> mbarrier; no thread
//     ever touches operand data and there is no __syncthreads in the main loop;
//   * the CTA tile is 128 x 256 (UMMA M=128, N<=256): per 32-wide K chunk 96 KB of operands feed 12 MMAs = 1536
//     tensor-pipe cycles, 62 B/clk — against 85 B/clk for a 128 x 128 tile, which the L2 (~42 B/clk/SM at full
//     chip, /opt/skills/guides/B300_MICROARCH.md) cannot deliver.  The accumulators fill the SM's TMEM:
//     columns [0,256) hold hi*hi, [256,512) the 2^-11-smaller correction products (summed separately, added once
//     in fp32 in the epilogue: the tensor core aligns addends to the accumulator's exponent);
//   * the epilogue is run by all 8 warps from TMEM (tcgen05.ld 32x32b.x16), functor-fused as in pp_tc.cuh.
// N tiles are 256 wide except the last, which is issued with N = the remainder rounded up to 16 (784 = 3*256 + 16:
// the tail tile costs 1/16 of a full one instead of a whole padded tile).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "pp_common.cuh"
#include "pp_tc.cuh"

namespace pp {
namespace tc2 {

constexpr int TM = 128;
constexpr int TN = 256;
constexpr int TK = 32;  // fp32 elements = one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int THREADS = 256;
constexpr int kPersistentMaxK = 320;  // see launch_tc2
constexpr uint32_t A_TILE = TM * TK * 4;  // 16 KB
constexpr uint32_t B_TILE = TN * TK * 4;  // 32 KB
constexpr uint32_t STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;  // A_hi | A_lo | B_hi | B_lo
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + slack to round the base up to 1024 (128-byte swizzle atoms)

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, SWIZZLE_128B: rows of 128 bytes, 8-row atoms of
// 1024 bytes (SBO); the leading offset is not used by swizzled K-major layouts.  K steps inside the 128-byte row advance the
// start address by 32 bytes (the swizzle XOR is applied to the absolute address bits, hence the 1024-byte alignment).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;  // layout type: SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

struct Maps {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

// grid (ceil(N/TN), ceil(M/TM), batch); block 256; dynamic smem SMEM_BYTES; one CTA per SM (all 512 TMEM columns).
template <class EP>
__global__ void __launch_bounds__(THREADS, 1) tc2_gemm_kernel(const __grid_constant__ Maps maps, int M, int N, int K, EP ep) {
    extern __shared__ uint8_t tc2_smem_raw[];
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (tc::smem_u32(tc2_smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tn = min(TN, ((N - n0) + 15) & ~15);  // UMMA N of this tile
    const int nchunk = (K + TK - 1) / TK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(&acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) { prefetch_map(&maps.a_hi); prefetch_map(&maps.a_lo); prefetch_map(&maps.b_hi); prefetch_map(&maps.b_lo); }
    if (warp == 2) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
#pragma unroll 1
            for (int c = 0; c < nchunk; c++) {
                const int s = c % STAGES, round = c / STAGES;
                if (round > 0) tc::mbar_wait(&empty[s], (round - 1) & 1);  // the MMAs that read this stage have completed
                const uint32_t st = base + s * STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                tma_load_3d(st, &maps.a_hi, c * TK, m0, b, &full[s]);
                tma_load_3d(st + A_TILE, &maps.a_lo, c * TK, m0, b, &full[s]);
                tma_load_3d(st + 2 * A_TILE, &maps.b_hi, c * TK, n0, b, &full[s]);
                tma_load_3d(st + 2 * A_TILE + B_TILE, &maps.b_lo, c * TK, n0, b, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            const uint32_t idesc = tc::make_idesc(TM, tn);
#pragma unroll 1
            for (int c = 0; c < nchunk; c++) {
                const int s = c % STAGES, round = c / STAGES;
                tc::mbar_wait(&full[s], round & 1);
                tc::fence_after_sync();
                const uint32_t a_hi = base + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                for (int kk = 0; kk < TK / 8; kk++) {  // one MMA consumes K = 8 fp32 = 32 bytes of the swizzle row
                    const uint32_t ko = kk * 32;
                    const uint32_t acc = (c > 0 || kk > 0) ? 1u : 0u;
                    tc::mma_tf32(tmem_d, desc_sw128(a_hi + ko), desc_sw128(b_hi + ko), idesc, acc);
                    tc::mma_tf32(tmem_d + TN, desc_sw128(a_hi + ko), desc_sw128(b_lo + ko), idesc, acc);
                    tc::mma_tf32(tmem_d + TN, desc_sw128(a_lo + ko), desc_sw128(b_hi + ko), idesc, 1u);
                }
                tc::mma_commit(&empty[s]);  // frees the stage when these MMAs have read it
            }
            tc::mma_commit(&acc_full);      // every MMA has completed: accumulators final
        }
        __syncwarp();
    }
    // ===== epilogue (all 8 warps): thread t of warp w owns accumulator row 32 (w % 4) + t; warps w, w + 4 alternate 16-column groups
    tc::mbar_wait(&acc_full, 0);
    tc::fence_after_sync();
    {
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int j = 16 * (warp >> 2); j < tn; j += 32) {
            float v[16], w[16];
            tc::tmem_ld16(lane_addr + j, v);
            tc::tmem_ld16(lane_addr + TN + j, w);
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] += w[i];
            if (m < M && n0 + j < N) ep.store16(b, m, n0 + j, v);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_d, 512);
}

// ---- persistent variant: TMEM double-buffered accumulators, epilogue overlapped with the next tile's main loop ------------
// ncu / timing of the kernel above on the 28x28 similarity shape (K = 256: 8 chunks per tile): the main loop is ~6 us of tensor
// work per tile, around it ~1.5 us of prologue (barrier init, TMEM allocation, first TMA round trip) and >= 2 us of epilogue
// (512 TMEM columns at 64 B/clk plus the stores) that nothing overlaps.  Here the three products of 3xTF32 accumulate into ONE
// 256-column tile (measured error 1.1e-6 -> see tests/test_gpu_tc.py; the separate correction tile of the kernel above halves
// the error and is kept for K > kPersistentMaxK), which leaves room for two accumulator buffers: a CTA per SM walks the tile list
// (n fastest, so neighbouring CTAs share A rows in L2), warp 0 streams operands, warp 1 issues MMAs into buffer i & 1, warps 4-7
// drain buffer (i - 1) & 1 meanwhile.  Barriers: full/empty per smem stage, acc_full/acc_empty per accumulator buffer.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane: two x16 loads in flight, one wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// grid = min(#tiles, #SMs) CTAs; block 256; dynamic smem SMEM_BYTES.
template <class EP>
__global__ void __launch_bounds__(THREADS, 1) tc2_gemm_persistent_kernel(const __grid_constant__ Maps maps, int M, int N, int K, int batch, EP ep) {
    extern __shared__ uint8_t tc2_smem_raw[];
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (tc::smem_u32(tc2_smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = (M + TM - 1) / TM, NT = (N + TN - 1) / TN;
    const int ntiles = batch * MT * NT;
    const int nchunk = (K + TK - 1) / TK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int u = 0; u < 2; u++) { tc::mbar_init(&acc_full[u], 1); tc::mbar_init(&acc_empty[u], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) { prefetch_map(&maps.a_hi); prefetch_map(&maps.a_lo); prefetch_map(&maps.b_hi); prefetch_map(&maps.b_lo); }
    if (warp == 2) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            int g = 0;  // chunks issued so far by this CTA (stage ring position)
#pragma unroll 1
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int nt = t % NT, mt = (t / NT) % MT, b = t / (NT * MT);
#pragma unroll 1
                for (int c = 0; c < nchunk; c++, g++) {
                    const int s = g % STAGES, round = g / STAGES;
                    if (round > 0) tc::mbar_wait(&empty[s], (round - 1) & 1);
                    const uint32_t st = base + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                    tma_load_3d(st, &maps.a_hi, c * TK, mt * TM, b, &full[s]);
                    tma_load_3d(st + A_TILE, &maps.a_lo, c * TK, mt * TM, b, &full[s]);
                    tma_load_3d(st + 2 * A_TILE, &maps.b_hi, c * TK, nt * TN, b, &full[s]);
                    tma_load_3d(st + 2 * A_TILE + B_TILE, &maps.b_lo, c * TK, nt * TN, b, &full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            int g = 0, i = 0;
#pragma unroll 1
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
                const int nt = t % NT;
                const int tn = min(TN, ((N - nt * TN) + 15) & ~15);
                const uint32_t idesc = tc::make_idesc(TM, tn);
                const int buf = i & 1, use = i >> 1;
                if (use > 0) tc::mbar_wait(&acc_empty[buf], (use - 1) & 1);  // the epilogue has drained this buffer
                tc::fence_after_sync();
                const uint32_t acc_addr = tmem_d + buf * TN;
#pragma unroll 1
                for (int c = 0; c < nchunk; c++, g++) {
                    const int s = g % STAGES, round = g / STAGES;
                    tc::mbar_wait(&full[s], round & 1);
                    tc::fence_after_sync();
                    const uint32_t a_hi = base + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                    for (int kk = 0; kk < TK / 8; kk++) {
                        const uint32_t ko = kk * 32;
                        // small products first: they are summed among themselves before the large one joins
                        tc::mma_tf32(acc_addr, desc_sw128(a_hi + ko), desc_sw128(b_lo + ko), idesc, (c > 0 || kk > 0) ? 1u : 0u);
                        tc::mma_tf32(acc_addr, desc_sw128(a_lo + ko), desc_sw128(b_hi + ko), idesc, 1u);
                        tc::mma_tf32(acc_addr, desc_sw128(a_hi + ko), desc_sw128(b_hi + ko), idesc, 1u);
                    }
                    tc::mma_commit(&empty[s]);
                }
                tc::mma_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {  // ===== epilogue warps: warp w owns TMEM lanes 32 (w % 4) .. + 31 =====
        int i = 0;
#pragma unroll 1
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
            const int nt = t % NT, mt = (t / NT) % MT, b = t / (NT * MT);
            const int n0 = nt * TN;
            const int tn = min(TN, ((N - n0) + 15) & ~15);
            const int buf = i & 1, use = i >> 1;
            tc::mbar_wait(&acc_full[buf], use & 1);
            tc::fence_after_sync();
            const int q = warp & 3;
            const int m = mt * TM + q * 32 + lane;
            const uint32_t lane_addr = tmem_d + buf * TN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int j = 0; j < tn; j += 32) {
                float v[32];
                if (j + 32 <= tn) {
                    tmem_ld32(lane_addr + j, v);
                } else {
                    tc::tmem_ld16(lane_addr + j, v);
                }
                if (m < M && n0 + j < N) ep.store16(b, m, n0 + j, v);
                if (m < M && j + 16 < tn && n0 + j + 16 < N) ep.store16(b, m, n0 + j + 16, v + 16);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_d, 512);
}

// hi = the bits a TF32 tensor core keeps, lo = x - hi (exact); n % 4 == 0, 16-byte aligned
__global__ void __launch_bounds__(256) split_kernel(const float4* __restrict__ x, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    hi[i] = h;
    lo[i] = l;
}
__device__ __forceinline__ void split1(float v, float& h, float& l) {
    h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    l = v - h;
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// plane [batch][rows][K] fp32 viewed as a 3-D tensor (K fastest); box = TK x box_rows x 1, 128-byte swizzle
static inline bool make_plane_map(CUtensorMap* tm, const float* base, int64_t batch, int rows, int K, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)rows * K * 4};
    cuuint32_t box[3] = {(cuuint32_t)TK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int num_sms() {  // of the current device
    static int n[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& v = n[dev & 63];
    if (v == 0) {
        int q = 0;
        cudaDeviceGetAttribute(&q, cudaDevAttrMultiProcessorCount, dev);
        v = q > 0 ? q : 148;
    }
    return v;
}

static inline bool applicable(int K, const void* a, const void* b, const void* c, const void* d) {
    static const int off = [] { const char* e = getenv("PIXPRO_B200_TC2"); return (e && e[0] == '0') ? 1 : 0; }();
    return !off && K % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}

// Planes: A_hi/A_lo [batch][M][K], B_hi/B_lo [batch][N][K].  Returns -1 when not applicable (caller uses pp_tc.cuh's kernel).
template <class EP>
static inline int launch_tc2(const char* what, int64_t batch, int M, int N, int K, const float* a_hi, const float* a_lo, const float* b_hi,
                             const float* b_lo, EP ep, cudaStream_t st) {
    if (!applicable(K, a_hi, a_lo, b_hi, b_lo) || batch > 65535) return -1;
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    if (!make_plane_map(&maps.a_hi, a_hi, batch, M, K, TM) || !make_plane_map(&maps.a_lo, a_lo, batch, M, K, TM) ||
        !make_plane_map(&maps.b_hi, b_hi, batch, N, K, TN) || !make_plane_map(&maps.b_lo, b_lo, batch, N, K, TN))
        return -1;
    static const int mode = [] { const char* e = getenv("PIXPRO_B200_TC2"); return e ? atoi(e) : 1; }();  // 2: force the one-tile-per-CTA kernel
    // one accumulator tile takes 3 K / 8 truncating accumulate steps (measured: 2.5e-6 of the result's scale at K = 256,
    // 6.5e-6 at K = 784); two tiles (the kernel above) take K / 8 into the large one: the long-K products stay there
    if ((K <= kPersistentMaxK && mode != 2) || mode == 3) {
        auto kern = tc2_gemm_persistent_kernel<EP>;
        static unsigned long long opted = 0;  // per template instantiation, one bit per device
        if (smem_opt_in(kern, (int)SMEM_BYTES, opted) != cudaSuccess) {
            set_error("%s: cudaFuncSetAttribute(%u B of shared memory) failed: %s", what, SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
            return PP_ERR_CUDA;
        }
        const int64_t ntiles = batch * (int64_t)((M + TM - 1) / TM) * ((N + TN - 1) / TN);
        if (ntiles < (1ll << 31)) {
            const unsigned grid = (unsigned)(ntiles < num_sms() ? ntiles : num_sms());
            PP_LAUNCH(what, st, kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, M, N, K, (int)batch, ep));
            return check_launch(what);
        }
    }
    auto kern = tc2_gemm_kernel<EP>;
    static unsigned long long opted = 0;  // per template instantiation, one bit per device
    if (smem_opt_in(kern, (int)SMEM_BYTES, opted) != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%u B of shared memory) failed: %s", what, SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
        return PP_ERR_CUDA;
    }
    dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, (unsigned)batch);
    PP_LAUNCH(what, st, kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, M, N, K, ep));
    return check_launch(what);
}

static inline int launch_split(const float* x, int64_t n, float* hi, float* lo, cudaStream_t st) {
    const int64_t n4 = n / 4;
    PP_LAUNCH("tc split", st, split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n4,
                                                                                        reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo)));
    return check_launch("tc split");
}

}  // namespace tc2
}  // namespace pp
