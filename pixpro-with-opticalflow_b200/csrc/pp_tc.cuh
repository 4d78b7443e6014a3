// pp_tc.cuh — tcgen05 (5th-gen tensor core) building blocks for the dense contractions of the
// pixel path at large grids (P = G*G >= 128: the 14x14 and 28x28 feature grids).
//
// Precision: the path is specified in fp32 with a 1e-5 relative bar, which single-pass TF32
// (10-bit mantissa) cannot meet.  Every fp32 operand is therefore split into hi = the 19 bits a
// TF32 tensor core keeps and lo = x - hi (exact), and each logical product is issued as three
// tcgen05.mma kind::tf32 instructions accumulating into the same TMEM tile:
//        a*b  ~=  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi        (dropped term ~2^-22 relative)
// with fp32 accumulation in TMEM — "3xTF32".
//
// Data path: operands are staged by the CTA's threads (not TMA: tiles need an elementwise
// prologue — normalisation, relu^γ, hi/lo split — that TMA cannot apply) into shared memory in
// the UMMA canonical K-major, no-swizzle layout: 8-row x 16-byte core matrices, core matrices of
// consecutive 8-row groups SBO bytes apart, the two 16-byte K halves of one instruction LBO bytes
// apart.  One elected thread issues the MMAs and commits them to an mbarrier; the accumulator
// tile lives in TMEM (128 lanes x N columns) and is read back with tcgen05.ld for the epilogue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "pp_common.cuh"

namespace pp {
namespace tc {

constexpr int TM = 128;  // accumulator rows  (UMMA M, cta_group::1)
constexpr int TN = 128;  // accumulator columns (UMMA N)
constexpr int TK = 16;   // K extent of one staged chunk (fp32 elements) = 2 MMA k-steps of 8; 66 KB per CTA -> 2 CTAs per SM
constexpr uint32_t SBO = 128;             // bytes between 8-row core matrices
constexpr uint32_t LBO = (TM / 8) * 128 + 32;  // bytes between 16-byte K columns of core matrices (+32: a quarter-warp of the row-wise staging
                                               // map = 2 rows x 4 K columns -> 16-byte slots 2*kc + (row & 1), all distinct: conflict-free STS.128)
constexpr uint32_t TILE_BYTES = (TK / 4) * LBO;  // one staged operand tile (TM == TN)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (uint64_t)(LBO >> 4) << 16;                // leading byte offset, bits [16,30)
    d |= (uint64_t)(SBO >> 4) << 32;                // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
    return d;                                       // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// tcgen05.commit: the mbarrier is arrived on when all previously issued MMAs of this thread completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// hi/lo split of 4 values and store as one 16-byte core-matrix row in each of the two tiles
__device__ __forceinline__ void split_store(uint8_t* tile_hi, uint8_t* tile_lo, int row, int kchunk, float4 v) {
    const uint32_t off = (uint32_t)kchunk * LBO + (uint32_t)(row >> 3) * SBO + (uint32_t)(row & 7) * 16;
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(tile_hi + off) = h;
    *reinterpret_cast<float4*>(tile_lo + off) = l;
}

// ---- batched C[b] = A[b] * B[b]^T  (A: M x K, B: N x K, both presented K-major by their loaders)
// LA/LB::load4(b, row, k) -> 4 consecutive-k values of operand row `row` (zeros out of range);
// LA/LB::kRowMajorK: true if k is the contiguous index in memory (stage row-wise for coalescing).
// EP::store16(b, m, n, v): 16 consecutive columns n..n+15 of row m.
// grid (ceil(N/TN), ceil(M/TM), batch); block 256 threads (8 warps stage operands, warps 0-3 run
// the epilogue, thread 0 issues the MMAs); dynamic smem = 2 stages x 4 tiles = 66 KB, 256 TMEM
// columns: two CTAs per SM, so one CTA's staging overlaps the other's tensor-core work
// (ncu, profiles/r01_m_*: with one 132 KB CTA per SM the tensor pipe sat idle 84 % of the time).
// Pipeline: the global loads of chunk c+1 are issued into registers before chunk c is split and
// stored, so their latency hides behind the store phase, the barrier and the tensor-core work.
constexpr int STAGES = 2;
constexpr int TC_THREADS = 256;
constexpr int ITEMS = (TM * (TK / 4)) / TC_THREADS;  // float4 items per thread per operand per chunk (= 4)
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;
constexpr uint32_t TC_SMEM_BYTES = STAGES * STAGE_BYTES + 64;

template <class L>
__device__ __forceinline__ void item_coords(int it, int& row, int& kc) {
    const int item = it * TC_THREADS + threadIdx.x;
    constexpr int KC = TK / 4;  // 16-byte K columns per chunk
    row = L::kRowMajorK ? (item / KC) : (item & (TM - 1));
    kc = L::kRowMajorK ? (item % KC) : (item >> 7);
}
template <class L>
__device__ __forceinline__ void load_operand(const L& ld, int64_t b, int row0, int k0, float4 r[ITEMS]) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        int row, kc;
        item_coords<L>(it, row, kc);
        r[it] = ld.load4(b, row0 + row, k0 + 4 * kc);
    }
}
template <class L>
__device__ __forceinline__ void store_operand(const float4 r[ITEMS], uint8_t* hi, uint8_t* lo) {
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        int row, kc;
        item_coords<L>(it, row, kc);
        split_store(hi, lo, row, kc, r[it]);
    }
}

template <class LA, class LB, class EP>
__global__ void __launch_bounds__(TC_THREADS, 2) tc_gemm_kernel(int M, int N, int K, LA la, LB lb, EP ep) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + STAGES * STAGE_BYTES);  // one per stage
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + STAGES);
    const int64_t b = blockIdx.z;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 2 * TN);  // columns [0,TN): hi*hi sums; [TN,2TN): the small correction terms
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc(TM, TN);
    const int nchunk = (K + TK - 1) / TK;
    // Register prefetch TWO chunks ahead: three register sets rotate (the loop is unrolled by three so the
    // rotation is a renaming, not a copy that would wait for the loads it copies).  The loads of chunk c+2 are
    // issued before chunk c is staged, so each has two iterations (stores, barrier, tensor-core work) to land:
    // the iteration no longer waits out a full global-load latency (ncu: long-scoreboard was the top stall).
    float4 ra[3][ITEMS], rb[3][ITEMS];
    load_operand(la, b, m0, 0, ra[0]);
    load_operand(lb, b, n0, 0, rb[0]);
    if (nchunk > 1) {
        load_operand(la, b, m0, TK, ra[1]);
        load_operand(lb, b, n0, TK, rb[1]);
    }
    auto body = [&](int c, const float4* ca, const float4* cb, float4* pa, float4* pb) {
        const int s = c & 1;
        uint8_t* st = tc_smem + s * STAGE_BYTES;
        if (c + 2 < nchunk) {  // prefetch chunk c+2 into the free register set
            load_operand(la, b, m0, (c + 2) * TK, pa);
            load_operand(lb, b, n0, (c + 2) * TK, pb);
        }
        if (c >= STAGES) mbar_wait(&bars[s], ((c >> 1) - 1) & 1);  // the MMAs that read this stage have completed
        store_operand<LA>(ca, st, st + TILE_BYTES);
        store_operand<LB>(cb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
        fence_smem_to_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
#pragma unroll
            for (int kk = 0; kk < TK / 8; kk++) {  // one MMA consumes K = 8 fp32 = two 16-byte columns
                const uint32_t ko = kk * 2 * LBO;
                const uint32_t acc = (c > 0 || kk > 0) ? 1u : 0u;
                // The tensor core aligns and truncates addends to the accumulator's exponent, so the
                // 2^-11-times-smaller correction products are summed in their own accumulator tile
                // and added to the main one once, in fp32, in the epilogue (measured: 2x lower error).
                mma_tf32(tmem_d, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, acc);
                mma_tf32(tmem_d + TN, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, acc);
                mma_tf32(tmem_d + TN, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, 1u);
            }
            mma_commit(&bars[s]);
        }
    };
#pragma unroll 1
    for (int c = 0; c < nchunk; c += 3) {
        body(c, ra[0], rb[0], ra[2], rb[2]);
        if (c + 1 < nchunk) body(c + 1, ra[1], rb[1], ra[0], rb[0]);
        if (c + 2 < nchunk) body(c + 2, ra[2], rb[2], ra[1], rb[1]);
    }
    {
        const int cl = nchunk - 1;
        mbar_wait(&bars[cl & 1], (cl >> 1) & 1);  // commit of the last chunk: every MMA has completed
        fence_after_sync();
    }
    // epilogue (all 8 warps): thread t of warp w owns accumulator row 32 (w % 4) + t (a warp can only read the TMEM
    // lanes of its own quarter); warps w and w + 4 split the columns.  (ncu: with warps 0-3 alone the other four
    // idled at the final barrier for ~30 % of a 16-chunk CTA's lifetime.)
    {
        const int q = warp & 3;
        const int m = m0 + q * 32 + (threadIdx.x & 31);
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
        const int j0 = (warp >> 2) * (TN / 2);
#pragma unroll 1
        for (int j = j0; j < j0 + TN / 2; j += 16) {
            float v[16], w[16];
            tmem_ld16(lane_addr + j, v);
            tmem_ld16(lane_addr + TN + j, w);
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] += w[i];
            if (m < M && n0 + j < N) ep.store16(b, m, n0 + j, v);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 2 * TN);
}

}  // namespace tc

// ---- loaders / epilogues shared by the PPM and loss contractions -----------------------------------
__device__ __forceinline__ float4 ldg4_guard(const float* q, int k, int K) {  // 4 consecutive values, zero beyond K
    if (k + 3 < K && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(q));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) v.x = __ldg(q);
    if (k + 1 < K) v.y = __ldg(q + 1);
    if (k + 2 < K) v.z = __ldg(q + 2);
    if (k + 3 < K) v.w = __ldg(q + 3);
    return v;
}
struct TcLdN {  // natural: memory [rows][K], k contiguous
    static constexpr bool kRowMajorK = true;
    const float* p;
    int rows, K;
    __device__ __forceinline__ float4 load4(int64_t b, int row, int k) const {
        if (row >= rows || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg4_guard(p + (b * rows + row) * (int64_t)K + k, k, K);
    }
};
struct TcLdT {  // transposed: operand row = spatial index i, k = channel c; memory [C][P] (i contiguous)
    static constexpr bool kRowMajorK = false;
    const float* p;
    int C, P;
    __device__ __forceinline__ float4 load4(int64_t b, int i, int c) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < P) {
            const float* q = p + (b * C + c) * (int64_t)P + i;
            if (c < C) v.x = __ldg(q);
            if (c + 1 < C) v.y = __ldg(q + P);
            if (c + 2 < C) v.z = __ldg(q + 2 * (int64_t)P);
            if (c + 3 < C) v.w = __ldg(q + 3 * (int64_t)P);
        }
        return v;
    }
};
// 4 consecutive columns n..n+3 of row m (the TMA-fed kernel's epilogue, pp_tc2.cuh): one 16-byte store when aligned
__device__ __forceinline__ void st4_guard(float* q, int n, int N, float4 v) {
    if (n + 3 < N && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
        *reinterpret_cast<float4*>(q) = v;
    } else {
        q[0] = v.x;
        if (n + 1 < N) q[1] = v.y;
        if (n + 2 < N) q[2] = v.z;
        if (n + 3 < N) q[3] = v.w;
    }
}
struct TcStN {  // out[b][m][n..n+15]
    static constexpr bool kAux = false;
    float* out;
    int M, N;
    __device__ __forceinline__ void store4(int64_t b, int m, int n, float4 v) const { st4_guard(out + (b * M + m) * (int64_t)N + n, n, N, v); }
    __device__ __forceinline__ void store16(int64_t b, int m, int n, const float v[16]) const {
        float* q = out + (b * M + m) * (int64_t)N + n;
        if (n + 15 < N && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 4; i++) reinterpret_cast<float4*>(q)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (n + i < N) q[i] = v[i];
        }
    }
};
template <class LA, class LB, class EP>
static inline int launch_tc(const char* what, int64_t B, int M, int N, int K, LA la, LB lb, EP ep, cudaStream_t st) {
    dim3 grid((N + tc::TN - 1) / tc::TN, (M + tc::TM - 1) / tc::TM, (unsigned)B);
    auto kern = tc::tc_gemm_kernel<LA, LB, EP>;
    static unsigned long long opted = 0;  // per template instantiation, one bit per device
    if (smem_opt_in(kern, (int)tc::TC_SMEM_BYTES, opted) != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%u B of shared memory) failed: %s", what, tc::TC_SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
        return PP_ERR_CUDA;
    }
    PP_LAUNCH(what, st, kern<<<grid, tc::TC_THREADS, tc::TC_SMEM_BYTES, st>>>(M, N, K, la, lb, ep));
    return check_launch(what);
}

// Grids of 14x14 and up (P >= 128) are large enough to fill 128-row tensor-core tiles
// (BASELINE.json north_star); PIXPRO_B200_NO_TC=1 forces the CUDA-core path for A/B runs.
static inline bool use_tensor_cores(int P) {
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("PIXPRO_B200_NO_TC");
        disabled = (e && e[0] == '1') ? 1 : 0;
    }
    return !disabled && P >= 128;
}


}  // namespace pp
