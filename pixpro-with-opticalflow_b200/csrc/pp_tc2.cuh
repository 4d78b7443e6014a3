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
//   * the CTA tile is 128 x 256 (UMMA M=128, N<=256): per 32 columns of K, 96 KB of operands feed 12 MMAs = 1536
//     tensor-pipe cycles, 62 B/clk — against 85 B/clk for a 128 x 128 tile, which the L2 (~42 B/clk/SM at full
//     chip, /opt/skills/guides/B300_MICROARCH.md; measured here: 41 B/clk/SM on 8 x 2048^3) cannot deliver.  The
//     accumulators fill the SM's TMEM: columns [0,256) hold hi*hi, [256,512) the correction products;
//   * the kernel is persistent (one CTA per SM walks the tile list) and the epilogue is run by 8 dedicated warps from TMEM
//     (tcgen05.ld 32x32b.x16), functor-fused as in pp_tc.cuh, while the producer already refills the ring for the next tile.
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
constexpr int THREADS = 512;  // warp 0: TMA producer, 1: MMA issuer, 2: TMEM allocation, 2-3 and 12-15: converters, 4-11: epilogue
constexpr int NCONV = 6 * 32;  // converter threads
constexpr uint32_t RING_BYTES = 192 * 1024;
constexpr int EPI_STRIDE = 20;  // floats per row of an epilogue warp's 32 x 16 transpose buffer (80 bytes: conflict-free 16-byte writes)
constexpr uint32_t EPI_BYTES = 8 * 32 * EPI_STRIDE * 4;  // 8 epilogue warps
constexpr uint32_t SMEM_BYTES = RING_BYTES + EPI_BYTES + 1024;  // + slack to round the base up to 1024 (swizzle atoms)

// One K chunk = one swizzle row: TK fp32 = 128 bytes (SWIZZLE_128B, 2 stages of 96 KB) or 64 bytes (SWIZZLE_64B, 4 stages of
// 48 KB).  Same bytes in flight; the deeper ring of smaller stages keeps the tensor pipe fed across a TMA round trip
// (measured with 2 x 96 KB: the refill of a stage — ~2 us for four boxes — outlasts the 0.8 us the other stage feeds the MMAs).
template <int TK>
struct Cfg {
    static_assert(TK == 32 || TK == 16, "one chunk = one 128- or 64-byte swizzle row");
    static constexpr uint32_t kRowBytes = TK * 4;
    static constexpr uint32_t kATile = TM * kRowBytes, kBTile = TN * kRowBytes;
    static constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBTile;  // A_hi | A_lo | B_hi | B_lo
    static constexpr int kStages = RING_BYTES / kStageBytes;
    static constexpr uint64_t kLayoutType = TK == 32 ? 2 : 4;  // cute::UMMA::LayoutType SWIZZLE_128B / SWIZZLE_64B
    static constexpr uint32_t kSBO = 8 * kRowBytes;            // one 8-row swizzle atom
};

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, swizzled: rows of one swizzle span, 8-row atoms
// SBO bytes apart; the leading offset is not used by swizzled K-major layouts.  K steps inside the row advance the start
// address by 32 bytes (the swizzle XOR is applied to absolute address bits, hence the 1024-byte aligned ring).
template <int TK>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(Cfg<TK>::kSBO >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    d |= Cfg<TK>::kLayoutType << 61;
    return d;
}

// MN-major operand (memory [K][rows], rows contiguous — e.g. a [C, P] feature map used as the P x C operand).  For 32-bit
// elements the only MN-major layout the tensor core reads is SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>:
// 32-byte chunks of a 128-byte line XORed with the line index mod 4; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  The tile is
// staged as boxes of 32 rows x TK k-lines (one 128-byte line per k): 4 k-lines form a 512-byte atom (SBO), consecutive 32-row
// groups are one box = TK * 128 bytes apart (LBO).  One MMA (K = 8) consumes 8 k-lines = 1024 bytes of every group.
template <int TK>
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((TK * 128) >> 4) << 16;  // LBO: next 32 rows
    d |= (uint64_t)(512 >> 4) << 32;         // SBO: next 4 k-lines
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(tc::smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane from each of two tiles: both loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t addr_a, uint32_t addr_b, float v[16], float w[16]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr_a));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr_b));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) { v[i] = __uint_as_float(r[i]); w[i] = __uint_as_float(r[16 + i]); }
}

// Two operand sets: the K loop runs over set 0 (K chunks [0, nchunk0)) and then, optionally, over set 1 — a second product
// accumulated into the same tile (C = A0 B0^T + A1 B1^T: how the PPM backward forms G + G^T without ever writing G).
struct Maps {
    CUtensorMap a_hi[2], a_lo[2], b_hi[2], b_lo[2];
    // MN-major operands only: the same plane viewed 4-D as [batch][row group][K][32 rows], so that ONE box brings all the 32-row
    // groups of a tile (TM / 32 or TN / 32) in the staged order [group][k-line][32 rows]; used for tiles whose groups are all
    // complete (a ragged last tile takes one 3-D box per group from a_lo / b_lo, whose row dimension clips properly)
    CUtensorMap a_grp[2], b_grp[2];      // the plane the lo slot receives (the single fp32 plane of a raw operand, else the lo plane)
    CUtensorMap a_grp_hi[2], b_grp_hi[2];  // pre-split MN-major operands: the hi plane
};
// in-place split of a staged fp32 tile: 16-byte chunk i of `lo` -> hi part to chunk i of `hi`, lo part back (NCONV converter threads)
__device__ __forceinline__ void split_tile(uint8_t* hi, uint8_t* lo, int n16, int tid) {
#pragma unroll 4
    for (int i = tid; i < n16; i += 192) {
        const float4 v = *reinterpret_cast<const float4*>(lo + 16 * i);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(hi + 16 * i) = h;
        *reinterpret_cast<float4*>(lo + 16 * i) = l;
    }
}
__device__ __forceinline__ bool m_in_range(int m, int M) { return m < M; }
// Base of epilogue functors: store4(b, m, n, float4) for 4 consecutive columns of one row; functors with kAux = true also
// provide aux4(b, m, n, M, N) (a second operand, loaded ahead), prefetch_line(b, m, n) and store4(b, m, n, v, aux).
struct EpNoAux {
    static constexpr bool kAux = false;
};

// Persistent: grid = min(#tiles, #SMs) CTAs, each walking the tile list t = blockIdx.x + i gridDim.x with n fastest, then m, then
// batch: the 28 tiles of a 784 x 784 sample run side by side on 28 SMs and request the same operand lines at the same time, which
// the L2 merges.  (Measured alternative: n slowest — balances the narrow tail tiles of N over all CTAs, 25 % fewer full tiles per
// busy CTA — is 5-15 % SLOWER: the kernel is bound by the chip-wide L2 request rate, not by per-SM work, so operand sharing in
// time beats load balance.  Starting every second group of 28 CTAs 3-10 us late, so that their store-bound epilogues would meet the
// others' L2-bound main loops, changed nothing either: profiles/r02_s_stagger.txt.)
// block 384; dynamic smem SMEM_BYTES; one CTA per SM (all 512 TMEM columns: columns [0,256) accumulate hi*hi, [256,512) the
// 2^-11-smaller correction products — the tensor core truncates addends to the accumulator's exponent at every accumulate
// step, so folding the 2 K/8 correction steps into the large accumulator triples its truncation bias: measured 3.0e-6 vs
// 1.2e-6 of the result's scale at K = 256, and past the path's 1e-5 bar once two contractions chain; the sums meet once, in
// fp32, in the epilogue).  The producer runs ahead of the MMA warp across tiles, so while the epilogue warps drain the
// accumulators the ring already holds the next tile's first chunks; barriers: full/empty per stage, acc_full/acc_empty.
// BLO = false: the B operand is exact in TF32 (e.g. a 0/1 mask) — its lo plane is neither loaded nor multiplied.
template <int TK, bool BLO, class EP>
__global__ void __launch_bounds__(THREADS, 1) tc2_gemm_kernel(const __grid_constant__ Maps maps, int M, int N, int K, int K1, int batch, int kb, int a_bcast, int raw, EP ep) {
    using C = Cfg<TK>;
    constexpr int STAGES = C::kStages;
    extern __shared__ uint8_t tc2_smem_raw[];
    __shared__ uint64_t full[STAGES], empty[STAGES], conv[STAGES], acc_full, acc_empty;
    __shared__ uint32_t tmem_slot;
    // raw & 1 (A) / raw & 2 (B): the operand comes as ONE fp32 plane (the lo map); TMA lands it in the stage's lo slot and the
    // converter warps (2, 3) split it there: hi -> the hi slot, lo back in place (same swizzled offsets in both slots).
    // raw & 4 / raw & 8: that fp32 plane is MN-major ([K][rows]) and is staged / described accordingly (desc_mnmajor).
    const bool a_raw = raw & 1, b_raw = raw & 2, a_mn = raw & 4, b_mn = raw & 8, any_raw = raw & 3;
    const uint32_t base = (tc::smem_u32(tc2_smem_raw) + 1023u) & ~1023u;
    float* epi = reinterpret_cast<float*>(tc2_smem_raw + (base - tc::smem_u32(tc2_smem_raw)) + RING_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = (M + TM - 1) / TM, NT = (N + TN - 1) / TN;
    // kb > 1 ("batch as K"): one output tile sums the products of kb consecutive batch entries, C[g] = sum_{b in group g} A[b] B[b]^T
    // (the weight gradient of the 1x1 convolution: K = the pixels of a sample, summed over samples); set 1 is not used then.
    const int ntiles = ((batch + kb - 1) / kb) * MT * NT;
    const int nchunk0 = (K + TK - 1) / TK;
    const int nchunk = kb * nchunk0 + (K1 + TK - 1) / TK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); tc::mbar_init(&conv[s], NCONV / 32); }
        tc::mbar_init(&acc_full, 1);
        tc::mbar_init(&acc_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) { prefetch_map(&maps.a_hi[0]); prefetch_map(&maps.a_lo[0]); prefetch_map(&maps.b_hi[0]); prefetch_map(&maps.b_lo[0]); }
    if (warp == 2) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            int g = 0;  // chunks issued so far by this CTA (ring position)
#pragma unroll 1
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int nt = t % NT, mt = (t / NT) % MT, bg = t / (NT * MT);
#pragma unroll 1
                for (int c = 0; c < nchunk; c++, g++) {
                    const int s = g % STAGES, round = g / STAGES;
                    if (round > 0) tc::mbar_wait(&empty[s], (round - 1) & 1);  // the MMAs that read this stage have completed
                    const uint32_t st = base + s * C::kStageBytes;
                    const int set = c >= kb * nchunk0 ? 1 : 0;
                    const int cc = set ? c - kb * nchunk0 : c;
                    const int k0 = (kb > 1 ? cc % nchunk0 : cc) * TK;
                    const int b = kb > 1 ? bg * kb + cc / nchunk0 : bg;  // a batch entry past the end reads as zeros (TMA fill)
                    // a ragged last MN-major B tile brings only the 32-row groups the MMA (N = tn) reads
                    const bool b_ragged = b_mn && nt * TN + TN > N;
                    const int b_grps = b_ragged ? (min(TN, N - nt * TN) + 31) / 32 : TN / 32;
                    const uint32_t b_plane = b_ragged ? (uint32_t)b_grps * (TK * 128) : C::kBTile;
                    mbar_arrive_expect_tx(&full[s], (a_raw ? 1 : 2) * C::kATile + ((BLO && !b_raw) ? 2 : 1) * b_plane);
                    if (!a_raw && !a_mn) tma_load_3d(st, &maps.a_hi[set], k0, mt * TM, a_bcast ? 0 : b, &full[s]);
                    if (a_mn) {  // boxes of 32 rows x TK k-lines
                        if (mt * TM + TM <= M) {
                            tma_load_4d(st + C::kATile, &maps.a_grp[set], 0, k0, mt * (TM / 32), a_bcast ? 0 : b, &full[s]);
                            if (!a_raw) tma_load_4d(st, &maps.a_grp_hi[set], 0, k0, mt * (TM / 32), a_bcast ? 0 : b, &full[s]);
                        } else {
#pragma unroll
                            for (int u = 0; u < TM / 32; u++) {
                                tma_load_3d(st + C::kATile + u * (TK * 128), &maps.a_lo[set], mt * TM + 32 * u, k0, a_bcast ? 0 : b, &full[s]);
                                if (!a_raw) tma_load_3d(st + u * (TK * 128), &maps.a_hi[set], mt * TM + 32 * u, k0, a_bcast ? 0 : b, &full[s]);
                            }
                        }
                    } else {
                        tma_load_3d(st + C::kATile, &maps.a_lo[set], k0, mt * TM, a_bcast ? 0 : b, &full[s]);
                    }
                    if (b_mn) {
                        if (nt * TN + TN <= N) {
                            tma_load_4d(st + 2 * C::kATile + C::kBTile, &maps.b_grp[set], 0, k0, nt * (TN / 32), b, &full[s]);
                            if (!b_raw) tma_load_4d(st + 2 * C::kATile, &maps.b_grp_hi[set], 0, k0, nt * (TN / 32), b, &full[s]);
                        } else {
#pragma unroll
                            for (int u = 0; u < TN / 32; u++)
                                if (u < b_grps) {
                                    tma_load_3d(st + 2 * C::kATile + C::kBTile + u * (TK * 128), &maps.b_lo[set], nt * TN + 32 * u, k0, b, &full[s]);
                                    if (!b_raw) tma_load_3d(st + 2 * C::kATile + u * (TK * 128), &maps.b_hi[set], nt * TN + 32 * u, k0, b, &full[s]);
                                }
                        }
                    } else if (b_raw) {
                        tma_load_3d(st + 2 * C::kATile + C::kBTile, &maps.b_lo[set], k0, nt * TN, b, &full[s]);
                    } else {
                        tma_load_3d(st + 2 * C::kATile, &maps.b_hi[set], k0, nt * TN, b, &full[s]);
                        if (BLO) tma_load_3d(st + 2 * C::kATile + C::kBTile, &maps.b_lo[set], k0, nt * TN, b, &full[s]);
                    }
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
                const int tn = min(TN, ((N - nt * TN) + 15) & ~15);  // UMMA N of this tile
                const uint32_t idesc = tc::make_idesc(TM, tn) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);  // operand major-ness bits
                if (i > 0) tc::mbar_wait(&acc_empty, (i - 1) & 1);  // the epilogue has drained the accumulators
                tc::fence_after_sync();
#pragma unroll 1
                for (int c = 0; c < nchunk; c++, g++) {
                    const int s = g % STAGES, round = g / STAGES;
                    tc::mbar_wait(any_raw ? &conv[s] : &full[s], round & 1);
                    tc::fence_after_sync();
                    const uint32_t a_hi = base + s * C::kStageBytes, a_lo = a_hi + C::kATile, b_hi = a_hi + 2 * C::kATile, b_lo = b_hi + C::kBTile;
#pragma unroll
                    for (int kk = 0; kk < TK / 8; kk++) {  // one MMA consumes K = 8 fp32: 32 bytes of a K-major row, one atom of an MN-major tile
                        const uint32_t acc = (c > 0 || kk > 0) ? 1u : 0u;
                        const uint64_t dah = a_mn ? desc_mnmajor<TK>(a_hi + kk * 1024) : desc_kmajor<TK>(a_hi + kk * 32);
                        const uint64_t dal = a_mn ? desc_mnmajor<TK>(a_lo + kk * 1024) : desc_kmajor<TK>(a_lo + kk * 32);
                        const uint64_t dbh = b_mn ? desc_mnmajor<TK>(b_hi + kk * 1024) : desc_kmajor<TK>(b_hi + kk * 32);
                        tc::mma_tf32(tmem_d, dah, dbh, idesc, acc);
                        if (BLO) {
                            const uint64_t dbl = b_mn ? desc_mnmajor<TK>(b_lo + kk * 1024) : desc_kmajor<TK>(b_lo + kk * 32);
                            tc::mma_tf32(tmem_d + TN, dah, dbl, idesc, acc);
                        }
                        tc::mma_tf32(tmem_d + TN, dal, dbh, idesc, BLO ? 1u : acc);
                    }
                    tc::mma_commit(&empty[s]);  // frees the stage when these MMAs have read it
                }
                tc::mma_commit(&acc_full);      // every MMA of the tile has completed: accumulators final
            }
        }
        __syncwarp();
    } else if (warp < 4 || warp >= 12) {
        // ===== converter warps (2, 3, 12-15): hi / lo split of operands that arrive as one fp32 plane.  Two warps could not keep
        // up with two MN-major operands (24 KB per 768 tensor-pipe cycles: measured 35 % slower main loops); six can. =====
        if (any_raw) {
            const int tid = warp < 4 ? threadIdx.x - 64 : threadIdx.x - 12 * 32 + 64;  // 0..NCONV-1
            int g = 0;
#pragma unroll 1
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int nt = t % NT;
                const int b16 = (b_mn && nt * TN + TN > N) ? ((min(TN, N - nt * TN) + 31) / 32) * (TK * 128 / 16) : (int)(C::kBTile / 16);
#pragma unroll 1
                for (int c = 0; c < nchunk; c++, g++) {
                    const int s = g % STAGES, round = g / STAGES;
                    tc::mbar_wait(&full[s], round & 1);
                    uint8_t* st = tc2_smem_raw + (base - tc::smem_u32(tc2_smem_raw)) + s * C::kStageBytes;
                    if (a_raw) split_tile(st, st + C::kATile, C::kATile / 16, tid);
                    if (b_raw) split_tile(st + 2 * C::kATile, st + 2 * C::kATile + C::kBTile, b16, tid);
                    tc::fence_smem_to_async();  // generic-proxy writes -> visible to the tensor core's operand reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&conv[s]);
                }
            }
        }
    } else {
        // ===== epilogue (warps 4-11): thread t of warp w owns accumulator row 32 (w % 4) + t (a warp reads only its own TMEM lane
        // quarter); the two warps of a quarter alternate 16-column groups =====
        const int q = warp & 3, half = (warp - 4) >> 2;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
        int i = 0;
#pragma unroll 1
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
            const int nt = t % NT, mt = (t / NT) % MT, b = t / (NT * MT);
            const int n0 = nt * TN;
            const int tn = min(TN, ((N - n0) + 15) & ~15);
            // TMEM hands every thread 16 consecutive columns of ITS row: stored like that, one warp instruction touches 32
            // different rows (32 L1 wavefronts for 512 bytes; measured: the stores, not the MMAs, bounded the K = 256
            // contractions).  Each warp therefore transposes its 32 x 16 block through shared memory and stores it as
            // 8 rows x 64 contiguous bytes per instruction.
            float* tb = epi + (warp - 4) * (32 * EPI_STRIDE);
            const int rr = lane >> 2, c4 = (lane & 3) * 4;
            const int mrow = mt * TM + q * 32 + rr;  // + 8 u
            // An epilogue that needs a second operand per element (EP::kAux: the derivative mask of gS reads S) would wait out a
            // DRAM round trip per 16-column group.  Its tile is pulled into L2 while the MMAs of this tile still run (these
            // warps have nothing else to do then), and the loads of group j + 32 are issued before group j is processed.
            float4 aux[4], aux_next[4];
            if constexpr (EP::kAux) {
                if (m_in_range(mt * TM + q * 32 + lane, M))
                    for (int jj = 32 * half; jj < tn; jj += 64)  // 128-byte lines of this warp's rows
                        if (n0 + jj < N) ep.prefetch_line(b, mt * TM + q * 32 + lane, n0 + jj);
            }
            tc::mbar_wait(&acc_full, i & 1);
            tc::fence_after_sync();
            if constexpr (EP::kAux) {
#pragma unroll
                for (int u = 0; u < 4; u++) aux_next[u] = ep.aux4(b, mrow + 8 * u, n0 + 16 * half + c4, M, N);
            }
#pragma unroll 1
            for (int j = 16 * half; j < tn; j += 32) {
                float v[16], w[16];
                if constexpr (EP::kAux) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        aux[u] = aux_next[u];
                        if (j + 32 < tn) aux_next[u] = ep.aux4(b, mrow + 8 * u, n0 + j + 32 + c4, M, N);
                    }
                }
                tmem_ld16x2(lane_addr + j, lane_addr + TN + j, v, w);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    *reinterpret_cast<float4*>(tb + lane * EPI_STRIDE + 4 * u) =
                        make_float4(v[4 * u] + w[4 * u], v[4 * u + 1] + w[4 * u + 1], v[4 * u + 2] + w[4 * u + 2], v[4 * u + 3] + w[4 * u + 3]);
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int row = 8 * u + rr;
                    const float4 x = *reinterpret_cast<const float4*>(tb + row * EPI_STRIDE + c4);
                    const int m = mrow + 8 * u, n = n0 + j + c4;
                    if (m < M && n < N) {
                        if constexpr (EP::kAux) ep.store4(b, m, n, x, aux[u]);
                        else ep.store4(b, m, n, x);
                    }
                }
                __syncwarp();
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_d, 512);
}

// hi = the bits a TF32 tensor core keeps, lo = x - hi (exact); n % 4 == 0, 16-byte aligned
static __global__ void __launch_bounds__(256) split_kernel(const float4* __restrict__ x, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo) {
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
// MN-major plane [batch][K][rows] fp32 viewed as a 3-D tensor (rows fastest); box = 32 rows x tk k-lines x 1, 128-byte swizzle of
// 32-byte atoms
static inline bool make_mn_map(CUtensorMap* tm, const float* base, int64_t batch, int rows, int K, int tk, int pitch) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)K, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * K * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)tk, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the same MN-major plane viewed 4-D: dims {32 rows, K, complete 32-row groups, batch}; box = 32 x tk x groups_per_tile x 1
static inline bool make_mn_group_map(CUtensorMap* tm, const float* base, int64_t batch, int rows, int K, int tk, int groups_per_tile, int pitch) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const int groups = rows / 32;  // complete groups only: a box never straddles the end of a k-line
    if (groups < groups_per_tile) {  // no tile is complete: the map is never used; encode a valid dummy (the 3-D map's geometry)
        return make_mn_map(tm, base, batch, rows, K, tk, pitch);
    }
    cuuint64_t dims[4] = {32, (cuuint64_t)K, (cuuint64_t)groups, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, 128, (cuuint64_t)pitch * K * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)tk, (cuuint32_t)groups_per_tile, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// plane [batch][rows][K] fp32 viewed as a 3-D tensor (K fastest); box = tk x box_rows x 1, swizzle span = one box row
static inline bool make_plane_map(CUtensorMap* tm, const float* base, int64_t batch, int rows, int K, int box_rows, int tk, int pitch) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows * pitch * 4};
    cuuint32_t box[3] = {(cuuint32_t)tk, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               tk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

static inline bool applicable(const void* a, const void* b, const void* c, const void* d) {
    static const int off = [] { const char* e = getenv("PIXPRO_B200_TC2"); return (e && e[0] == '0') ? 1 : 0; }();
    return !off && a && c && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}

template <int TK, bool BLO, class EP>
static inline int launch_cfg(const char* what, const Maps& maps, int64_t batch, int kb, int M, int N, int K, int K1, int a_bcast, int raw, EP ep, cudaStream_t st) {
    auto kern = tc2_gemm_kernel<TK, BLO, EP>;
    static unsigned long long opted = 0;  // per template instantiation, one bit per device
    if (smem_opt_in(kern, (int)SMEM_BYTES, opted) != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%u B of shared memory) failed: %s", what, SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
        return PP_ERR_CUDA;
    }
    const int64_t ntiles = ((batch + kb - 1) / kb) * (int64_t)((M + TM - 1) / TM) * ((N + TN - 1) / TN);
    const unsigned grid = (unsigned)(ntiles < num_sms() ? ntiles : num_sms());
    PP_LAUNCH(what, st, kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, M, N, K, K1, (int)batch, kb, a_bcast, raw, ep));
    return check_launch(what);
}

// Planes: A_hi/A_lo [batch][M][K] (or [M][K] shared by every batch entry: a_bcast), B_hi/B_lo [batch][N][K].
// Returns -1 when not applicable (caller uses pp_tc.cuh's kernel).
// PIXPRO_B200_TC2: 0 = off, 1 (default) = 64-byte chunks x 4 stages, 2 = 128-byte chunks x 2 stages (A/B switch).
struct Operands {  // hi / lo planes of one product: A [batch][M][K] (or [M][K] with a_bcast), B [batch][N][K]
    const float *a_hi, *a_lo, *b_hi, *b_lo;  // an operand given as ONE fp32 plane (split in the kernel): hi = the plane, lo = nullptr
    int K;
    bool a_mn = false, b_mn = false;  // the operand's plane(s) are MN-major: memory [batch][K][rows] (rows contiguous) instead of [batch][rows][K]
    // allocated length (floats, a multiple of 4) of the contiguous dimension when it is padded: K for a K-major plane, rows for an
    // MN-major one; 0 = dense.  E.g. the [C, 49] maps of the 7x7 grid copied to a pitch of 52 floats: TMA needs 16-byte strides.
    int a_pitch = 0, b_pitch = 0;
};
template <bool BLO = true, class EP>
static inline int launch_tc2_sets(const char* what, int64_t batch, int M, int N, const Operands* sets, int nsets, EP ep, cudaStream_t st,
                                  bool a_bcast, int kb = 1) {
    if (batch > 65535 || nsets < 1 || nsets > 2 || kb < 1 || (kb > 1 && nsets != 1)) return -1;
    const bool a_raw = sets[0].a_lo == nullptr, b_raw = BLO && sets[0].b_lo == nullptr;
    const bool a_mn = sets[0].a_mn, b_mn = sets[0].b_mn;
    if (b_mn && !BLO) return -1;
    for (int i = 0; i < nsets; i++) {
        if ((sets[i].a_lo == nullptr) != a_raw || (BLO && (sets[i].b_lo == nullptr) != b_raw)) return -1;  // same form in both sets
        if (sets[i].a_mn != a_mn || sets[i].b_mn != b_mn) return -1;
        if (!applicable(sets[i].a_hi, sets[i].a_lo, sets[i].b_hi, sets[i].b_lo)) return -1;
        // 16-byte strides: the (possibly padded) contiguous dimension of every plane is a multiple of 4 floats
        const int ap = sets[i].a_pitch ? sets[i].a_pitch : (a_mn ? M : sets[i].K), bp = sets[i].b_pitch ? sets[i].b_pitch : (b_mn ? N : sets[i].K);
        if (ap % 4 != 0 || bp % 4 != 0 || ap < (a_mn ? M : sets[i].K) || bp < (b_mn ? N : sets[i].K)) return -1;
    }
    if (batch * (int64_t)((M + TM - 1) / TM) * ((N + TN - 1) / TN) >= (1ll << 31)) return -1;
    static const int mode = [] { const char* e = getenv("PIXPRO_B200_TC2"); return e ? atoi(e) : 1; }();
    const int tk = (mode == 2 && !a_mn && !b_mn) ? 32 : 16;
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    const int64_t abatch = a_bcast ? 1 : batch;
    for (int i = 0; i < 2; i++) {
        const Operands& o = sets[i < nsets ? i : 0];
        // raw operand: the single fp32 plane is what the "lo" map describes (it is loaded into the stage's lo slot)
        const float* a_lo_src = a_raw ? o.a_hi : o.a_lo;                  // what the stage's lo slot receives
        const float* b_lo_src = (BLO && !b_raw) ? o.b_lo : o.b_hi;
        const int ap = o.a_pitch ? o.a_pitch : (a_mn ? M : o.K), bp = o.b_pitch ? o.b_pitch : (b_mn ? N : o.K);
        bool ok = a_mn ? (make_mn_map(&maps.a_lo[i], a_lo_src, abatch, M, o.K, tk, ap) && make_mn_map(&maps.a_hi[i], o.a_hi, abatch, M, o.K, tk, ap) &&
                          make_mn_group_map(&maps.a_grp[i], a_lo_src, abatch, M, o.K, tk, TM / 32, ap) &&
                          make_mn_group_map(&maps.a_grp_hi[i], o.a_hi, abatch, M, o.K, tk, TM / 32, ap))
                       : (make_plane_map(&maps.a_hi[i], o.a_hi, abatch, M, o.K, TM, tk, ap) && make_plane_map(&maps.a_lo[i], a_lo_src, abatch, M, o.K, TM, tk, ap));
        ok = ok && (b_mn ? (make_mn_map(&maps.b_lo[i], b_lo_src, batch, N, o.K, tk, bp) && make_mn_map(&maps.b_hi[i], o.b_hi, batch, N, o.K, tk, bp) &&
                            make_mn_group_map(&maps.b_grp[i], b_lo_src, batch, N, o.K, tk, TN / 32, bp) &&
                            make_mn_group_map(&maps.b_grp_hi[i], o.b_hi, batch, N, o.K, tk, TN / 32, bp))
                         : (make_plane_map(&maps.b_hi[i], o.b_hi, batch, N, o.K, TN, tk, bp) && make_plane_map(&maps.b_lo[i], b_lo_src, batch, N, o.K, TN, tk, bp)));
        if (!ok) return -1;
    }
    const int K0 = sets[0].K, K1 = nsets > 1 ? sets[1].K : 0;
    const int raw = (a_raw ? 1 : 0) | (b_raw ? 2 : 0) | (a_mn ? 4 : 0) | (b_mn ? 8 : 0);
    return tk == 32 ? launch_cfg<32, BLO>(what, maps, batch, kb, M, N, K0, K1, a_bcast ? 1 : 0, raw, ep, st)
                    : launch_cfg<16, BLO>(what, maps, batch, kb, M, N, K0, K1, a_bcast ? 1 : 0, raw, ep, st);
}

// Planes: A_hi/A_lo [batch][M][K] (or [M][K] shared by every batch entry: a_bcast), B_hi/B_lo [batch][N][K].
// Returns -1 when not applicable (caller uses pp_tc.cuh's kernel).
// PIXPRO_B200_TC2: 0 = off, 1 (default) = 64-byte chunks x 4 stages, 2 = 128-byte chunks x 2 stages (A/B switch).
template <bool BLO = true, class EP>
static inline int launch_tc2(const char* what, int64_t batch, int M, int N, int K, const float* a_hi, const float* a_lo, const float* b_hi,
                             const float* b_lo, EP ep, cudaStream_t st, bool a_bcast = false) {
    const Operands o{a_hi, a_lo, b_hi, b_lo, K};
    return launch_tc2_sets<BLO>(what, batch, M, N, &o, 1, ep, st, a_bcast);
}

static inline int launch_split(const float* x, int64_t n, float* hi, float* lo, cudaStream_t st) {
    const int64_t n4 = n / 4;
    PP_LAUNCH("tc split", st, split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n4,
                                                                                        reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo)));
    return check_launch("tc split");
}

}  // namespace tc2
}  // namespace pp
