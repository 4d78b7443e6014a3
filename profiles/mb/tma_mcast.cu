// tma_mcast.cu — does TMA multicast inside a 2-CTA cluster raise the operand bandwidth a GEMM main loop sees?
// Every CTA streams K chunks of an A tile (128 rows x 64 B) and a B tile (256 rows x 64 B) through a 4-stage ring, as
// pp_tc2.cuh's producer does (no MMA: a consumer thread releases each stage as soon as it is full).
//   mode 0: each CTA loads its A tile and the whole B tile itself                       (48 KB from L2 per CTA and chunk)
//   mode 1: the two CTAs of a cluster share the B tile: each loads half of it with .multicast::cluster to both
//           (A 16 KB + B 16 KB = 32 KB from L2 per CTA and chunk, 48 KB delivered)
// Reports delivered bytes / clk / SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_mcast tma_mcast.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#ifndef KFLOATS
#define KFLOATS 16
#endif
constexpr int KF = KFLOATS;                      // K chunk: 16 floats = 64-byte rows (SWIZZLE_64B), 32 = 128-byte rows
constexpr int MAXST = KF == 16 ? 4 : 2;
constexpr int AROWS = 128, BROWS = 256;
constexpr uint32_t A_BYTES = AROWS * KF * 4, B_BYTES = BROWS * KF * 4, STAGE_BYTES = A_BYTES + B_BYTES;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24) && !done; spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
k(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb, const __grid_constant__ CUtensorMap tmbh, int mode,
  int iters, int rows_total, int kchunks, int* fail, int STAGES) {
    extern __shared__ __align__(1024) uint8_t ring[];
    __shared__ uint64_t full[MAXST], empty[MAXST];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode ? 2 : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync();
    const uint32_t base = smem_u32(ring);
    // tiles: cluster c works on A row block (2c + rank) and B row block c (256 rows), all K chunks, repeated
    const int cl = blockIdx.x >> 1;
    const int arow = ((2 * cl + rank) * AROWS) % (rows_total - AROWS), brow = (cl * BROWS) % (rows_total - BROWS);
    if (warp == 0 && lane == 0) {  // producer
        for (int g = 0; g < iters; g++) {
            const int s = g % STAGES, round = g / STAGES;
            if (round > 0 && !mbar_wait(&empty[s], (round - 1) & 1)) { atomicAdd(fail, 1); break; }
            const uint32_t st = base + s * STAGE_BYTES;
            const int k0 = (g % kchunks) * KF;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(STAGE_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(st), "l"(reinterpret_cast<uint64_t>(&tma)), "r"(k0), "r"(arow), "r"(smem_u32(&full[s])) : "memory");
            if (mode == 0) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(st + A_BYTES), "l"(reinterpret_cast<uint64_t>(&tmb)), "r"(k0), "r"(brow), "r"(smem_u32(&full[s])) : "memory");
            } else {  // my half of the B tile, to the same offset of both CTAs; each CTA's barrier receives both halves' bytes
                const uint16_t mask = 3;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                             ::"r"(st + A_BYTES + rank * (B_BYTES / 2)), "l"(reinterpret_cast<uint64_t>(&tmbh)), "r"(k0), "r"(brow + (int)rank * (BROWS / 2)),
                               "r"(smem_u32(&full[s])), "h"(mask) : "memory");
            }
        }
    } else if (warp == 1 && lane == 0) {  // consumer: release the stage in every CTA that writes into it
        for (int g = 0; g < iters; g++) {
            const int s = g % STAGES, round = g / STAGES;
            if (!mbar_wait(&full[s], round & 1)) { atomicAdd(fail, 1); break; }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
            if (mode) {
                uint32_t raddr;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(&empty[s])), "r"(rank ^ 1u));
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
            }
        }
    }
    __syncthreads();
    cluster_sync();  // no CTA leaves while its peer may still arrive on its barriers / write its shared memory
}

int main(int argc, char** argv) {
    const int STAGES = argc > 1 ? (atoi(argv[1]) < MAXST ? atoi(argv[1]) : MAXST) : MAXST;   // ring depth
    const int grid_arg = argc > 2 ? atoi(argv[2]) : 0;  // CTAs (0 = one per SM)
    const int ROWS = 32768, K = 256;  // 32 MB fp32 matrix: L2-resident after the first pass
    float* d; cudaMalloc(&d, (size_t)ROWS * K * 4); cudaMemset(d, 0, (size_t)ROWS * K * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    auto mk = [&](CUtensorMap* tm, int box_rows) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)ROWS}; cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {KF, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
        return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   KF == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap tma, tmb, tmbh;
    if (mk(&tma, AROWS) || mk(&tmb, BROWS) || mk(&tmbh, BROWS / 2)) { printf("encode failed\n"); return 1; }
    int* fail; cudaMalloc(&fail, 4); cudaMemset(fail, 0, 4);
    const int smem = MAXST * STAGE_BYTES + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = (grid_arg ? grid_arg : sms) & ~1, iters = 4096;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k<<<grid, 64, smem>>>(tma, tmb, tmbh, mode, iters, ROWS, K / KF, fail, STAGES);
            cudaEventRecord(e1);
            cudaError_t er = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            int hf = 0; cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost);
            const double delivered = (double)iters * STAGE_BYTES, clk = ms * 1e-3 * 1.965e9;
            const double from_l2 = (double)iters * (mode ? A_BYTES + B_BYTES / 2 : STAGE_BYTES);
            if (rep) printf("mode %d (%s): %s fail=%d  %.3f ms  delivered %.1f B/clk/SM  read from L2 %.1f B/clk/SM  (%d CTAs, %d stages)\n", mode,
                            mode ? "B tile multicast inside 2-CTA clusters" : "every CTA loads everything", cudaGetErrorString(er), hf, ms,
                            delivered / clk, from_l2 / clk, grid, STAGES);
        }
    }
    return 0;
}
