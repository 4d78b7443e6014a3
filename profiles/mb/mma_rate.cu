// Ceiling of the tcgen05 kind::tf32 issue pattern used by pp_tc.cuh: back-to-back 128x128x8 MMAs on operands that are
// already in shared memory (no staging at all), same no-swizzle K-major descriptors, 1 or 2 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pixpro-with-opticalflow_b200/csrc/pp_tc.cuh"
using namespace pp::tc;
__global__ void __launch_bounds__(128) k(int iters, float* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * STAGE_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < 2 * (int)STAGE_BYTES / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f / (1 + (i % 7));
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) tmem_alloc(slot, 2 * TN);
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(TM, TN);
        for (int c = 0; c < iters; c++) {
            const uint32_t a_hi = smem_u32(sm + (c & 1) * STAGE_BYTES), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
#pragma unroll
            for (int kk = 0; kk < TK / 8; kk++) {
                const uint32_t ko = kk * 2 * LBO;
                mma_tf32(tmem_d, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, 1u);
                mma_tf32(tmem_d + TN, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
                mma_tf32(tmem_d + TN, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, 1u);
            }
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    fence_after_sync();
    float v[16];
    tmem_ld16(tmem_d + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = v[0];
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_d, 2 * TN);
}
int main() {
    float* o; cudaMalloc(&o, 148 * 2 * 128 * 4);
    const int smem = 2 * STAGE_BYTES + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int per_sm = 1; per_sm <= 2; per_sm++) for (int rep = 0; rep < 2; rep++) {
        const int iters = 4000;
        cudaEventRecord(e0);
        k<<<148 * per_sm, 128, smem>>>(iters, o);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 148.0 * per_sm * iters * (TK / 8) * 3 * 2.0 * TM * TN * 8;
        printf("%d CTA/SM: %.3f ms, %.1f TFLOP/s of tf32 MMA work (%.1f useful 3xTF32)  err=%s\n", per_sm, ms, flops / ms / 1e9, flops / ms / 3e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
