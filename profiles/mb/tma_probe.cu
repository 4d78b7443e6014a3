// Probe: one 3-D TMA box load (W x H x planes fp32 tensor, box BW x BH x 2) incl. out-of-frame coordinates.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BW = 48, BH = 48;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1, int sel, int ox, int oy, int pl, float* out, int* status) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* tm = sel ? &tm0 : &tm1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(2 * BW * BH * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(ox), "r"(oy), "r"(pl), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) *status = done;
    if (!done) return;
    const float* s = (const float*)sm;
    for (int i = threadIdx.x; i < 2 * BW * BH; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int W = 1280, H = 720, P = 8;
    float* d; cudaMalloc(&d, (size_t)W * H * P * 4);
    float* h = (float*)malloc((size_t)W * H * P * 4);
    for (size_t i = 0; i < (size_t)W * H * P; i++) h[i] = (float)(i % 1000003);
    cudaMemcpy(d, h, (size_t)W * H * P * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: err=%d q=%d p=%p\n", (int)e, (int)q, p);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, P}; cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {BW, BH, 2}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    float* out; int* st; cudaMalloc(&out, 2 * BW * BH * 4); cudaMalloc(&st, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BW * BH * 4);
    int cases[6][3] = {{100, 200, 2}, {-5, -7, 0}, {1260, 700, 6}, {13, 31, 4}, {-5, 40, 0}, {40, -7, 2}};
    float* ho = (float*)malloc(2 * BW * BH * 4);
    for (int c = 0; c < 6; c++) {
        if (only >= 0 && c != only) continue;
        int ox = cases[c][0], oy = cases[c][1], pl = cases[c][2];
        k<<<1, 128, 2 * BW * BH * 4>>>(tm, tm, c & 1, ox, oy, pl, out, st);
        cudaError_t er = cudaDeviceSynchronize();
        int hs = -1; cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(ho, out, 2 * BW * BH * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int ch = 0; ch < 2; ch++) for (int y = 0; y < BH; y++) for (int x = 0; x < BW; x++) {
            int gx = ox + x, gy = oy + y;
            float ref = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[((size_t)(pl + ch) * H + gy) * W + gx];
            if (ho[(ch * BH + y) * BW + x] != ref) bad++;
        }
        printf("case %d (%d,%d,%d): sync=%s done=%d mismatches=%ld\n", c, ox, oy, pl, cudaGetErrorString(er), hs, bad);
    }
    return 0;
}
