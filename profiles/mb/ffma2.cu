#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    if (MODE == 0) {
        for (int i = 0; i < iters; i++) {
            a0 = __fmaf_rn(a0, s, a1); a1 = __fmaf_rn(a1, s, a2); a2 = __fmaf_rn(a2, s, a3); a3 = __fmaf_rn(a3, s, a4);
            a4 = __fmaf_rn(a4, s, a5); a5 = __fmaf_rn(a5, s, a6); a6 = __fmaf_rn(a6, s, a7); a7 = __fmaf_rn(a7, s, a0);
        }
    } else {
        u64 p0, p1, p2, p3, ps;
        asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a0), "f"(a1)); asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a2), "f"(a3));
        asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a4), "f"(a5)); asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a6), "f"(a7));
        asm("mov.b64 %0, {%1,%2};" : "=l"(ps) : "f"(s), "f"(s));
        for (int i = 0; i < iters; i++) { p0 = fma2(p0, ps, p1); p1 = fma2(p1, ps, p2); p2 = fma2(p2, ps, p3); p3 = fma2(p3, ps, p0); }
        asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0)); asm("mov.b64 {%0,%1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2)); asm("mov.b64 {%0,%1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    float* o; cudaMalloc(&o, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 2; mode++) for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 1024>>>(o, iters, 0.999f); else k<1><<<148 * 8, 1024>>>(o, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 148.0 * 8 * 1024 * iters * 8 * 2;
        printf("mode %d: %.3f ms  %.1f TFLOP/s fp32 (%s)\n", mode, ms, flops / ms / 1e9, mode ? "FFMA2, 4 per iter" : "FFMA, 8 per iter");
    }
    return 0;
}
