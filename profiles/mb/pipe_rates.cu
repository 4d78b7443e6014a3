// pipe_rates.cu — issue / pipe throughput of the instruction classes the flow kernels are made of, on
// sm_100a (B200).  Each test runs ITERS iterations of 8 independent dependency chains of one instruction
// class (or an interleaved mix) on 148*4 CTAs x 512 threads and reports warp-instructions per clock per
// SM sub-partition (SMSP).  1.0 = one issue slot per clock; 0.5 = the pipe accepts a warp instruction
// every other clock.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

enum { T_FFMA, T_FFMA2, T_FADD2, T_FMUL2, T_FADD, T_FMUL, T_IMAD, T_IADD3, T_LOP3, T_F2I, T_I2F, T_FSETP_SEL, T_LDS,
       T_MIX_FFMA2_IADD3, T_MIX_FFMA2_FADD, T_MIX_FFMA2_LDS, T_MIX_FFMA_IADD3, T_MIX_FFMA2_I2F, T_FADD2_RM, T_MIX_FFMA2_FFMA,
       T_IMNMX, T_LEA, T_MIX3, T_COUNT };
static const char* NAMES[T_COUNT] = {"FFMA (3 reg)", "FFMA2", "FADD2", "FMUL2", "FADD", "FMUL", "IMAD", "IADD3", "LOP3", "F2I.FLOOR", "I2FP",
                                     "FSETP+SEL", "LDS.32", "FFMA2 + IADD3 (1:1)", "FFMA2 + FADD (1:1)", "FFMA2 + LDS (1:1)",
                                     "FFMA + IADD3 (1:1)", "FFMA2 + I2FP (1:1)", "FADD2.RM", "FFMA2 + FFMA (1:1)", "VIMNMX", "LEA",
                                     "FFMA2 + IADD3 + LDS (1:1:1)"};
static const int PER_ITER[T_COUNT] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 16, 8, 16, 16, 16, 16, 16, 8, 16, 8, 8, 24};

template <int T>
__global__ void __launch_bounds__(512) k(float* out, int iters, float s, int si) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = s;
    sm[threadIdx.x + 512] = s;
    __syncthreads();
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    u64 p0, p1, p2, p3, p4, p5, p6, p7, ps;
#define PK(n) asm("mov.b64 %0, {%1,%2};" : "=l"(p##n) : "f"(a##n), "f"(a##n + 0.5f));
    REP8(PK)
    asm("mov.b64 %0, {%1,%2};" : "=l"(ps) : "f"(s), "f"(s));
    unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    for (int it = 0; it < iters; it++) {
#define FFMA(n) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a##n) : "f"(s));
#define FFMA2(n) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p##n) : "l"(ps));
#define FADD2(n) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p##n) : "l"(ps));
#define FADD2RM(n) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(p##n) : "l"(ps));
#define FMUL2(n) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p##n) : "l"(ps));
#define FADD(n) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a##n) : "f"(s));
#define FMUL(n) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a##n) : "f"(s));
#define IMAD(n) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(i##n) : "r"(si));
#define IADD3(n) asm volatile("add.s32 %0, %0, %1;" : "+r"(i##n) : "r"(si));
#define LOP3(n) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i##n) : "r"(si));
#define F2I(n) asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(i##n) : "f"(a##n));
#define I2F(n) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a##n) : "r"(i##n));
#define FSETPSEL(n) asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %1, %0, q;}" : "+f"(a##n) : "f"(s));
#define LDS(n) asm volatile("ld.shared.f32 %0, [%1+" #n "*128];" : "=f"(a##n) : "r"(sa));
#define IMNMX(n) asm volatile("min.s32 %0, %0, %1;" : "+r"(i##n) : "r"(si));
#define LEA(n) asm volatile("{.reg .b32 t; shl.b32 t, %0, 2; add.s32 %0, t, %1;}" : "+r"(i##n) : "r"(si));
        if (T == T_FFMA) { REP8(FFMA) }
        if (T == T_FFMA2) { REP8(FFMA2) }
        if (T == T_FADD2) { REP8(FADD2) }
        if (T == T_FADD2_RM) { REP8(FADD2RM) }
        if (T == T_FMUL2) { REP8(FMUL2) }
        if (T == T_FADD) { REP8(FADD) }
        if (T == T_FMUL) { REP8(FMUL) }
        if (T == T_IMAD) { REP8(IMAD) }
        if (T == T_IADD3) { REP8(IADD3) }
        if (T == T_LOP3) { REP8(LOP3) }
        if (T == T_F2I) { REP8(F2I) }
        if (T == T_I2F) { REP8(I2F) }
        if (T == T_FSETP_SEL) { REP8(FSETPSEL) }
        if (T == T_LDS) { REP8(LDS) }
        if (T == T_IMNMX) { REP8(IMNMX) }
        if (T == T_LEA) { REP8(LEA) }
#define MIX_A(n) FFMA2(n) IADD3(n)
#define MIX_B(n) FFMA2(n) FADD(n)
#define MIX_C(n) FFMA2(n) LDS(n)
#define MIX_D(n) FFMA(n) IADD3(n)
#define MIX_E(n) FFMA2(n) I2F(n)
#define MIX_F(n) FFMA2(n) FFMA(n)
#define MIX_G(n) FFMA2(n) IADD3(n) LDS(n)
        if (T == T_MIX_FFMA2_IADD3) { REP8(MIX_A) }
        if (T == T_MIX_FFMA2_FADD) { REP8(MIX_B) }
        if (T == T_MIX_FFMA2_LDS) { REP8(MIX_C) }
        if (T == T_MIX_FFMA_IADD3) { REP8(MIX_D) }
        if (T == T_MIX_FFMA2_I2F) { REP8(MIX_E) }
        if (T == T_MIX_FFMA2_FFMA) { REP8(MIX_F) }
        if (T == T_MIX3) { REP8(MIX_G) }
    }
    float b0, b1;
    float acc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7);
#define UNPK(n) asm("mov.b64 {%0,%1}, %2;" : "=f"(b0), "=f"(b1) : "l"(p##n)); acc += b0 + b1;
    REP8(UNPK)
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int T>
void run(float* o, int iters, double mhz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas = 148 * 4;  // 4 CTAs x 16 warps = 64 warps per SM, 16 per SMSP
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<T><<<ctas, 512>>>(o, iters, 0.999f, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // warp instructions per SMSP = ctas/148 CTAs per SM * 16 warps / 4 SMSPs * iters * PER_ITER
    const double wi = (double)ctas / 148 * 16 / 4 * iters * PER_ITER[T];
    const double clocks = best * 1e-3 * mhz * 1e6;
    printf("%-28s %8.3f ms   %.3f warp-instr/clk/SMSP (at %.0f MHz)\n", NAMES[T], best, wi / clocks, mhz);
}

int main(int argc, char** argv) {
    float* o;
    cudaMalloc(&o, 148 * 4 * 512 * 4);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double mhz = argc > 1 ? atof(argv[1]) : clk_khz / 1000.0;
    const int iters = 20000;
    run<T_FFMA>(o, iters, mhz); run<T_FFMA2>(o, iters, mhz); run<T_FADD2>(o, iters, mhz); run<T_FADD2_RM>(o, iters, mhz);
    run<T_FMUL2>(o, iters, mhz); run<T_FADD>(o, iters, mhz); run<T_FMUL>(o, iters, mhz); run<T_IMAD>(o, iters, mhz);
    run<T_IADD3>(o, iters, mhz); run<T_LOP3>(o, iters, mhz); run<T_IMNMX>(o, iters, mhz); run<T_LEA>(o, iters, mhz);
    run<T_F2I>(o, iters, mhz); run<T_I2F>(o, iters, mhz);
    run<T_FSETP_SEL>(o, iters, mhz); run<T_LDS>(o, iters, mhz);
    run<T_MIX_FFMA2_IADD3>(o, iters, mhz); run<T_MIX_FFMA2_FADD>(o, iters, mhz); run<T_MIX_FFMA2_LDS>(o, iters, mhz);
    run<T_MIX_FFMA_IADD3>(o, iters, mhz); run<T_MIX_FFMA2_I2F>(o, iters, mhz); run<T_MIX_FFMA2_FFMA>(o, iters, mhz);
    run<T_MIX3>(o, iters, mhz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
