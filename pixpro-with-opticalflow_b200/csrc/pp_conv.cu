// pp_conv.cu — the PPM's value transform (1x1 convolution, contrast/models/PixPro.py:21-23,
// :300, applied at :343) forward and backward on the tcgen05 3xTF32 kernel of pp_tc.cuh.
//
// SURVEY.md §8(f) rank 3 ("value_transform fused into the PPM prologue"): at the benchmark size
// (2B = 128 samples, 256 -> 256 channels, 7x7 grid) cuDNN's fp32 conv + its weight/bias gradient
// kernels cost ~190 us per step — more than every PPM/loss kernel together — while the three
// contractions are only 2.5 GFLOP.  Here each is ONE tensor-core GEMM over the joint (sample,
// pixel) index n = b*P + p, read straight from the [B,C,P] layout (no im2col, no transposes):
//     y [o][n] = Σ_c W[o][c] x[c][n] + bias[o]            M=Cout N=B*P K=Cin
//     dx[c][n] = Σ_o W[o][c] dy[o][n]                      M=Cin  N=B*P K=Cout
//     dW[o][c] = Σ_n dy[o][n] x[c][n]                      M=Cout N=Cin K=B*P  (split-K, deterministic reduce)
//     db[o]    = Σ_n dy[o][n]
// fp32-accurate (3xTF32, ~1e-6 relative), checked against torch in tests/test_gpu_tc.py.
#include "pp_common.cuh"
#include "pp_tc.cuh"
#include "pp_tc2.cuh"

namespace pp {

// joint index n -> (b, p) without an integer division per element: P is small, use float reciprocal
// with an exact fix-up (n < 2^24).
struct DivP {
    int P;
    float inv;
    __device__ __forceinline__ void operator()(int n, int& b, int& p) const {
        b = (int)((float)n * inv);
        p = n - b * P;
        if (p < 0) { b -= 1; p += P; }
        if (p >= P) { b += 1; p -= P; }
    }
};
static DivP make_divp(int P) { return DivP{P, 1.0f / (float)P}; }

// operand row = joint index n, k = channel: t[b][k][p]  (x or dy viewed as [N][C], "transposed")
struct LdNC {
    static constexpr bool kRowMajorK = false;
    const float* t;
    int C, NP;  // channels, B*P
    DivP dp;
    __device__ __forceinline__ float4 load4(int64_t, int n, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < NP) {
            int b, p;
            dp(n, b, p);
            const float* q = t + ((int64_t)b * C + k) * dp.P + p;
            if (k < C) v.x = __ldg(q);
            if (k + 1 < C) v.y = __ldg(q + dp.P);
            if (k + 2 < C) v.z = __ldg(q + 2 * dp.P);
            if (k + 3 < C) v.w = __ldg(q + 3 * dp.P);
        }
        return v;
    }
};
// operand row = channel, k = joint index n of split `s`: t[b][row][p], n = s*KS + k
struct LdCN {
    static constexpr bool kRowMajorK = true;
    const float* t;
    int C, NP, KS;
    DivP dp;
    __device__ __forceinline__ float4 load4(int64_t s, int row, int k) const {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (row < C && k < KS) {
            const int n0 = (int)s * KS + k;
            int b, p;
            dp(n0, b, p);  // one (sample, pixel) split per item; the next three joint indices follow by increment
            const float* q = t + ((int64_t)b * C + row) * dp.P + p;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k + u < KS && n0 + u < NP) v[u] = __ldg(q);
                ++q;
                if (++p == dp.P) { p = 0; q += (int64_t)(C - 1) * dp.P; }  // row `row` of the next sample
            }
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
};
// W [Cout][Cin]: rows o, k = c (natural)  /  rows c, k = o (transposed)
struct LdW {
    static constexpr bool kRowMajorK = true;
    const float* w;
    int rows, K;
    __device__ __forceinline__ float4 load4(int64_t, int row, int k) const {
        if (row >= rows || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg4_guard(w + (int64_t)row * K + k, k, K);
    }
};
struct LdWT {
    static constexpr bool kRowMajorK = false;
    const float* w;
    int Cout, Cin;  // operand rows = c (Cin), k = o (Cout)
    __device__ __forceinline__ float4 load4(int64_t, int c, int o) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cin) {
            const float* q = w + (int64_t)o * Cin + c;
            if (o < Cout) v.x = __ldg(q);
            if (o + 1 < Cout) v.y = __ldg(q + Cin);
            if (o + 2 < Cout) v.z = __ldg(q + 2 * Cin);
            if (o + 3 < Cout) v.w = __ldg(q + 3 * Cin);
        }
        return v;
    }
};
// out[b][m][p] (+ bias[m]) for the 16 joint indices n..n+15
struct StCN {
    float* out;
    const float* bias;  // may be null
    int C, NP;
    DivP dp;
    __device__ __forceinline__ void store16(int64_t, int m, int n, const float v[16]) const {
        const float bv = bias ? __ldg(bias + m) : 0.0f;
        int b, p;
        dp(n, b, p);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (n + u < NP) out[((int64_t)b * C + m) * dp.P + p] = v[u] + bv;
            if (++p == dp.P) { p = 0; b++; }
        }
    }
};

// dW[o][c] = Σ_s part[s][o][c];  db[o] = Σ_n dy[o][n]   (fixed summation order)
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ part, int splits, int total,
                                                                 float* __restrict__ dw) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    float s = 0.0f;
    for (int k = 0; k < splits; k++) s += __ldg(part + (int64_t)k * total + e);
    dw[e] = s;
}
// db[o] = Σ_b Σ_p dy[b][o][p] in two deterministic passes: one warp per (b, o) row (coalesced, B*C warps in
// flight), then one thread per channel sums the B row sums in order.  (One block per channel looping over B*P
// elements took 0.36 ms at B=64, 28x28 and sat on the backward's critical path.)
__global__ void __launch_bounds__(256) conv_bias_rows_kernel(const float* __restrict__ dy, int rows, int P,
                                                              float* __restrict__ part) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* q = dy + (int64_t)row * P;
    float s = 0.0f;
    for (int p = lane; p < P; p += 32) s += __ldg(q + p);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[row] = s;
}
// block = 32 channels x 8 row groups: warp w sums rows b = w, w + 8, ... of channel 32 blockIdx.x + lane, then the eight
// partial sums are added in warp order (deterministic)
__global__ void __launch_bounds__(256) conv_bias_grad_kernel(const float* __restrict__ part, int B, int C, float* __restrict__ db) {
    __shared__ float red[8][33];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, o = blockIdx.x * 32 + lane;
    float s = 0.0f;
    if (o < C)
        for (int b = w; b < B; b += 8) s += __ldg(part + (int64_t)b * C + o);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && o < C) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) t += red[k][lane];
        db[o] = t;
    }
}

constexpr int kWgradSplitK = 128;  // joint indices per split

// ---- TMA-fed route (pp_tc2.cuh): per-sample contractions, every operand streamed by TMA in place and split by the kernel --------
//     y [b][o][p] = Σ_c W[o][c] x[b][c][p] + bias[o]     A = W (K-major, shared by every sample),   B = x[b]  read MN-major
//     dx[b][c][p] = Σ_o W[o][c] dy[b][o][p]               A = W read MN-major (rows c, K lines o),   B = dy[b] read MN-major
//     dW[o][c]    = Σ_b Σ_p dy[b][o][p] x[b][c][p]        per sample: A = dy[b], B = x[b], both K-major; deterministic sum over b
// No transposes, no plane passes, no workspace beyond the per-sample partials of dW.
static inline bool conv_tc2(int Cin, int Cout, int P) {
    static const int off = [] { const char* e = getenv("PIXPRO_B200_TC2"); return (e && e[0] == '0') ? 1 : 0; }();
    return !off && use_tensor_cores(P) && Cin % 4 == 0 && Cout % 4 == 0 && P % 4 == 0;
}
// Grids the in-place route cannot stream (P % 4 != 0: a [C, 49] map of the 7x7 grid has 196-byte rows, TMA strides are multiples
// of 16 bytes) stay on the thread-staged kernel.  A route through the TMA-fed kernel on copies padded to a pitch of 52 floats
// (per-sample y and dx, dW summed over 4 samples per tile: pp_tc_gemm_ex's pitch / kb features) was built and measured at
// B = 128, 256 -> 256, 7x7: dgrad 36.9 -> 29.2 us, wgrad 33.5 -> 25.0 us, forward unchanged, but with the pad launches the step got
// SLOWER (0.689 -> 0.710 ms dense, 0.218 -> 0.237 ms sparse; profiles/r02_v_conv7.txt, r02_w_convpad_step.txt) — these contractions
// are bound by the latency chain of one tile, not by operand staging — and was removed.
struct TcStBias {  // out[b][m][n..n+3] = v + bias[m]
    static constexpr bool kAux = false;
    float* out;
    const float* bias;  // may be null
    int M, N;
    __device__ __forceinline__ void store4(int64_t b, int m, int n, float4 v) const {
        const float bv = bias ? __ldg(bias + m) : 0.0f;
        st4_guard(out + (b * M + m) * (int64_t)N + n, n, N, make_float4(v.x + bv, v.y + bv, v.z + bv, v.w + bv));
    }
    __device__ __forceinline__ void store16(int64_t b, int m, int n, const float v[16]) const {
        const float bv = bias ? __ldg(bias + m) : 0.0f;
        float* q = out + (b * M + m) * (int64_t)N + n;
#pragma unroll
        for (int u = 0; u < 16; u++)
            if (n + u < N) q[u] = v[u] + bv;
    }
};

}  // namespace pp

using namespace pp;

extern "C" {

int64_t pp_conv1x1_fwd_workspace(int64_t B, int Cin, int Cout, int P) {
    (void)B; (void)Cin; (void)Cout; (void)P;
    return 0;  // the TMA route streams x and W in place (kept in the ABI: earlier builds staged planes here)
}

int pp_conv1x1_fwd(const float* x, const float* w, const float* bias, int64_t B, int Cin, int Cout, int P, float* y,
                   void* workspace, void* stream) {
    PP_REQUIRE(x && w && y, "pp_conv1x1_fwd: null pointer");
    PP_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && P > 0 && B * P < (1 << 24), "pp_conv1x1_fwd: bad shape");
    const int NP = (int)(B * P);
    (void)workspace;
    if (conv_tc2(Cin, Cout, P)) {
        tc2::Operands o{w, nullptr, x, nullptr, Cin};
        o.b_mn = true;
        const int rc = tc2::launch_tc2_sets("conv1x1 fwd (tcgen05)", B, Cout, P, &o, 1, TcStBias{y, bias, Cout, P}, (cudaStream_t)stream, true);
        if (rc >= 0) return rc;
    }
    return launch_tc("conv1x1 fwd (tcgen05)", 1, Cout, NP, Cin, LdW{w, Cout, Cin}, LdNC{x, Cin, NP, make_divp(P)},
                     StCN{y, bias, Cout, NP, make_divp(P)}, (cudaStream_t)stream);
}

int64_t pp_conv1x1_bwd_workspace(int64_t B, int Cin, int Cout, int P) {
    const int64_t splits = (B * P + kWgradSplitK - 1) / kWgradSplitK;
    int64_t f = splits * Cin * Cout + B * Cout;  // split-K partials of dW, then the row sums of db
    if (conv_tc2(Cin, Cout, P)) f += B * (int64_t)Cout * Cin;  // wgrad: per-sample partials of dW
    return f * (int64_t)sizeof(float);
}

int pp_conv1x1_bwd(const float* x, const float* w, const float* dy, int64_t B, int Cin, int Cout, int P, float* dx,
                   float* dw, float* db, void* workspace, void* stream) {
    PP_REQUIRE(x && w && dy && workspace, "pp_conv1x1_bwd: null pointer");
    PP_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && P > 0 && B * P < (1 << 24), "pp_conv1x1_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int NP = (int)(B * P);
    int rc;
    const int64_t base_f = (int64_t)((NP + kWgradSplitK - 1) / kWgradSplitK) * Cin * Cout + B * Cout;
    const bool tma = conv_tc2(Cin, Cout, P);
    float* part_b = (float*)workspace + base_f;  // per-sample partials of dW (TMA route)
    if (dx && tma) {
        tc2::Operands o{w, nullptr, dy, nullptr, Cout};
        o.a_mn = o.b_mn = true;
        rc = tc2::launch_tc2_sets("conv1x1 dgrad (tcgen05)", B, Cin, P, &o, 1, TcStBias{dx, nullptr, Cin, P}, st, true);
        if (rc > 0) return rc;
        if (rc == 0) dx = nullptr;  // done
    }
    if (dw && tma) {
        // both operands as they are ([B,C,P]: K = the pixel index is contiguous), split by the kernel's converter warps
        rc = tc2::launch_tc2("conv1x1 wgrad (tcgen05)", B, Cout, Cin, P, dy, nullptr, x, nullptr, TcStN{part_b, Cout, Cin}, st);
        if (rc > 0) return rc;
        if (rc == 0) {
            const int total = Cout * Cin;
            PP_LAUNCH("conv1x1 wgrad reduce", st, conv_wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(part_b, (int)B, total, dw));
            rc = check_launch("conv_wgrad_reduce_kernel");
            if (rc) return rc;
            dw = nullptr;  // done
        }
    }
    if (dx) {
        rc = launch_tc("conv1x1 dgrad (tcgen05)", 1, Cin, NP, Cout, LdWT{w, Cout, Cin}, LdNC{dy, Cout, NP, make_divp(P)},
                       StCN{dx, nullptr, Cin, NP, make_divp(P)}, st);
        if (rc) return rc;
    }
    if (dw) {
        const int splits = (NP + kWgradSplitK - 1) / kWgradSplitK;
        PP_REQUIRE(splits <= 65535, "pp_conv1x1_bwd: too many split-K slices");
        float* part = (float*)workspace;
        rc = launch_tc("conv1x1 wgrad (tcgen05)", splits, Cout, Cin, kWgradSplitK, LdCN{dy, Cout, NP, kWgradSplitK, make_divp(P)},
                       LdCN{x, Cin, NP, kWgradSplitK, make_divp(P)}, TcStN{part, Cout, Cin}, st);
        if (rc) return rc;
        const int total = Cout * Cin;
        PP_LAUNCH("conv1x1 wgrad reduce", st, conv_wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(part, splits, total, dw));
        rc = check_launch("conv_wgrad_reduce_kernel");
        if (rc) return rc;
    }
    if (db) {
        // its own workspace region: the three gradients may run concurrently on different streams
        float* rows = (float*)workspace + (int64_t)((NP + kWgradSplitK - 1) / kWgradSplitK) * Cin * Cout;
        const int nrows = (int)(B * Cout);
        PP_LAUNCH("conv1x1 bias grad", st, conv_bias_rows_kernel<<<(nrows + 7) / 8, 256, 0, st>>>(dy, nrows, P, rows));
        PP_LAUNCH("conv1x1 bias grad reduce", st, conv_bias_grad_kernel<<<(Cout + 31) / 32, 256, 0, st>>>(rows, (int)B, Cout, db));
        rc = check_launch("conv_bias_grad_kernel");
        if (rc) return rc;
    }
    return PP_OK;
}

}  // extern "C"
