// pp_ppm.cu — Pixel Propagation Module forward/backward (sm_100a), generic-size fp32 path.
//
// Reference restated: PixPro.featprop, contrast/models/PixPro.py:339-363, and the caller's
// F.normalize (:380,385).  For one sample with x = feat [C,P], v = value_transform(feat) [C,P]:
//     x̂ = x / max(‖x‖_c, 1e-12),  v̂ likewise
//     S = x̂ᵀ x̂ [P,P];  A = (clamp(S, min=cv) [+1e-6 if γ<1])^γ
//     Y = v̂ Aᵀ [C,P];  out = Y / max(‖Y‖_c, 1e-12)        (if final_norm)
// Backward (validated against torch autograd through the oracle):
//     gy = (g − ŷ (g·ŷ)) / ‖Y‖ ;  gA = gyᵀ v̂ ;  gv̂ = gy A ;  gS = gA ∘ A'(S) [S ≥ cv]
//     gx̂ = x̂ (gS + gSᵀ) ;  d_feat_sim = normbwd(x, gx̂) ;  d_val = normbwd(v, gv̂)
//
// This file is the size-generic fp32 CUDA-core path: a batched 64x64x16 register-tiled GEMM
// whose operand loads and epilogue are functors, so the normalisations, relu^γ, its
// derivative and the (gS+gSᵀ) symmetrisation are all fused into the contractions and no
// normalised copy of x or v is ever written.  Column norms use one warp-shuffle reduction
// kernel.  saved = [nx | nv | ny | S] per batch (3·B·P + B·P·P floats).
#include <math.h>
#include <stdlib.h>

#include "pp_common.cuh"
#include "pp_ppm.cuh"
#include "pp_tc.cuh"
#include "pp_tc2.cuh"

namespace pp {

// ---- column norms over C: n[b,i] = max(sqrt(Σ_c u[b,c,i]^2), eps) --------------------------
// grid (ceil(P/32), B), block (32, 8): threadIdx.x = column, threadIdx.y strides over C.
__global__ void __launch_bounds__(256) colnorm_kernel(const float* __restrict__ u, int C, int P, float* __restrict__ nrm) {
    __shared__ float red[8][33];
    int i = blockIdx.x * 32 + threadIdx.x;
    int64_t b = blockIdx.y;
    float s = 0.0f;
    if (i < P)
        for (int c = threadIdx.y; c < C; c += 8) {
            float v = __ldg(u + (b * C + c) * (int64_t)P + i);
            s = fmaf(v, v, s);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && i < P) {
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++) t += red[r][threadIdx.x];
        nrm[b * P + i] = fmaxf(sqrtf(t), kNormEps);
    }
}

// out[b,c,i] = y[b,c,i] / n[b,i]
__global__ void __launch_bounds__(256) coldiv_kernel(const float* y, const float* __restrict__ nrm, int C, int P, int64_t total,
                                                      float* out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int i = (int)(e % P);
    int64_t b = e / ((int64_t)C * P);
    out[e] = y[e] / nrm[b * P + i];
}

// out[b,c,i] = y[b,c,i] / n[b,i] (n == nullptr: y itself) and its hi / lo split in the same [B,C,P] layout: the planes the
// similarity contractions stream MN-major (splitting 24 KB of operands per K chunk inside the contraction kernel costs it a
// quarter of its tensor-pipe rate — shared-memory bandwidth — so operands that feed a P x P output are split here, once)
__global__ void __launch_bounds__(256) coldiv_split_kernel(const float* __restrict__ y, const float* __restrict__ nrm, int C, int P, int64_t total,
                                                            float* __restrict__ out, float* __restrict__ hi, float* __restrict__ lo) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    float v = __ldg(y + e);
    if (nrm) {
        int i = (int)(e % P);
        int64_t b = e / ((int64_t)C * P);
        v = v / nrm[b * P + i];
    }
    if (out) out[e] = v;
    float h, l;
    tc2::split1(v, h, l);
    hi[e] = h;
    lo[e] = l;
}

// normalisation backward: out = (g − û (g·û)) / n, û = u / n_u  (û given by u and its norm,
// or directly when u is already normalised: pass nu = nullptr).
// grid (ceil(P/32), B), block (32,8).
__global__ void __launch_bounds__(256) normbwd_kernel(const float* __restrict__ g, const float* __restrict__ u,
                                                       const float* __restrict__ nu, const float* __restrict__ ndiv, int C,
                                                       int P, float* __restrict__ out) {
    __shared__ float red[8][33];
    __shared__ float dot[32];
    int i = blockIdx.x * 32 + threadIdx.x;
    int64_t b = blockIdx.y;
    float rn = 1.0f, rdiv = 1.0f;
    if (i < P) {
        rn = nu ? 1.0f / nu[b * P + i] : 1.0f;
        rdiv = 1.0f / ndiv[b * P + i];
    }
    float s = 0.0f;
    if (i < P)
        for (int c = threadIdx.y; c < C; c += 8) {
            int64_t o = (b * C + c) * (int64_t)P + i;
            s = fmaf(__ldg(g + o), __ldg(u + o) * rn, s);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++) t += red[r][threadIdx.x];
        dot[threadIdx.x] = t;
    }
    __syncthreads();
    if (i < P) {
        float d = dot[threadIdx.x];
        for (int c = threadIdx.y; c < C; c += 8) {
            int64_t o = (b * C + c) * (int64_t)P + i;
            out[o] = (__ldg(g + o) - __ldg(u + o) * rn * d) * rdiv;
        }
    }
}

// ---- the same passes with the sample's 32-column tile held in registers: every tensor is read ONCE ------------------
// colnorm_kernel + coldiv(_split)_kernel read their input twice (norm pass, then scale pass: 2 launches), normbwd_kernel
// reads g and u twice inside one launch, and normbwd + coldiv_split write gy only to read it back.  At the 28x28 grid these
// passes move 1.9 GB per step against 1.0 GB for one read per tensor and cost 0.40 ms of a 2.1 ms step.  Here a block owns
// 32 columns x all C <= 8 RMAX channels of one sample: thread (x, y) keeps rows y, y + 8, ... of column x in registers, the
// column sums take the same order as in the two-pass kernels (bit-identical norms and outputs), and the scaled values leave as
// the fp32 tensor and / or its hi / lo planes.  grid (ceil(P/32), B), block (32, 8).
constexpr int kRegRows = 32;  // C <= 256
struct NormJob {  // one tensor of a colnorm_scale launch
    const float* u;
    float *nrm, *out, *hi, *lo;
};
struct NormJobs {
    NormJob j[2];  // blockIdx.z selects: feat and val of a forward share one launch (fuller waves, one launch latency)
};
template <bool SPLIT>
__global__ void __launch_bounds__(256) colnorm_scale_kernel(NormJobs jobs, int C, int P) {
    __shared__ float red[8][33];
    __shared__ float nsh[32];
    const NormJob& jb = jobs.j[blockIdx.z];
    const float* u = jb.u;
    float *nrm = jb.nrm, *out = jb.out, *hi = jb.hi, *lo = jb.lo;
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t base = b * C * (int64_t)P + i;
    float v[kRegRows];
#pragma unroll
    for (int r = 0; r < kRegRows; r++) {
        const int c = threadIdx.y + 8 * r;
        v[r] = (i < P && c < C) ? u[base + c * (int64_t)P] : 0.0f;  // plain loads: `out` may alias `u` (in-place final norm)
    }
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < kRegRows; r++)
        if (threadIdx.y + 8 * r < C) s = fmaf(v[r], v[r], s);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++) t += red[r][threadIdx.x];
        t = fmaxf(sqrtf(t), kNormEps);
        nsh[threadIdx.x] = t;
        if (i < P) nrm[b * P + i] = t;
    }
    __syncthreads();
    const float n = nsh[threadIdx.x];
    if (i < P) {
#pragma unroll
        for (int r = 0; r < kRegRows; r++) {
            const int c = threadIdx.y + 8 * r;
            if (c < C) {
                const int64_t o = base + c * (int64_t)P;
                const float q = v[r] / n;
                if (out) out[o] = q;
                if (SPLIT) {
                    float h, l;
                    tc2::split1(q, h, l);
                    hi[o] = h;
                    lo[o] = l;
                }
            }
        }
    }
}
// normbwd_kernel with g and u in registers; out (fp32, may be null when SPLIT) and / or the hi / lo planes of the result
struct NormBwdJob {
    const float *g, *u, *nu, *ndiv;
    float *out, *hi, *lo;
};
struct NormBwdJobs {
    NormBwdJob j[2];  // blockIdx.z selects: d_feat and d_val of a backward share one launch
};
template <bool SPLIT>
__global__ void __launch_bounds__(256) normbwd_reg_kernel(NormBwdJobs jobs, int C, int P) {
    __shared__ float red[8][33];
    __shared__ float dot[32];
    const NormBwdJob& jb = jobs.j[blockIdx.z];
    const float* __restrict__ g = jb.g;
    const float* __restrict__ u = jb.u;
    const float* __restrict__ nu = jb.nu;
    const float* __restrict__ ndiv = jb.ndiv;
    float* __restrict__ out = jb.out;
    float* __restrict__ hi = jb.hi;
    float* __restrict__ lo = jb.lo;
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t base = b * C * (int64_t)P + i;
    float rn = 1.0f, rdiv = 1.0f;
    if (i < P) {
        rn = nu ? 1.0f / nu[b * P + i] : 1.0f;
        rdiv = 1.0f / ndiv[b * P + i];
    }
    float gv[kRegRows], uv[kRegRows];
#pragma unroll
    for (int r = 0; r < kRegRows; r++) {
        const int c = threadIdx.y + 8 * r;
        const bool in = i < P && c < C;
        gv[r] = in ? __ldg(g + base + c * (int64_t)P) : 0.0f;
        uv[r] = in ? __ldg(u + base + c * (int64_t)P) : 0.0f;
    }
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < kRegRows; r++)
        if (threadIdx.y + 8 * r < C) s = fmaf(gv[r], uv[r] * rn, s);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++) t += red[r][threadIdx.x];
        dot[threadIdx.x] = t;
    }
    __syncthreads();
    if (i < P) {
        const float d = dot[threadIdx.x];
#pragma unroll
        for (int r = 0; r < kRegRows; r++) {
            const int c = threadIdx.y + 8 * r;
            if (c < C) {
                const int64_t o = base + c * (int64_t)P;
                const float q = (gv[r] - uv[r] * rn * d) * rdiv;
                if (out) out[o] = q;
                if (SPLIT) {
                    float h, l;
                    tc2::split1(q, h, l);
                    hi[o] = h;
                    lo[o] = l;
                }
            }
        }
    }
}
// PIXPRO_B200_PPMREG=0: the two-pass kernels (A/B switch)
static inline bool ppm_reg_passes(int C) {
    static const int off = [] { const char* e = getenv("PIXPRO_B200_PPMREG"); return (e && e[0] == '0') ? 1 : 0; }();
    return !off && C <= 8 * kRegRows;
}

static inline NormJobs norm_jobs(NormJob a, NormJob b = NormJob{}) {
    NormJobs j;
    j.j[0] = a;
    j.j[1] = b;
    return j;
}
static inline NormBwdJobs normbwd_jobs(NormBwdJob a, NormBwdJob b = NormBwdJob{}) {
    NormBwdJobs j;
    j.j[0] = a;
    j.j[1] = b;
    return j;
}

// ---- batched register-tiled GEMM with functor operands ------------------------------------
// Cmn = Σ_k A(m,k) B(k,n);  LA/LB: element loaders (b, m|n, k) -> float;  EP: epilogue.
constexpr int BM = 64, BN = 64, BK = 16;

template <class LA, class LB, class EP>
__global__ void __launch_bounds__(256) bgemm_kernel(int M, int N, int K, LA la, LB lb, EP ep) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int64_t b = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // cooperative loads: 16x64 elements each, 4 per thread
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int e = threadIdx.x + r * 256;
            if (LA::kMajorM) {
                int mm = e & 63, kk = e >> 6;
                As[kk][mm] = (m0 + mm < M && k0 + kk < K) ? la(b, m0 + mm, k0 + kk) : 0.0f;
            } else {
                int kk = e & 15, mm = e >> 4;
                As[kk][mm] = (m0 + mm < M && k0 + kk < K) ? la(b, m0 + mm, k0 + kk) : 0.0f;
            }
            if (LB::kMajorN) {
                int nn = e & 63, kk = e >> 6;
                Bs[kk][nn] = (n0 + nn < N && k0 + kk < K) ? lb(b, n0 + nn, k0 + kk) : 0.0f;
            } else {
                int kk = e & 15, nn = e >> 4;
                Bs[kk][nn] = (n0 + nn < N && k0 + kk < K) ? lb(b, n0 + nn, k0 + kk) : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) ep(b, m, n, acc[i][j]);
        }
}

// Operand loaders.  kMajorM/kMajorN = the m (or n) index is the contiguous one in memory, so
// the cooperative load walks it fastest (coalesced).
// x[b, k=c, m=i] * rn[b,i]      ([C,P] tensor read "transposed": m contiguous)
struct LoadColScaled {  // element (m=i, k=c) of x̂ᵀ
    static constexpr bool kMajorM = true, kMajorN = true;
    const float *x, *rn;  // rn may be null (no scaling)
    int C, P;
    __device__ __forceinline__ float operator()(int64_t b, int i, int c) const {
        float v = __ldg(x + (b * C + c) * (int64_t)P + i);
        return rn ? v / __ldg(rn + b * P + i) : v;
    }
};
// x[b, m=c, k=j] / n[b,j]        ([C,P] tensor read row-wise: k contiguous)
struct LoadRowScaled {
    static constexpr bool kMajorM = false, kMajorN = false;
    const float *x, *rn;
    int C, P;
    __device__ __forceinline__ float operator()(int64_t b, int c, int j) const {
        float v = __ldg(x + (b * C + c) * (int64_t)P + j);
        return rn ? v / __ldg(rn + b * P + j) : v;
    }
};
// B(k=j, n=i) = f(S[b,i,j])      (S row i contiguous in j = k)
struct LoadActS {
    static constexpr bool kMajorM = false, kMajorN = false;
    const float* S;
    int P;
    Act act;
    __device__ __forceinline__ float operator()(int64_t b, int i, int j) const {
        return act.f(__ldg(S + (b * P + i) * (int64_t)P + j));
    }
};
// B(k=i, n=j) = f(S[b,i,j])      (n contiguous)
struct LoadActS_T {
    static constexpr bool kMajorM = true, kMajorN = true;
    const float* S;
    int P;
    Act act;
    __device__ __forceinline__ float operator()(int64_t b, int j, int i) const {
        return act.f(__ldg(S + (b * P + i) * (int64_t)P + j));
    }
};
// B(k=j, n=i) = gS[b,i,j] + gS[b,j,i]
struct LoadSym {
    static constexpr bool kMajorM = false, kMajorN = false;
    const float* gS;
    int P;
    __device__ __forceinline__ float operator()(int64_t b, int i, int j) const {
        return __ldg(gS + (b * P + i) * (int64_t)P + j) + __ldg(gS + (b * P + j) * (int64_t)P + i);
    }
};

struct EpStore {  // out[b,m,n] = v
    float* out;
    int M, N;
    __device__ __forceinline__ void operator()(int64_t b, int m, int n, float v) const { out[(b * M + m) * (int64_t)N + n] = v; }
};
struct EpGradS {  // gS[b,i,j] = v * A'(S[b,i,j])
    float* gS;
    const float* S;
    int P;
    Act act;
    __device__ __forceinline__ void operator()(int64_t b, int i, int j, float v) const {
        int64_t o = (b * P + i) * (int64_t)P + j;
        gS[o] = v * act.df(__ldg(S + o));
    }
};

// ---- tensor-core route for large grids (tcgen05 3xTF32, pp_tc.cuh) --------------------------------
// Specialised operand loaders for the tensor-core route (operands are pre-normalised once by
// coldiv_kernel, so loads are plain vector / coalesced accesses without divisions).
struct TcLdActN {  // relu^γ(S[row][k..k+3])  (S symmetric: also serves A[k][row])
    static constexpr bool kRowMajorK = true;
    const float* S;
    int P;
    Act act;
    __device__ __forceinline__ float4 load4(int64_t b, int row, int k) const {
        if (row >= P || k >= P) return make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = ldg4_guard(S + (b * P + row) * (int64_t)P + k, k, P);
        v.x = act.f(v.x);
        v.y = k + 1 < P ? act.f(v.y) : 0.f;
        v.z = k + 2 < P ? act.f(v.z) : 0.f;
        v.w = k + 3 < P ? act.f(v.w) : 0.f;
        return v;
    }
};
struct TcLdSymN {  // gS[row][k..k+3] + gS[k..k+3][row]
    static constexpr bool kRowMajorK = true;
    const float* gS;
    int P;
    __device__ __forceinline__ float4 load4(int64_t b, int row, int k) const {
        if (row >= P || k >= P) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = gS + b * (int64_t)P * P;
        float4 v = ldg4_guard(base + row * (int64_t)P + k, k, P);
        v.x += __ldg(base + k * (int64_t)P + row);
        if (k + 1 < P) v.y += __ldg(base + (k + 1) * (int64_t)P + row);
        if (k + 2 < P) v.z += __ldg(base + (k + 2) * (int64_t)P + row);
        if (k + 3 < P) v.w += __ldg(base + (k + 3) * (int64_t)P + row);
        return v;
    }
};
struct TcStGradS {  // gS[b][i][j] = v * A'(S[b][i][j])
    float* gS;
    const float* S;
    int P;
    Act act;
    __device__ __forceinline__ void store16(int64_t b, int i, int j, const float v[16]) const {
        const int64_t o = (b * P + i) * (int64_t)P + j;
#pragma unroll
        for (int u = 0; u < 16; u++)
            if (j + u < P) gS[o + u] = v[u] * act.df(__ldg(S + o + u));
    }
};

// ---- TMA-fed tensor-core route (pp_tc2.cuh) ---------------------------------------------------------------------------------
// [B,C,P] operands (x̂, v̂, gy) are split into hi / lo planes once, in their own layout, by the pass that normalises them
// (coldiv_split_kernel), and streamed by TMA in place — K-major where the contraction runs over pixels, MN-major where it runs
// over channels: no transposed copies exist.  The P x P operands (relu^gamma(S), gS + gS^T) are written as planes by the
// epilogue of the contraction that produces them.
// Epilogue of the similarity contraction: S itself (kept for the backward) and relu^gamma(S) already split into the planes the
// propagation contractions stream — the activation pass over S is gone.  Requires P % 4 == 0 (the TMA route does).
struct TcStSAct {
    static constexpr bool kAux = false;
    float *S, *hi, *lo;
    int P;
    Act act;
    __device__ __forceinline__ void store4(int64_t b, int i, int j, float4 v) const {  // P % 4 == 0: j + 3 < P, 16-byte aligned
        const int64_t o = (b * P + i) * (int64_t)P + j;
        float4 h, l;
        tc2::split1(act.f(v.x), h.x, l.x);
        tc2::split1(act.f(v.y), h.y, l.y);
        tc2::split1(act.f(v.z), h.z, l.z);
        tc2::split1(act.f(v.w), h.w, l.w);
        *reinterpret_cast<float4*>(S + o) = v;
        *reinterpret_cast<float4*>(hi + o) = h;
        *reinterpret_cast<float4*>(lo + o) = l;
    }
    __device__ __forceinline__ void store16(int64_t b, int i, int j, const float v[16]) const {
        const int64_t o = (b * P + i) * (int64_t)P + j;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (j + 4 * q < P) {
                float4 h, l;
                tc2::split1(act.f(v[4 * q]), h.x, l.x);
                tc2::split1(act.f(v[4 * q + 1]), h.y, l.y);
                tc2::split1(act.f(v[4 * q + 2]), h.z, l.z);
                tc2::split1(act.f(v[4 * q + 3]), h.w, l.w);
                *reinterpret_cast<float4*>(S + o + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                *reinterpret_cast<float4*>(hi + o + 4 * q) = h;
                *reinterpret_cast<float4*>(lo + o + 4 * q) = l;
            }
        }
    }
};
// Epilogue of the backward's similarity-gradient contraction.  The accumulator holds G + G^T (two products, see below); with S
// (hence A'(S)) symmetric, (G + G^T) o A'(S) IS gS + gS^T — written straight as the hi / lo planes the gx contraction streams.
struct TcStSymPlanes {
    float *hi, *lo;
    const float* S;
    int P;
    Act act;
    static constexpr bool kAux = true;  // S[b][i][j..j+3], loaded ahead of the accumulator read
    __device__ __forceinline__ float4 aux4(int64_t b, int i, int j, int M, int N) const {
        if (i >= P || j >= P) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(S + (b * P + i) * (int64_t)P + j));
    }
    __device__ __forceinline__ void prefetch_line(int64_t b, int i, int j) const {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(S + (b * P + i) * (int64_t)P + j));
    }
    __device__ __forceinline__ void store4(int64_t b, int i, int j, float4 v, float4 s) const {
        const int64_t o = (b * P + i) * (int64_t)P + j;
        float4 h, l;
        tc2::split1(v.x * act.df(s.x), h.x, l.x);
        tc2::split1(v.y * act.df(s.y), h.y, l.y);
        tc2::split1(v.z * act.df(s.z), h.z, l.z);
        tc2::split1(v.w * act.df(s.w), h.w, l.w);
        *reinterpret_cast<float4*>(hi + o) = h;
        *reinterpret_cast<float4*>(lo + o) = l;
    }
};
// The TMA route needs 16-byte row strides in every plane: C % 4 == 0 and P % 4 == 0 (PIXPRO_B200_TC2=0 disables it).
// Pre-split planes of the [B,C,P] operands pay off at large grids (28x28: the similarity contractions are 15-25 % faster than with
// the split done by the contraction kernel's converter warps, which costs it shared-memory bandwidth); at 14x14 the extra pass
// over the operands costs more than it saves (measured 1.16 vs 1.11 ms/step), so there the fp32 tensors are streamed as they are.
static inline bool ppm_presplit(int P) { return P >= 512; }
static inline bool ppm_tc2(int C, int P) {
    static const int off = [] { const char* e = getenv("PIXPRO_B200_TC2"); return (e && e[0] == '0') ? 1 : 0; }();
    return !off && use_tensor_cores(P) && C % 4 == 0 && P % 4 == 0;
}
template <class LA, class LB, class EP>
static int launch_bgemm(const char* what, int64_t B, int M, int N, int K, LA la, LB lb, EP ep, cudaStream_t st) {
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, (unsigned)B);
    PP_LAUNCH(what, st, (bgemm_kernel<LA, LB, EP><<<grid, 256, 0, st>>>(M, N, K, la, lb, ep)));
    return check_launch(what);
}

struct Saved {
    float *nx, *nv, *ny, *S, *xh, *vh;  // xh, vh: normalised operands (generic path only)
    // TMA route only (ppm_tc2): hi / lo planes of relu^gamma(S) [B,P,P], and of x̂ and v̂ in their own [B,C,P] layout
    float *a_hi, *a_lo, *x_hi, *x_lo, *v_hi, *v_lo;
};
static Saved carve_saved(void* p, int64_t B, int C, int P) {
    Saved s;
    float* f = (float*)p;
    s.nx = f; f += B * P;
    s.nv = f; f += B * P;
    s.ny = f; f += B * P;
    s.S = f; f += B * (int64_t)P * P;
    s.xh = f; f += B * (int64_t)C * P;
    s.vh = f; f += B * (int64_t)C * P;
    const int64_t pp2 = B * (int64_t)P * P, cp = B * (int64_t)C * P;
    s.a_hi = f; f += pp2;
    s.a_lo = f; f += pp2;
    s.x_hi = f; f += cp;
    s.x_lo = f; f += cp;
    s.v_hi = f; f += cp;
    s.v_lo = f;
    return s;
}

}  // namespace pp

using namespace pp;

extern "C" {

int64_t pp_ppm_saved_bytes(int64_t B, int C, int P) {
    int64_t f = 3 * B * P + B * (int64_t)P * P;
    if (!ppm_small_supported(C, P)) {
        f += 2 * B * (int64_t)C * P;  // normalised operands kept for the backward contractions
        if (ppm_tc2(C, P)) f += 2 * B * (int64_t)P * P + 4 * B * (int64_t)C * P;  // hi / lo planes of relu^gamma(S), x̂, v̂
    }
    return f * (int64_t)sizeof(float);
}

int64_t pp_ppm_bwd_workspace(int64_t B, int C, int P) {
    // gy [B,C,P] + gS [B,P,P] + gvh [B,C,P] + gxh [B,C,P]
    int64_t f = 3 * B * (int64_t)C * P + B * (int64_t)P * P;
    if (!ppm_small_supported(C, P) && ppm_tc2(C, P)) f += 2 * B * (int64_t)P * P + 2 * B * (int64_t)C * P;  // (gS+gS^T) planes, gy planes
    return f * (int64_t)sizeof(float);
}

int pp_ppm_fwd(const float* feat, const float* val, int64_t B, int C, int P, double gamma, double clamp_value, int final_norm,
               float* out, void* saved, void* stream) {
    PP_REQUIRE(feat && val && out && saved, "pp_ppm_fwd: null pointer");
    PP_REQUIRE(B > 0 && B <= 65535 && C > 0 && P > 0, "pp_ppm_fwd: bad shape B=%lld C=%d P=%d", (long long)B, C, P);
    cudaStream_t st = (cudaStream_t)stream;
    Saved sv = carve_saved(saved, B, C, P);
    Act act = make_act(gamma, clamp_value);
    if (ppm_small_supported(C, P))  // e.g. the 7x7 grid: one block per sample, one launch
        return ppm_fwd_small(feat, val, B, C, P, act, final_norm, out, sv.nx, sv.nv, sv.ny, sv.S, st);
    dim3 nb((P + 31) / 32, (unsigned)B), nt(32, 8);
    const bool reg = ppm_reg_passes(C) && (ppm_tc2(C, P) || use_tensor_cores(P));  // norm + scale (+ split) in one pass per tensor
    int rc = PP_OK;
    if (!reg) {
        PP_LAUNCH("ppm colnorm", st, colnorm_kernel<<<nb, nt, 0, st>>>(feat, C, P, sv.nx));
        PP_LAUNCH("ppm colnorm", st, colnorm_kernel<<<nb, nt, 0, st>>>(val, C, P, sv.nv));
        rc = check_launch("colnorm_kernel");
        if (rc) return rc;
    }
    // S[i][j] = Σ_c x̂[c][i] x̂[c][j]
    const int64_t total = B * (int64_t)C * P;
    if (ppm_tc2(C, P)) {
        // TMA route: normalise once (x̂, v̂ kept for the backward), then stream them in place
        const bool pre = ppm_presplit(P);
        const dim3 nb2(nb.x, nb.y, 2);
        if (reg && pre) {
            PP_LAUNCH("ppm colnorm+div", st, colnorm_scale_kernel<true><<<nb2, nt, 0, st>>>(
                norm_jobs(NormJob{feat, sv.nx, sv.xh, sv.x_hi, sv.x_lo}, NormJob{val, sv.nv, sv.vh, sv.v_hi, sv.v_lo}), C, P));
        } else if (reg) {
            PP_LAUNCH("ppm colnorm+div", st, colnorm_scale_kernel<false><<<nb2, nt, 0, st>>>(
                norm_jobs(NormJob{feat, sv.nx, sv.xh, nullptr, nullptr}, NormJob{val, sv.nv, sv.vh, nullptr, nullptr}), C, P));
        } else if (pre) {
            PP_LAUNCH("ppm coldiv", st, coldiv_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, sv.nx, C, P, total, sv.xh, sv.x_hi, sv.x_lo));
            PP_LAUNCH("ppm coldiv", st, coldiv_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(val, sv.nv, C, P, total, sv.vh, sv.v_hi, sv.v_lo));
        } else {
            PP_LAUNCH("ppm coldiv", st, coldiv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, sv.nx, C, P, total, sv.xh));
            PP_LAUNCH("ppm coldiv", st, coldiv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(val, sv.nv, C, P, total, sv.vh));
        }
        rc = check_launch("ppm coldiv");
        if (rc) return rc;
        const float *x_hi = pre ? sv.x_hi : sv.xh, *x_lo = pre ? sv.x_lo : nullptr, *v_hi = pre ? sv.v_hi : sv.vh, *v_lo = pre ? sv.v_lo : nullptr;
        {   // S[i][j] = Σ_c x̂[c][i] x̂[c][j]: both operands are x̂ read MN-major ([C][P]: the channel index is the K line)
            tc2::Operands o{x_hi, x_lo, x_hi, x_lo, C};
            o.a_mn = o.b_mn = true;
            rc = tc2::launch_tc2_sets("ppm S (tcgen05)", B, P, P, &o, 1, TcStSAct{sv.S, sv.a_hi, sv.a_lo, P, act}, st, false);
        }
        if (rc > 0) return rc;
        if (rc == 0) rc = tc2::launch_tc2("ppm Y (tcgen05)", B, C, P, P, v_hi, v_lo, sv.a_hi, sv.a_lo, TcStN{out, C, P}, st);
        if (rc > 0) return rc;
        if (rc < 0) {  // tensor maps could not be encoded: the thread-staged kernels read the same normalised tensors
            rc = launch_tc("ppm S (tcgen05)", B, P, P, C, TcLdT{sv.xh, C, P}, TcLdT{sv.xh, C, P}, TcStSAct{sv.S, sv.a_hi, sv.a_lo, P, act}, st);
            if (rc) return rc;
            rc = launch_tc("ppm Y (tcgen05)", B, C, P, P, TcLdN{sv.vh, C, P}, TcLdActN{sv.S, P, act}, TcStN{out, C, P}, st);
            if (rc) return rc;
        }
    } else if (use_tensor_cores(P)) {
        // normalise once (x̂, v̂ kept for backward), then the two contractions on the tensor cores
        if (reg) {
            PP_LAUNCH("ppm colnorm+div", st, colnorm_scale_kernel<false><<<dim3(nb.x, nb.y, 2), nt, 0, st>>>(
                norm_jobs(NormJob{feat, sv.nx, sv.xh, nullptr, nullptr}, NormJob{val, sv.nv, sv.vh, nullptr, nullptr}), C, P));
        } else {
            PP_LAUNCH("ppm coldiv", st, coldiv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(feat, sv.nx, C, P, total, sv.xh));
            PP_LAUNCH("ppm coldiv", st, coldiv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(val, sv.nv, C, P, total, sv.vh));
        }
        rc = check_launch("ppm coldiv");
        if (rc) return rc;
        rc = launch_tc("ppm S (tcgen05)", B, P, P, C, TcLdT{sv.xh, C, P}, TcLdT{sv.xh, C, P}, TcStN{sv.S, P, P}, st);
        if (rc) return rc;
        rc = launch_tc("ppm Y (tcgen05)", B, C, P, P, TcLdN{sv.vh, C, P}, TcLdActN{sv.S, P, act}, TcStN{out, C, P}, st);
        if (rc) return rc;
    } else {
        rc = launch_bgemm("ppm S", B, P, P, C, LoadColScaled{feat, sv.nx, C, P}, LoadColScaled{feat, sv.nx, C, P},
                          EpStore{sv.S, P, P}, st);
        if (rc) return rc;
        rc = launch_bgemm("ppm Y", B, C, P, P, LoadRowScaled{val, sv.nv, C, P}, LoadActS{sv.S, P, act}, EpStore{out, C, P}, st);
        if (rc) return rc;
    }
    if (final_norm) {
        if (ppm_reg_passes(C)) {
            PP_LAUNCH("ppm colnorm+div", st, colnorm_scale_kernel<false><<<nb, nt, 0, st>>>(norm_jobs(NormJob{out, sv.ny, out, nullptr, nullptr}), C, P));
        } else {
            PP_LAUNCH("ppm colnorm", st, colnorm_kernel<<<nb, nt, 0, st>>>(out, C, P, sv.ny));
            PP_LAUNCH("ppm coldiv", st, coldiv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, sv.ny, C, P, total, out));
        }
        rc = check_launch("ppm final normalize");
    }
    return rc;
}

int pp_ppm_bwd(const float* feat, const float* val, const float* out, const float* g, const void* saved, int64_t B, int C,
               int P, double gamma, double clamp_value, int final_norm, float* d_feat_sim, float* d_val, void* workspace,
               void* stream) {
    PP_REQUIRE(feat && val && out && g && saved && d_feat_sim && d_val && workspace, "pp_ppm_bwd: null pointer");
    PP_REQUIRE(B > 0 && B <= 65535 && C > 0 && P > 0, "pp_ppm_bwd: bad shape B=%lld C=%d P=%d", (long long)B, C, P);
    cudaStream_t st = (cudaStream_t)stream;
    Saved sv = carve_saved(const_cast<void*>(saved), B, C, P);
    Act act = make_act(gamma, clamp_value);
    if (ppm_small_supported(C, P))
        return ppm_bwd_small(feat, val, out, g, B, C, P, act, final_norm, sv.nx, sv.nv, sv.ny, sv.S, d_feat_sim, d_val, st);
    float* gy = (float*)workspace;
    float* gS = gy + B * (int64_t)C * P;
    float* gvh = gS + B * (int64_t)P * P;
    float* gxh = gvh + B * (int64_t)C * P;
    dim3 nb((P + 31) / 32, (unsigned)B), nt(32, 8);
    int rc;
    const float* gyp = g;
    const bool regp = ppm_reg_passes(C);
    // pre-split TMA route with the final normalisation: gy leaves its pass directly as hi / lo planes (and as fp32: the
    // fallback below reads it), one launch instead of normbwd + split
    const bool fuse_gy = final_norm && regp && ppm_tc2(C, P) && ppm_presplit(P) && (reinterpret_cast<uintptr_t>(gy) & 15) == 0;
    if (final_norm && !fuse_gy) {  // gy = (g − ŷ (g·ŷ)) / ny ; `out` is ŷ
        if (regp) PP_LAUNCH("ppm normbwd", st, normbwd_reg_kernel<false><<<nb, nt, 0, st>>>(normbwd_jobs(NormBwdJob{g, out, nullptr, sv.ny, gy, nullptr, nullptr}), C, P));
        else PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(g, out, nullptr, sv.ny, C, P, gy));
        rc = check_launch("ppm normbwd(out)");
        if (rc) return rc;
    }
    if (final_norm) gyp = gy;
    if (ppm_tc2(C, P) && (reinterpret_cast<uintptr_t>(gyp) & 15) == 0) {  // gy is streamed by TMA as it is: 16-byte aligned
        float* f = gxh + B * (int64_t)C * P;
        const int64_t pp2 = B * (int64_t)P * P;
        const bool pre = ppm_presplit(P);
        float *sym_hi = f, *sym_lo = sym_hi + pp2;
        const float *gy_hi = gyp, *gy_lo = nullptr, *x_hi = sv.xh, *x_lo = nullptr, *v_hi = sv.vh, *v_lo = nullptr;
        if (pre) {
            float *gh = sym_lo + pp2, *gl = gh + B * (int64_t)C * P;
            gy_hi = gh; gy_lo = gl; x_hi = sv.x_hi; x_lo = sv.x_lo; v_hi = sv.v_hi; v_lo = sv.v_lo;
            const int64_t total = B * (int64_t)C * P;
            if (fuse_gy) PP_LAUNCH("ppm normbwd", st, normbwd_reg_kernel<true><<<nb, nt, 0, st>>>(normbwd_jobs(NormBwdJob{g, out, nullptr, sv.ny, gy, gh, gl}), C, P));
            else PP_LAUNCH("ppm coldiv", st, coldiv_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gyp, nullptr, C, P, total, nullptr, gh, gl));
            rc = check_launch("ppm split(gy)");
            if (rc) return rc;
        }
        // (gS + gS^T)[i][j] = (G + G^T)[i][j] A'(S[i][j]) with G = gy^T v̂: ONE contraction with two operand sets accumulates
        // gy^T v̂ and v̂^T gy into the same tile (K = 2C), and its epilogue writes the symmetrised gradient directly as the planes
        // the gx contraction streams — gS itself, its store, and the transposing symmetrisation pass never exist.
        tc2::Operands sets[2] = {{gy_hi, gy_lo, v_hi, v_lo, C}, {v_hi, v_lo, gy_hi, gy_lo, C}};  // [C][P] tensors read MN-major
        sets[0].a_mn = sets[0].b_mn = sets[1].a_mn = sets[1].b_mn = true;
        rc = tc2::launch_tc2_sets("ppm gS (tcgen05)", B, P, P, sets, 2, TcStSymPlanes{sym_hi, sym_lo, sv.S, P, act}, st, false);
        if (rc > 0) return rc;
        if (rc == 0) {
            // gv̂[c][j] = Σ_i gy[c][i] A[i][j]; A is symmetric, so the B operand row j is row j of the saved relu^γ(S) planes
            rc = tc2::launch_tc2("ppm gvh (tcgen05)", B, C, P, P, gy_hi, gy_lo, sv.a_hi, sv.a_lo, TcStN{gvh, C, P}, st);
            if (rc) return rc < 0 ? PP_ERR_CUDA : rc;
            rc = tc2::launch_tc2("ppm gxh (tcgen05)", B, C, P, P, x_hi, x_lo, sym_hi, sym_lo, TcStN{gxh, C, P}, st);
            if (rc) return rc < 0 ? PP_ERR_CUDA : rc;
        } else {
            rc = launch_tc("ppm gS (tcgen05)", B, P, P, C, TcLdT{gyp, C, P}, TcLdT{sv.vh, C, P}, TcStGradS{gS, sv.S, P, act}, st);
            if (rc) return rc;
            rc = launch_tc("ppm gvh (tcgen05)", B, C, P, P, TcLdN{gyp, C, P}, TcLdActN{sv.S, P, act}, TcStN{gvh, C, P}, st);
            if (rc) return rc;
            rc = launch_tc("ppm gxh (tcgen05)", B, C, P, P, TcLdN{sv.xh, C, P}, TcLdSymN{gS, P}, TcStN{gxh, C, P}, st);
            if (rc) return rc;
        }
        if (regp) {
            PP_LAUNCH("ppm normbwd", st, normbwd_reg_kernel<false><<<dim3(nb.x, nb.y, 2), nt, 0, st>>>(
                normbwd_jobs(NormBwdJob{gxh, sv.xh, nullptr, sv.nx, d_feat_sim, nullptr, nullptr},
                             NormBwdJob{gvh, sv.vh, nullptr, sv.nv, d_val, nullptr, nullptr}), C, P));
        } else {
            PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(gxh, sv.xh, nullptr, sv.nx, C, P, d_feat_sim));
            PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(gvh, sv.vh, nullptr, sv.nv, C, P, d_val));
        }
        return check_launch("ppm normbwd(in)");
    }
    // gS[i][j] = (Σ_c gy[c][i] v̂[c][j]) A'(S[i][j])
    if (use_tensor_cores(P)) {
        rc = launch_tc("ppm gS (tcgen05)", B, P, P, C, TcLdT{gyp, C, P}, TcLdT{sv.vh, C, P}, TcStGradS{gS, sv.S, P, act}, st);
        if (rc) return rc;
        // gv̂[c][j] = Σ_i gy[c][i] A[i][j]; A is symmetric, so the B operand row j reads S[j][:] contiguously
        rc = launch_tc("ppm gvh (tcgen05)", B, C, P, P, TcLdN{gyp, C, P}, TcLdActN{sv.S, P, act}, TcStN{gvh, C, P}, st);
        if (rc) return rc;
        rc = launch_tc("ppm gxh (tcgen05)", B, C, P, P, TcLdN{sv.xh, C, P}, TcLdSymN{gS, P}, TcStN{gxh, C, P}, st);
        if (rc) return rc;
        PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(gxh, sv.xh, nullptr, sv.nx, C, P, d_feat_sim));
        PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(gvh, sv.vh, nullptr, sv.nv, C, P, d_val));
        return check_launch("ppm normbwd(in)");
    }
    // gS[i][j] = (Σ_c gy[c][i] v̂[c][j]) A'(S[i][j])
    rc = launch_bgemm("ppm gS", B, P, P, C, LoadColScaled{gyp, nullptr, C, P}, LoadColScaled{val, sv.nv, C, P},
                      EpGradS{gS, sv.S, P, act}, st);
    if (rc) return rc;
    rc = launch_bgemm("ppm gvh", B, C, P, P, LoadRowScaled{gyp, nullptr, C, P}, LoadActS_T{sv.S, P, act}, EpStore{gvh, C, P}, st);
    if (rc) return rc;
    rc = launch_bgemm("ppm gxh", B, C, P, P, LoadRowScaled{feat, sv.nx, C, P}, LoadSym{gS, P}, EpStore{gxh, C, P}, st);
    if (rc) return rc;
    PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(gxh, feat, sv.nx, sv.nx, C, P, d_feat_sim));
    PP_LAUNCH("ppm normbwd", st, normbwd_kernel<<<nb, nt, 0, st>>>(gvh, val, sv.nv, sv.nv, C, P, d_val));
    return check_launch("ppm normbwd(in)");
}

}  // extern "C"
