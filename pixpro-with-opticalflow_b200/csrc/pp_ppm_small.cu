// pp_ppm_small.cu — Pixel Propagation Module for small grids (P = G*G <= 64, e.g. the 7x7
// C5 grid of the published 224x224 crops): ONE thread block per sample keeps x̂, v̂, the
// P x P similarity and every intermediate in shared memory, so forward is one launch and
// backward is one launch, and HBM sees only the compulsory traffic (read feat, val [, out, g],
// write out / gradients + the small `saved` block).  fp32 CUDA-core arithmetic: at P = 49 the
// contractions are 49x49x256 per sample — far below the size where tensor cores pay off
// (BASELINE.json north_star: tcgen05 "only at feature grids large enough to be dense").
//
// Reference restated: PixPro.featprop + F.normalize, contrast/models/PixPro.py:339-363,380.
// Formulas and the `saved` layout ([nx | nv | ny | S]) are those of pp_ppm.cu (generic path),
// which remains the fallback for larger grids.
#include <math.h>

#include "pp_common.cuh"
#include "pp_ppm.cuh"
#include "pp_small.cuh"

namespace pp {

struct SmallArgs {
    const float *feat, *val, *out_in, *g;
    float *out, *d_feat, *d_val;
    float *nx, *nv, *ny, *S;  // saved
    int C, P, final_norm;
    Act act;
};

// 512 threads per block: one block per sample keeps one SM busy, so its 16 warps (instead of 8) are
// what hides the shared-memory and FMA latencies of the short serial phases (measured: issue slots
// 30 % busy with 8 warps).  The two contractions split their work over the two 256-thread halves
// (gram_tile_split: channel halves; row_times_mat_split: column halves).  1024 threads (with PMAX = 52 to make room for the
// two extra gram partials) were measured in round 2: no change (backward 61-64 us, forward 37 us) — the phases are bound by
// their shared-memory instruction count and the barriers between them, not by the number of warps.
constexpr int PT = 512;
constexpr int PRED = PT / 64;  // partial rows of col_reduce

// shared memory (floats): nbuf x [C*P + SLACK] | nmat x [PMAX*PS] | norms 3*PMAX | red PRED*PMAX | dot PMAX | rcp 3*PMAX
// | gram partials (PT/256 - 1) x [PMAX*PS]
__host__ __device__ inline size_t small_smem_bytes(int C, int P, bool bwd) {
    size_t f = (size_t)(bwd ? 3 : 2) * ((size_t)C * P + SLACK) + (size_t)(bwd ? 3 : 1) * PMAX * PS + (7 + PRED) * PMAX +
               (size_t)(PT / 256 - 1) * PMAX * PS;
    return f * sizeof(float);
}

__global__ void __launch_bounds__(PT) ppm_fwd_small_kernel(SmallArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int C = a.C, P = a.P, CP = C * P;
    float* xs = smem;
    float* vs = xs + CP + SLACK;
    float* Sm = vs + CP + SLACK;
    float* nrm = Sm + PMAX * PS;  // [0] nx, [1] nv, [2] ny
    float* red = nrm + 3 * PMAX;
    float* rcp = red + PRED * PMAX;  // 3*PMAX reciprocals
    float* gpart = rcp + 4 * PMAX;   // gram partials of the upper thread halves (after the unused dot slot)
    const int64_t b = blockIdx.x;
    stage_dense<PT>(xs, a.feat + b * (int64_t)CP, CP);
    stage_dense<PT>(vs, a.val + b * (int64_t)CP, CP);
    for (int e = threadIdx.x; e < PMAX * PS; e += PT) Sm[e] = 0.0f;
    if (threadIdx.x < SLACK) { xs[CP + threadIdx.x] = 0.0f; vs[CP + threadIdx.x] = 0.0f; }
    __syncthreads();
    col_reduce<PT>(C, P, red, nrm, [&](int c, int i) { float t = xs[c * P + i]; return t * t; });
    col_reduce<PT>(C, P, red, nrm + PMAX, [&](int c, int i) { float t = vs[c * P + i]; return t * t; });
    if (threadIdx.x < P) {
        float n0 = fmaxf(sqrtf(nrm[threadIdx.x]), kNormEps), n1 = fmaxf(sqrtf(nrm[PMAX + threadIdx.x]), kNormEps);
        nrm[threadIdx.x] = n0;
        nrm[PMAX + threadIdx.x] = n1;
        a.nx[b * P + threadIdx.x] = n0;
        a.nv[b * P + threadIdx.x] = n1;
    }
    __syncthreads();
    // reciprocals once per column; x * (1/n) instead of x / n (floating tolerance, 12x fewer instructions)
    if (threadIdx.x < P) {
        rcp[threadIdx.x] = 1.0f / nrm[threadIdx.x];
        rcp[PMAX + threadIdx.x] = 1.0f / nrm[PMAX + threadIdx.x];
    }
    __syncthreads();
    for_each_ci<PT>(C, P, [&](int e, int, int i) {  // x̂, v̂ in place (PixPro.py:344,348)
        xs[e] = xs[e] * rcp[i];
        vs[e] = vs[e] * rcp[PMAX + i];
    });
    __syncthreads();
    // S = x̂ᵀx̂ (:354); A = relu^γ(S) kept in smem (bitwise symmetric), raw S saved for backward
    float* Sg = a.S + b * (int64_t)P * P;
    const Act act = a.act;
    gram_tile_split<PT>(xs, xs, C, P, gpart, [&](int i, int j, float s) {
        Sg[i * P + j] = s;
        Sm[i * PS + j] = act.f(s);
    });
    __syncthreads();
    // Y = v̂ Aᵀ (:361), A symmetric -> rows of A; result overwrites x̂
    row_times_mat_split<PT>(vs, Sm, C, P, xs);
    __syncthreads();
    float* o = a.out + b * (int64_t)CP;
    if (a.final_norm) {  // :380
        col_reduce<PT>(C, P, red, nrm + 2 * PMAX, [&](int c, int i) { float t = xs[c * P + i]; return t * t; });
        if (threadIdx.x < P) {
            float n2 = fmaxf(sqrtf(nrm[2 * PMAX + threadIdx.x]), kNormEps);
            nrm[2 * PMAX + threadIdx.x] = n2;
            a.ny[b * P + threadIdx.x] = n2;
        }
        if (threadIdx.x < P) rcp[2 * PMAX + threadIdx.x] = 1.0f / nrm[2 * PMAX + threadIdx.x];
        __syncthreads();
        for_each_ci<PT>(C, P, [&](int e, int, int i) { o[e] = xs[e] * rcp[2 * PMAX + i]; });
    } else {
        for (int e = threadIdx.x; e < CP; e += PT) o[e] = xs[e];
    }
}

__global__ void __launch_bounds__(PT) ppm_bwd_small_kernel(SmallArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int C = a.C, P = a.P, CP = C * P;
    float* xs = smem;              // x̂
    float* vs = xs + CP + SLACK;   // v̂
    float* gs = vs + CP + SLACK;   // g -> gy -> gv̂ -> gx̂
    float* Sm = gs + CP + SLACK;   // raw S
    float* Am = Sm + PMAX * PS;    // A = relu^γ(S)
    float* Gm = Am + PMAX * PS;    // gS, then gS + gSᵀ
    float* nrm = Gm + PMAX * PS;
    float* red = nrm + 3 * PMAX;
    float* dot = red + PRED * PMAX;
    float* rcp = dot + PMAX;  // 3*PMAX reciprocals of the norms
    float* gpart = rcp + 3 * PMAX;  // gram partials of the upper thread halves
    const int64_t b = blockIdx.x;
    const int64_t off = b * (int64_t)CP;
    const float* yh = a.out_in + off;
    stage_dense<PT>(xs, a.feat + off, CP);
    stage_dense<PT>(vs, a.val + off, CP);
    stage_dense<PT>(gs, a.g + off, CP);
    if (threadIdx.x < P) {
        rcp[threadIdx.x] = 1.0f / a.nx[b * P + threadIdx.x];
        rcp[PMAX + threadIdx.x] = 1.0f / a.nv[b * P + threadIdx.x];
        rcp[2 * PMAX + threadIdx.x] = a.final_norm ? 1.0f / a.ny[b * P + threadIdx.x] : 1.0f;
    }
    for (int e = threadIdx.x; e < 3 * PMAX * PS; e += PT) Sm[e] = 0.0f;  // Sm, Am, Gm
    if (threadIdx.x < SLACK) { xs[CP + threadIdx.x] = 0.0f; vs[CP + threadIdx.x] = 0.0f; gs[CP + threadIdx.x] = 0.0f; }
    __syncthreads();
    const Act act = a.act;
    {
        const float* Sg = a.S + b * (int64_t)P * P;
        int e = threadIdx.x, i = e / P, j = e - i * P;
        const int dr = PT / P, dj = PT - dr * P;
        for (; e < P * P; e += PT) {
            float s = __ldg(Sg + e);
            Sm[i * PS + j] = s;
            Am[i * PS + j] = act.f(s);
            i += dr; j += dj;
            if (j >= P) { j -= P; i += 1; }
        }
    }
    for_each_ci<PT>(C, P, [&](int e, int, int i) {
        xs[e] = xs[e] * rcp[i];
        vs[e] = vs[e] * rcp[PMAX + i];
    });
    __syncthreads();
    if (a.final_norm) {  // gy = (g − ŷ (g·ŷ)) / ‖Y‖
        col_reduce<PT>(C, P, red, dot, [&](int c, int i) { return gs[c * P + i] * __ldg(yh + c * P + i); });
        for_each_ci<PT>(C, P, [&](int e, int, int i) { gs[e] = (gs[e] - __ldg(yh + e) * dot[i]) * rcp[2 * PMAX + i]; });
        __syncthreads();
    }
    // gS[i][j] = (Σ_c gy[c][i] v̂[c][j]) A'(S[i][j])
    gram_tile_split<PT>(gs, vs, C, P, gpart, [&](int i, int j, float t) { Gm[i * PS + j] = t * act.df(Sm[i * PS + j]); });
    __syncthreads();
    // symmetrise gS in place (each unordered pair is owned by one thread)
    {
        int e = threadIdx.x, i = e / P, j = e - i * P;
        const int dr = PT / P, dj = PT - dr * P;
        for (; e < P * P; e += PT) {
            if (i < j) {
                float t = Gm[i * PS + j] + Gm[j * PS + i];
                Gm[i * PS + j] = t;
                Gm[j * PS + i] = t;
            } else if (i == j) {
                Gm[i * PS + i] = 2.0f * Gm[i * PS + i];
            }
            i += dr; j += dj;
            if (j >= P) { j -= P; i += 1; }
        }
    }
    // gv̂[c][j] = Σ_i gy[c][i] A[i][j]: rows are thread-private and fully accumulated in registers
    // before the write-back, so gv̂ overwrites gy in place.
    row_times_mat_split<PT>(gs, Am, C, P, gs);
    __syncthreads();
    // d_val = (gv̂ − v̂ (gv̂·v̂)) / ‖v‖
    col_reduce<PT>(C, P, red, dot, [&](int c, int i) { return gs[c * P + i] * vs[c * P + i]; });
    for_each_ci<PT>(C, P, [&](int e, int, int i) { a.d_val[off + e] = (gs[e] - vs[e] * dot[i]) * rcp[PMAX + i]; });
    __syncthreads();
    // gx̂[c][i] = Σ_j x̂[c][j] (gS + gSᵀ)[j][i]  -> into gs (free now)
    row_times_mat_split<PT>(xs, Gm, C, P, gs);
    __syncthreads();
    col_reduce<PT>(C, P, red, dot, [&](int c, int i) { return gs[c * P + i] * xs[c * P + i]; });
    for_each_ci<PT>(C, P, [&](int e, int, int i) { a.d_feat[off + e] = (gs[e] - xs[e] * dot[i]) * rcp[i]; });
}

// opt in to > 48 KB of dynamic shared memory, once per process, with the result checked
static int ensure_small_attrs() {
    static unsigned long long opted_f = 0, opted_b = 0;  // one bit per device
    cudaError_t e1 = smem_opt_in(ppm_fwd_small_kernel, 227 * 1024, opted_f);
    cudaError_t e2 = smem_opt_in(ppm_bwd_small_kernel, 227 * 1024, opted_b);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        set_error("ppm small: cudaFuncSetAttribute failed: %s / %s", cudaGetErrorString(e1), cudaGetErrorString(e2));
        return PP_ERR_CUDA;
    }
    return PP_OK;
}

bool ppm_small_supported(int C, int P) {
    return P <= PMAX && ((C * P) % 4 == 0) && C <= PT && PT % C == 0 && small_smem_bytes(C, P, true) <= 226 * 1024;
}

int ppm_fwd_small(const float* feat, const float* val, int64_t B, int C, int P, Act act, int final_norm, float* out,
                  float* nx, float* nv, float* ny, float* S, cudaStream_t st) {
    size_t smem = small_smem_bytes(C, P, false);
    int rc = ensure_small_attrs();
    if (rc) return rc;
    SmallArgs a{};
    a.feat = feat; a.val = val; a.out = out; a.nx = nx; a.nv = nv; a.ny = ny; a.S = S;
    a.C = C; a.P = P; a.final_norm = final_norm; a.act = act;
    PP_LAUNCH("ppm_fwd_small", st, ppm_fwd_small_kernel<<<(unsigned)B, PT, smem, st>>>(a));
    return check_launch("ppm_fwd_small_kernel");
}

int ppm_bwd_small(const float* feat, const float* val, const float* out, const float* g, int64_t B, int C, int P, Act act,
                  int final_norm, const float* nx, const float* nv, const float* ny, const float* S, float* d_feat,
                  float* d_val, cudaStream_t st) {
    size_t smem = small_smem_bytes(C, P, true);
    int rc = ensure_small_attrs();
    if (rc) return rc;
    SmallArgs a{};
    a.feat = feat; a.val = val; a.out_in = out; a.g = g; a.d_feat = d_feat; a.d_val = d_val;
    a.nx = const_cast<float*>(nx); a.nv = const_cast<float*>(nv); a.ny = const_cast<float*>(ny); a.S = const_cast<float*>(S);
    a.C = C; a.P = P; a.final_norm = final_norm; a.act = act;
    PP_LAUNCH("ppm_bwd_small", st, ppm_bwd_small_kernel<<<(unsigned)B, PT, smem, st>>>(a));
    return check_launch("ppm_bwd_small_kernel");
}

}  // namespace pp
