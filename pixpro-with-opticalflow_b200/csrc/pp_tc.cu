// pp_tc.cu — tcgen05 batched GEMM entry used by tests / microbenchmarks:
// C[b] = A[b] * B[b]^T in fp32 accuracy (3xTF32), A [batch,M,K], B [batch,N,K], C [batch,M,N].
#include "pp_common.cuh"
#include "pp_tc.cuh"
#include "pp_tc2.cuh"

namespace pp {
namespace tc {

struct LoadRowK {  // row-major [rows, K], k contiguous
    static constexpr bool kRowMajorK = true;
    const float* p;
    int rows, K;
    __device__ __forceinline__ float4 load4(int64_t b, int row, int k) const {
        if (row >= rows || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float* q = p + (b * rows + row) * (int64_t)K + k;
        if (k + 3 < K && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(q));
        float4 v;
        v.x = __ldg(q);
        v.y = k + 1 < K ? __ldg(q + 1) : 0.f;
        v.z = k + 2 < K ? __ldg(q + 2) : 0.f;
        v.w = k + 3 < K ? __ldg(q + 3) : 0.f;
        return v;
    }
};

struct StoreC {
    static constexpr bool kAux = false;
    float* c;
    int M, N;
    __device__ __forceinline__ void store4(int64_t b, int m, int n, float4 v) const { st4_guard(c + (b * M + m) * (int64_t)N + n, n, N, v); }
    __device__ __forceinline__ void store16(int64_t b, int m, int n, const float v[16]) const {
        float* q = c + (b * M + m) * (int64_t)N + n;
        if (n + 15 < N && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 4; i++) reinterpret_cast<float4*>(q)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            return;
        }
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (n + i < N) q[i] = v[i];
    }
};

}  // namespace tc
}  // namespace pp

using namespace pp;

extern "C" int pp_tc_gemm_nt(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, void* stream) {
    PP_REQUIRE(A && B && C, "pp_tc_gemm_nt: null pointer");
    PP_REQUIRE(batch > 0 && batch <= 65535 && M > 0 && N > 0 && K > 0, "pp_tc_gemm_nt: bad shape");
    return launch_tc("tc_gemm_nt", batch, M, N, K, tc::LoadRowK{A, M, K}, tc::LoadRowK{B, N, K}, tc::StoreC{C, M, N},
                     (cudaStream_t)stream);
}

// The same product through the TMA-fed kernel (pp_tc2.cuh): the operands are split once into hi / lo planes in `workspace`
// (pp_tc_gemm_nt_workspace bytes), then streamed by TMA.  Falls back to the kernel above when K % 4 != 0.
extern "C" int64_t pp_tc_gemm_nt_workspace(int64_t batch, int M, int N, int K) {
    return 2 * batch * ((int64_t)M + N) * K * (int64_t)sizeof(float);
}

extern "C" int pp_tc_gemm_nt_ws(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, void* workspace, void* stream) {
    PP_REQUIRE(A && B && C && workspace, "pp_tc_gemm_nt_ws: null pointer");
    PP_REQUIRE(batch > 0 && batch <= 65535 && M > 0 && N > 0 && K > 0, "pp_tc_gemm_nt_ws: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    float* a_hi = (float*)workspace;
    float* a_lo = a_hi + batch * (int64_t)M * K;
    float* b_hi = a_lo + batch * (int64_t)M * K;
    float* b_lo = b_hi + batch * (int64_t)N * K;
    static const int planes = [] { const char* e = getenv("PIXPRO_B200_TC2_PLANES"); return (e && e[0] == '1') ? 1 : 0; }();
    if (K % 4 == 0 && tc2::applicable(A, B, workspace, C)) {
        int rc;
        if (planes) {  // A/B switch: operands pre-split into hi / lo planes by a separate pass
            rc = tc2::launch_split(A, batch * (int64_t)M * K, a_hi, a_lo, st);
            if (rc) return rc;
            rc = tc2::launch_split(B, batch * (int64_t)N * K, b_hi, b_lo, st);
            if (rc) return rc;
            rc = tc2::launch_tc2("tc2_gemm_nt", batch, M, N, K, a_hi, a_lo, b_hi, b_lo, tc::StoreC{C, M, N}, st);
        } else {       // default: fp32 operands streamed as they are, split by the kernel's converter warps
            rc = tc2::launch_tc2("tc2_gemm_nt", batch, M, N, K, A, nullptr, B, nullptr, tc::StoreC{C, M, N}, st);
        }
        if (rc >= 0) return rc;
    }
    return launch_tc("tc_gemm_nt", batch, M, N, K, tc::LoadRowK{A, M, K}, tc::LoadRowK{B, N, K}, tc::StoreC{C, M, N}, st);
}

// General layouts through the TMA-fed kernel: a_mn / b_mn = the operand is stored MN-major, i.e. A as [batch][K][M] (resp. B as
// [batch][K][N]) — how a [C, P] feature map serves as the P x C operand of the similarity contraction without a transposing pass.
// fp32 operands are streamed as they are and split into hi / lo by the kernel's converter warps.  No fallback: returns
// PP_ERR_INVALID when the shape is not streamable (K, and the contiguous extent of an MN-major operand, must be multiples of 4).
extern "C" int pp_tc_gemm_ws(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, int a_mn, int b_mn, void* stream) {
    return pp_tc_gemm_ex(A, B, C, batch, M, N, K, a_mn, b_mn, 0, 0, 1, stream);
}

// The same with padded operands and sums over batch entries: a_pitch / b_pitch = allocated length (floats, % 4 == 0; 0 = dense) of
// each operand's contiguous dimension (K for a K-major operand, M resp. N for an MN-major one) — how a [C, 49] map copied to a
// pitch of 52 floats becomes streamable; kb > 1: C[g] = sum over the kb batch entries of group g of A[b] B[b]^T, C is
// [ceil(batch / kb)][M][N] (the weight gradient of the 1x1 convolution sums per-sample products this way).
extern "C" int pp_tc_gemm_ex(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, int a_mn, int b_mn, int a_pitch,
                             int b_pitch, int kb, void* stream) {
    PP_REQUIRE(A && B && C, "pp_tc_gemm_ex: null pointer");
    PP_REQUIRE(batch > 0 && batch <= 65535 && M > 0 && N > 0 && K > 0 && kb >= 1 && a_pitch >= 0 && b_pitch >= 0, "pp_tc_gemm_ex: bad shape");
    tc2::Operands o{A, nullptr, B, nullptr, K};
    o.a_mn = a_mn != 0;
    o.b_mn = b_mn != 0;
    o.a_pitch = a_pitch;
    o.b_pitch = b_pitch;
    const int rc = tc2::launch_tc2_sets("tc2_gemm", batch, M, N, &o, 1, tc::StoreC{C, M, N}, (cudaStream_t)stream, false, kb);
    PP_REQUIRE(rc >= 0, "pp_tc_gemm_ex: shape / alignment not streamable by TMA (contiguous extents or pitches %% 4, 16-byte aligned pointers)");
    return rc;
}
