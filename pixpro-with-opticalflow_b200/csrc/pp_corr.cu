// pp_corr.cu — RAFT all-pairs correlation volume, its average-pool pyramid and the windowed bilinear lookup
// (the job the reference's missing `alt_cuda_corr` extension / its torch CorrBlock did).
//
// Reference restated (paths relative to the reference repo):
//   CorrBlock.corr      contrast/flow/corr.py:52-60    corr[b,i,j] = <fmap1[b,:,i], fmap2[b,:,j]> / sqrt(dim)
//   CorrBlock.__init__  contrast/flow/corr.py:12-28    pyramid: avg_pool2d(2, stride 2) over the target plane, num_levels-1 times
//   CorrBlock.__call__  contrast/flow/corr.py:30-50    per level a (2r+1)^2 window of bilinear samples around coords / 2^level
//   bilinear_sampler    contrast/flow/utils/utils.py:64-78   2*x/(W-1)-1 -> grid_sample(align_corners=True, zeros padding)
//
// The contraction runs on the tcgen05 tensor cores (3xTF32, fp32-accurate: pp_tc.cuh) with both feature maps read in
// place in their [B, D, h*w] layout and the 1/sqrt(D) division in the epilogue; pooling and lookup are HBM / latency
// bound gather kernels with the reference's own per-op rounding (normalise -> unnormalise -> floor -> weights -> fma chain).
#include <math.h>

#include "pp_common.cuh"
#include "pp_tc.cuh"
#include "pp_tc2.cuh"

namespace pp {

struct TcStDiv {
    static constexpr bool kAux = false;  // out[b][m][n..n+15] = v / s  (tensor / tensor true division of the reference: IEEE on every device)
    float* out;
    int M, N;
    float s;
    __device__ __forceinline__ void store4(int64_t b, int m, int n, float4 v) const {
        st4_guard(out + (b * M + m) * (int64_t)N + n, n, N, make_float4(__fdiv_rn(v.x, s), __fdiv_rn(v.y, s), __fdiv_rn(v.z, s), __fdiv_rn(v.w, s)));
    }
    __device__ __forceinline__ void store16(int64_t b, int m, int n, const float v[16]) const {
        float* q = out + (b * M + m) * (int64_t)N + n;
        if (n + 15 < N && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                reinterpret_cast<float4*>(q)[i] = make_float4(__fdiv_rn(v[4 * i], s), __fdiv_rn(v[4 * i + 1], s), __fdiv_rn(v[4 * i + 2], s),
                                                              __fdiv_rn(v[4 * i + 3], s));
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (n + i < N) q[i] = __fdiv_rn(v[i], s);
        }
    }
};

// CUDA-core contraction for planes too small for 128-wide tensor-core tiles (h*w < 128): one thread per output.
__global__ void __launch_bounds__(256) corr_small_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int D, int P,
                                                          float s, float* __restrict__ out, int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % P);
    const int64_t r = idx / P;
    const int i = (int)(r % P);
    const int64_t b = r / P;
    const float* a = f1 + b * D * (int64_t)P + i;
    const float* c = f2 + b * D * (int64_t)P + j;
    float acc = 0.0f;
    for (int d = 0; d < D; d++) acc = fmaf(__ldg(a + d * (int64_t)P), __ldg(c + d * (int64_t)P), acc);
    out[idx] = __fdiv_rn(acc, s);
}

// avg_pool2d(kernel 2, stride 2) of `planes` planes [h, w] -> [h/2, w/2]: ATen sums the window row-major
// ((a + b) + c) + d and divides by 4 (exact).  One thread per output; a warp reads two 256-byte row segments.
__global__ void __launch_bounds__(256) corr_pool_kernel(const float* __restrict__ in, int64_t planes, int h, int w, float* __restrict__ out) {
    const int ho = h >> 1, wo = w >> 1;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = planes * ho * wo;
    if (idx >= total) return;
    const int x = (int)(idx % wo);
    const int64_t r = idx / wo;
    const int y = (int)(r % ho);
    const int64_t p = r / ho;
    const float* q = in + (p * h + 2 * y) * (int64_t)w + 2 * x;
    const float2 top = __ldg(reinterpret_cast<const float2*>(q));  // 2x is even and w*... alignment: see launcher
    float a = top.x, b = top.y, c = __ldg(q + w), d = __ldg(q + w + 1);
    out[idx] = mul(add(add(add(a, b), c), d), 0.25f);
}
__global__ void __launch_bounds__(256) corr_pool_kernel_unaligned(const float* __restrict__ in, int64_t planes, int h, int w, float* __restrict__ out) {
    const int ho = h >> 1, wo = w >> 1;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = planes * ho * wo;
    if (idx >= total) return;
    const int x = (int)(idx % wo);
    const int64_t r = idx / wo;
    const int y = (int)(r % ho);
    const int64_t p = r / ho;
    const float* q = in + (p * h + 2 * y) * (int64_t)w + 2 * x;
    out[idx] = mul(add(add(add(__ldg(q), __ldg(q + 1)), __ldg(q + w)), __ldg(q + w + 1)), 0.25f);
}

constexpr int CORR_MAX_LEVELS = 8;
struct LookupArgs {
    const float* level[CORR_MAX_LEVELS];  // level l: [B*P, h_l, w_l]
    int hl[CORR_MAX_LEVELS], wl[CORR_MAX_LEVELS];
    const float* coords;  // [B, 2, h, w]: x then y, in level-0 pixels
    float* out;           // [B, L*K*K, h, w]
    int64_t B;
    int P, L, r, K;
    int rcp;  // div_mode of the tensor / python-scalar divisions in bilinear_sampler
};

// block = 256 threads = 32 consecutive query pixels of one sample x 8 lanes each, one pyramid level per blockIdx.y.
// The 8 lanes of a query walk its K*K window; results go through shared memory so that the output (channel-major:
// [L*K*K, h*w]) is written as 128-byte rows of 32 consecutive queries.
__global__ void __launch_bounds__(256) corr_lookup_kernel(LookupArgs a) {
    extern __shared__ float lk_smem[];  // [K*K][33]
    const int l = blockIdx.y, K = a.K, KK = K * K;
    const int64_t q0 = (int64_t)blockIdx.x * 32;  // first query of this block, flattened over (b, p)
    const int qq = threadIdx.x >> 3, sl = threadIdx.x & 7;
    const int64_t q = q0 + qq;
    const int64_t BP = a.B * a.P;
    if (q < BP) {
        const int64_t b = q / a.P;
        const int p = (int)(q - b * a.P);
        const float cx = __ldg(a.coords + (b * 2) * a.P + p), cy = __ldg(a.coords + (b * 2 + 1) * a.P + p);
        const float inv = 1.0f / (float)(1 << l);
        const float ccx = mul(cx, inv), ccy = mul(cy, inv);  // coords / 2**i: exact either way (power of two)
        const int H = a.hl[l], W = a.wl[l];
        const float* plane = a.level[l] + q * (int64_t)H * W;
        const ScalarDiv dw = {(float)(W - 1), 1.0f / (float)(W - 1), a.rcp}, dh = {(float)(H - 1), 1.0f / (float)(H - 1), a.rcp};
        const float half_w = (float)(W - 1) / 2.0f, half_h = (float)(H - 1) / 2.0f;
        for (int k = sl; k < KK; k += 8) {
            // corr.py:37-43: delta = stack(meshgrid(dy, dx), -1) is ADDED to (x, y): window index (i, j) shifts x by d[i], y by d[j]
            const int i = k / K, j = k - i * K;
            const float x = add(ccx, (float)(i - a.r)), y = add(ccy, (float)(j - a.r));
            const float gx = norm_coord(x, dw), gy = norm_coord(y, dh);             // utils.py:68-69
            const Taps t = make_taps(gx, gy, W, H, half_w, half_h);
            const float vnw = (t.inx0 && t.iny0) ? __ldg(plane + t.y0 * W + t.x0) : 0.0f;
            const float vne = (t.inx1 && t.iny0) ? __ldg(plane + t.y0 * W + t.x0 + 1) : 0.0f;
            const float vsw = (t.inx0 && t.iny1) ? __ldg(plane + (t.y0 + 1) * W + t.x0) : 0.0f;
            const float vse = (t.inx1 && t.iny1) ? __ldg(plane + (t.y0 + 1) * W + t.x0 + 1) : 0.0f;
            lk_smem[k * 33 + qq] = combine(t, vnw, vne, vsw, vse);
        }
    }
    __syncthreads();
    // write: thread -> (channel k, query qq); a warp covers one channel row of 32 queries
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t qw = q0 + lane;
    if (qw < BP) {
        const int64_t b = qw / a.P;
        const int p = (int)(qw - b * a.P);
        for (int k = wrp; k < KK; k += 8)
            a.out[(b * (int64_t)(a.L * KK) + (int64_t)l * KK + k) * a.P + p] = lk_smem[k * 33 + lane];
    }
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_corr_volume(const float* fmap1, const float* fmap2, int64_t B, int D, int h, int w, float* corr, void* stream) {
    PP_REQUIRE(B >= 0 && B <= 65535 && D > 0 && h > 0 && w > 0, "pp_corr_volume: bad shape B=%lld D=%d h=%d w=%d", (long long)B, D, h, w);
    if (B == 0) return PP_OK;
    PP_REQUIRE(fmap1 && fmap2 && corr, "pp_corr_volume: null pointer");
    const int P = h * w;
    const float s = sqrtf((float)D);  // torch.sqrt(torch.tensor(dim).float())
    cudaStream_t st = (cudaStream_t)stream;
    if (use_tensor_cores(P)) {
        // both feature maps read in place, MN-major ([D][h*w]: the channel index is the K line), by the TMA-fed kernel
        tc2::Operands o{fmap1, nullptr, fmap2, nullptr, D};
        o.a_mn = o.b_mn = true;
        const int rc = tc2::launch_tc2_sets("corr volume (tcgen05)", B, P, P, &o, 1, TcStDiv{corr, P, P, s}, st, false);
        if (rc >= 0) return rc;
    }
    if (use_tensor_cores(P))
        return launch_tc("corr volume (tcgen05)", B, P, P, D, TcLdT{fmap1, D, P}, TcLdT{fmap2, D, P}, TcStDiv{corr, P, P, s}, st);
    const int64_t total = B * P * (int64_t)P;
    PP_LAUNCH("corr volume", st, (corr_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(fmap1, fmap2, D, P, s, corr, total)));
    return check_launch("corr_small_kernel");
}

int pp_corr_pool(const float* in, int64_t planes, int h, int w, float* out, void* stream) {
    PP_REQUIRE(planes >= 0 && h >= 2 && w >= 2, "pp_corr_pool: bad shape planes=%lld h=%d w=%d", (long long)planes, h, w);
    if (planes == 0) return PP_OK;
    PP_REQUIRE(in && out, "pp_corr_pool: null pointer");
    const int64_t total = planes * (h >> 1) * (int64_t)(w >> 1);
    PP_REQUIRE((total + 255) / 256 < (1ll << 31), "pp_corr_pool: problem too large");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nb = (unsigned)((total + 255) / 256);
    if ((w & 1) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0)
        PP_LAUNCH("corr pool", st, (corr_pool_kernel<<<nb, 256, 0, st>>>(in, planes, h, w, out)));
    else
        PP_LAUNCH("corr pool", st, (corr_pool_kernel_unaligned<<<nb, 256, 0, st>>>(in, planes, h, w, out)));
    return check_launch("corr_pool_kernel");
}

int pp_corr_lookup(const float* const* levels, int num_levels, const float* coords, int64_t B, int h, int w, int radius, int div_mode,
                   float* out, void* stream) {
    PP_REQUIRE(num_levels >= 1 && num_levels <= CORR_MAX_LEVELS && radius >= 0 && radius <= 8, "pp_corr_lookup: bad levels %d / radius %d", num_levels, radius);
    PP_REQUIRE(B >= 0 && h > 0 && w > 0, "pp_corr_lookup: bad shape");
    if (B == 0) return PP_OK;
    PP_REQUIRE(levels && coords && out, "pp_corr_lookup: null pointer");
    LookupArgs a;
    int hl = h, wl = w;
    for (int l = 0; l < num_levels; l++) {
        PP_REQUIRE(levels[l] != nullptr, "pp_corr_lookup: level %d is null", l);
        PP_REQUIRE(hl >= 2 && wl >= 2, "pp_corr_lookup: level %d is %dx%d (the reference divides by size-1)", l, hl, wl);
        a.level[l] = levels[l];
        a.hl[l] = hl; a.wl[l] = wl;
        hl >>= 1; wl >>= 1;
    }
    a.coords = coords; a.out = out; a.B = B; a.P = h * w; a.L = num_levels; a.r = radius; a.K = 2 * radius + 1; a.rcp = div_mode;
    const int64_t BP = B * (int64_t)a.P;
    dim3 grid((unsigned)((BP + 31) / 32), (unsigned)num_levels);
    const size_t smem = (size_t)a.K * a.K * 33 * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    PP_LAUNCH("corr lookup", st, (corr_lookup_kernel<<<grid, 256, smem, st>>>(a)));
    return check_launch("corr_lookup_kernel");
}

}  // extern "C"
