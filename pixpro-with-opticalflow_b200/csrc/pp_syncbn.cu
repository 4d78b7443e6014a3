// pp_syncbn.cu — synchronised batch normalisation of the pre-training step in THREE launches per direction.
//
// Reference: the model converts every BatchNorm of encoder / projector (and their momentum twins) to
// torch.nn.SyncBatchNorm (contrast/models/PixPro.py:289-292, 315-317) and runs under DDP (main_pretrain.py:78).
// torch's SyncBatchNorm issues, per layer and direction, ~10 small launches and one all_gather from a Python
// autograd.Function (batch_norm_stats, empty/cat/split glue, all_gather, batch_norm_gather_stats_with_counts,
// batch_norm_elemt; backward: batch_norm_backward_reduce, all_reduce, batch_norm_backward_elemt + divisions); with 318
// layer calls per step the multi-GPU step is bound by the host's issue rate, not by the GPUs or the links
// (profiles/r01_s3_pretrain_ddp.txt: 90 ms on one GPU, 150 ms on two).  Here, per layer:
//   forward : pp_bn_stats    per-channel local (mean, M2, count) -> [2C+1] fp32, one launch
//             (torch.distributed.all_gather_into_tensor of that ONE vector — the only collective)
//             pp_bn_apply    global mean / invstd from the gathered rows, running-stat update, y = (x - mean) invstd w + b
//   backward: pp_bn_bwd_stats   per-channel (sum dy, sum dy (x - mean)) -> [2C]; also grad_weight, grad_bias (local sums,
//             as torch: DDP reduces parameter gradients later)
//             (all_reduce(SUM) of [2C])
//             pp_bn_bwd_apply   dx = (dy - mean_dy - (x - mean) invstd^2 mean_dy_xmu) invstd w
// Layouts: NCHW-contiguous and channels_last (NHWC memory, what the trainer uses); dtypes fp32 and bf16 (autocast);
// statistics always fp32, deterministic (fixed partition, partials reduced in order by the last block to finish).
#include <cuda_bf16.h>
#include <math.h>

#include "pp_common.cuh"

namespace pp {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_SLABS = 64;

template <class T>
__device__ __forceinline__ float bn_ld(const T* p);
template <>
__device__ __forceinline__ float bn_ld<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float bn_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <class T>
__device__ __forceinline__ void bn_st(T* p, float v);
template <>
__device__ __forceinline__ void bn_st<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void bn_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// Element addressing of a logical [N, C, HW] tensor.  NHWC: ((n HW + s) C + c); NCHW: ((n C + c) HW + s).
// Work decomposition common to all four kernels: rows r = n HW + s in [0, R); a "slab" is a contiguous range of rows.
// NHWC kernels: block = 32 channel lanes x 8 row warps... lanes walk channels (contiguous), warps walk rows.
// NCHW kernels: block = one channel x one slab; threads walk s (contiguous) inside each n.
struct BnShape {
    int64_t N;
    int C, HW;
    int nhwc;
};

// ---- forward statistics -------------------------------------------------------------------------------------
// partial[slab][2][C]: sum and sum of squares of (x - pivot_c), pivot_c = the channel's first element (a shift that
// removes the E[x^2] - E[x]^2 cancellation for channels whose mean is large against their deviation); finalised by
// the last block into out[0..C) = local mean, out[C..2C) = M2 = sum (x - mean_local)^2, out[2C] = local count — the
// triple torch's SyncBatchNorm gathers too; ranks are combined in bn_apply_kernel with the parallel-variance formula,
// so no rank ever forms a raw second moment (fp32 would lose the variance of a channel with |mean| >> deviation).
template <class T, bool NHWC>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const T* __restrict__ x, BnShape sh, int nslab, float* __restrict__ partial,
                                                               unsigned int* __restrict__ ticket, float* __restrict__ out) {
    const int C = sh.C, HW = sh.HW;
    const int64_t R = sh.N * HW;
    const int slab = blockIdx.y;
    const int64_t r0 = R * slab / nslab, r1 = R * (slab + 1) / nslab;
    __shared__ float red[2][8][33];
    __shared__ bool last;
    if (NHWC) {
        const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
        float s = 0.f, q = 0.f;
        if (c < C) {
            const float pv = bn_ld(x + c);
            for (int64_t r = r0 + w; r < r1; r += 8) {
                const float d = bn_ld(x + r * C + c) - pv;
                s += d;
                q = fmaf(d, d, q);
            }
        }
        red[0][w][threadIdx.x & 31] = s;
        red[1][w][threadIdx.x & 31] = q;
        __syncthreads();
        if (w == 0 && c < C) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { ts += red[0][k][threadIdx.x]; tq += red[1][k][threadIdx.x]; }
            partial[((int64_t)slab * 2 + 0) * C + c] = ts;
            partial[((int64_t)slab * 2 + 1) * C + c] = tq;
        }
    } else {
        const int c = blockIdx.x;
        const float pv = bn_ld(x + (int64_t)c * HW);
        float s = 0.f, q = 0.f;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += BN_THREADS) {
            const int64_t n = r / HW;
            const int sp = (int)(r - n * HW);
            const float d = bn_ld(x + (n * C + c) * HW + sp) - pv;
            s += d;
            q = fmaf(d, d, q);
        }
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5][0] = s; red[1][threadIdx.x >> 5][0] = q; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { ts += red[0][k][0]; tq += red[1][k][0]; }
            partial[((int64_t)slab * 2 + 0) * C + c] = ts;
            partial[((int64_t)slab * 2 + 1) * C + c] = tq;
        }
    }
    // last block to finish reduces the partials in slab order (deterministic)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        last = atomicAdd(ticket, 1u) == total - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const float cnt = (float)R;
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        float ts = 0.f, tq = 0.f;
        for (int k = 0; k < nslab; k++) {
            ts += __ldcg(partial + ((int64_t)k * 2 + 0) * C + c);
            tq += __ldcg(partial + ((int64_t)k * 2 + 1) * C + c);
        }
        const float pv = bn_ld(x + (NHWC ? (int64_t)c : (int64_t)c * HW));
        // about the pivot: mean = pv + ts/n;  sum (x-mean)^2 = tq - ts^2/n
        const float dm = ts / cnt, mean = pv + dm;
        out[c] = mean;
        out[C + c] = fmaxf(tq - ts * dm, 0.f);
    }
    if (threadIdx.x == 0) {
        out[2 * C] = cnt;
        *ticket = 0;  // ready for the next launch on this stream
    }
}

// ---- forward apply ------------------------------------------------------------------------------------------
// stats[r][0..C) = mean of rank r, [C..2C) = its M2, [2C] = its count (after the all-gather; nranks rows of 2C+1).
struct BnApplyArgs {
    const float *stats, *weight, *bias;   // weight / bias may be null (affine=False)
    int nranks;
    float *running_mean, *running_var;    // may be null
    float *save_mean, *save_invstd;       // [C], for the backward
    float eps, momentum;
};
__device__ __forceinline__ void bn_moments(const float* stats, int nranks, int C, int c, float eps, float& mean, float& invstd, float& var,
                                           float& cnt) {
    const int S = 2 * C + 1;
    cnt = 0.f;
    float sm = 0.f;
    for (int r = 0; r < nranks; r++) {
        const float n = stats[r * S + 2 * C];
        cnt += n;
        sm = fmaf(n, stats[r * S + c], sm);
    }
    mean = sm / cnt;
    float m2 = 0.f;
    for (int r = 0; r < nranks; r++) {  // Chan et al.: M2 = sum_r M2_r + n_r (mean_r - mean)^2
        const float d = stats[r * S + c] - mean;
        m2 += stats[r * S + C + c] + stats[r * S + 2 * C] * d * d;
    }
    var = m2 / cnt;  // biased
    invstd = rsqrtf(var + eps);
}
template <class T, bool NHWC>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, BnShape sh, int nslab, BnApplyArgs a) {
    const int C = sh.C, HW = sh.HW;
    const int64_t R = sh.N * HW;
    const int slab = blockIdx.y;
    const int64_t r0 = R * slab / nslab, r1 = R * (slab + 1) / nslab;
    const int c = NHWC ? blockIdx.x * 32 + (threadIdx.x & 31) : blockIdx.x;
    if (c >= C) return;
    float mean, invstd, var, cnt;
    bn_moments(a.stats, a.nranks, C, c, a.eps, mean, invstd, var, cnt);
    const float g = a.weight ? __ldg(a.weight + c) * invstd : invstd;
    const float sft = (a.bias ? __ldg(a.bias + c) : 0.f) - mean * g;
    if (slab == 0 && (NHWC ? (threadIdx.x >> 5) == 0 : threadIdx.x == 0)) {
        a.save_mean[c] = mean;
        a.save_invstd[c] = invstd;
        if (c == 0) a.save_invstd[C] = cnt;  // total element count, read by the backward
        if (a.running_mean) {  // F.batch_norm: running = (1 - m) running + m stat, unbiased variance
            a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
            const float unb = cnt > 1.f ? var * cnt / (cnt - 1.f) : var;
            a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unb;
        }
    }
    if (NHWC) {
        for (int64_t r = r0 + (threadIdx.x >> 5); r < r1; r += 8) bn_st(y + r * C + c, fmaf(bn_ld(x + r * C + c), g, sft));
    } else {
        for (int64_t r = r0 + threadIdx.x; r < r1; r += BN_THREADS) {
            const int64_t n = r / HW;
            const int64_t o = (n * C + c) * HW + (r - n * HW);
            bn_st(y + o, fmaf(bn_ld(x + o), g, sft));
        }
    }
}

// ---- backward statistics: out[0..C) = sum dy, out[C..2C) = sum dy (x - mean); grad_bias = sum dy, grad_weight = sum dy (x-mean) invstd
template <class T, bool NHWC>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_stats_kernel(const T* __restrict__ dy, const T* __restrict__ x, BnShape sh, int nslab,
                                                                   const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                                                   float* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                                   float* __restrict__ out, float* __restrict__ grad_weight,
                                                                   float* __restrict__ grad_bias) {
    const int C = sh.C, HW = sh.HW;
    const int64_t R = sh.N * HW;
    const int slab = blockIdx.y;
    const int64_t r0 = R * slab / nslab, r1 = R * (slab + 1) / nslab;
    __shared__ float red[2][8][33];
    __shared__ bool last;
    if (NHWC) {
        const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
        float s = 0.f, q = 0.f;
        if (c < C) {
            const float mean = __ldg(save_mean + c);
            for (int64_t r = r0 + w; r < r1; r += 8) {
                const float g = bn_ld(dy + r * C + c);
                s += g;
                q = fmaf(g, bn_ld(x + r * C + c) - mean, q);
            }
        }
        red[0][w][threadIdx.x & 31] = s;
        red[1][w][threadIdx.x & 31] = q;
        __syncthreads();
        if (w == 0 && c < C) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { ts += red[0][k][threadIdx.x]; tq += red[1][k][threadIdx.x]; }
            partial[((int64_t)slab * 2 + 0) * C + c] = ts;
            partial[((int64_t)slab * 2 + 1) * C + c] = tq;
        }
    } else {
        const int c = blockIdx.x;
        const float mean = __ldg(save_mean + c);
        float s = 0.f, q = 0.f;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += BN_THREADS) {
            const int64_t n = r / HW;
            const int64_t o = (n * C + c) * HW + (r - n * HW);
            const float g = bn_ld(dy + o);
            s += g;
            q = fmaf(g, bn_ld(x + o) - mean, q);
        }
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5][0] = s; red[1][threadIdx.x >> 5][0] = q; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { ts += red[0][k][0]; tq += red[1][k][0]; }
            partial[((int64_t)slab * 2 + 0) * C + c] = ts;
            partial[((int64_t)slab * 2 + 1) * C + c] = tq;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        last = atomicAdd(ticket, 1u) == total - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        float ts = 0.f, tq = 0.f;
        for (int k = 0; k < nslab; k++) {
            ts += __ldcg(partial + ((int64_t)k * 2 + 0) * C + c);
            tq += __ldcg(partial + ((int64_t)k * 2 + 1) * C + c);
        }
        out[c] = ts;
        out[C + c] = tq;
        if (grad_bias) grad_bias[c] = ts;
        if (grad_weight) grad_weight[c] = tq * __ldg(save_invstd + c);
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// ---- backward apply: sums[0..C) = total sum dy, [C..2C) = total sum dy (x - mean) (after the all-reduce), count = total count
template <class T, bool NHWC>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, BnShape sh,
                                                                   int nslab, const float* __restrict__ save_mean,
                                                                   const float* __restrict__ save_invstd, const float* __restrict__ weight,
                                                                   const float* __restrict__ sums, const float* __restrict__ total_count) {
    const int C = sh.C, HW = sh.HW;
    const int64_t R = sh.N * HW;
    const int slab = blockIdx.y;
    const int64_t r0 = R * slab / nslab, r1 = R * (slab + 1) / nslab;
    const int c = NHWC ? blockIdx.x * 32 + (threadIdx.x & 31) : blockIdx.x;
    if (c >= C) return;
    const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c);
    const float g = (weight ? __ldg(weight + c) : 1.f) * invstd;
    const float inv_count = 1.f / __ldg(total_count);
    const float mdy = sums[c] * inv_count;                      // mean of dy
    const float k = sums[C + c] * inv_count * invstd * invstd;  // mean of dy (x-mean), over the variance
    // dx = (dy - mdy - (x - mean) k) g
    if (NHWC) {
        for (int64_t r = r0 + (threadIdx.x >> 5); r < r1; r += 8) {
            const int64_t o = r * C + c;
            bn_st(dx + o, (bn_ld(dy + o) - mdy - (bn_ld(x + o) - mean) * k) * g);
        }
    } else {
        for (int64_t r = r0 + threadIdx.x; r < r1; r += BN_THREADS) {
            const int64_t n = r / HW;
            const int64_t o = (n * C + c) * HW + (r - n * HW);
            bn_st(dx + o, (bn_ld(dy + o) - mdy - (bn_ld(x + o) - mean) * k) * g);
        }
    }
}

// slabs of rows per channel tile: enough blocks to fill the machine, never more than BN_MAX_SLABS
static int bn_slabs(const BnShape& sh) {
    const int64_t R = sh.N * sh.HW;
    const int ctiles = sh.nhwc ? (sh.C + 31) / 32 : sh.C;
    int64_t want = (148 * 4 + ctiles - 1) / ctiles;
    const int64_t min_rows = sh.nhwc ? 64 : 1024;  // rows per slab worth a block
    if (want > (R + min_rows - 1) / min_rows) want = (R + min_rows - 1) / min_rows;
    if (want < 1) want = 1;
    if (want > BN_MAX_SLABS) want = BN_MAX_SLABS;
    return (int)want;
}
static dim3 bn_grid(const BnShape& sh, int nslab) { return dim3(sh.nhwc ? (sh.C + 31) / 32 : sh.C, nslab); }

}  // namespace pp

using namespace pp;

extern "C" {

int64_t pp_bn_workspace(int C) {
    // partial[BN_MAX_SLABS][2][C] floats + one ticket (kept zero between launches), 16-byte aligned
    return ((int64_t)BN_MAX_SLABS * 2 * C) * (int64_t)sizeof(float) + 16;
}

static int bn_check(const char* what, int64_t N, int C, int HW, int layout, int dtype) {
    PP_REQUIRE(N > 0 && C > 0 && HW > 0 && N * HW < (1ll << 40), "%s: bad shape N=%lld C=%d HW=%d", what, (long long)N, C, HW);
    PP_REQUIRE(layout == PP_LAYOUT_NCHW || layout == PP_LAYOUT_NHWC, "%s: bad layout %d", what, layout);
    PP_REQUIRE(dtype == PP_DTYPE_F32 || dtype == PP_DTYPE_BF16, "%s: bad dtype %d", what, dtype);
    PP_REQUIRE(C <= 65535 * 32, "%s: too many channels", what);
    return PP_OK;
}

int pp_bn_stats(const void* x, int64_t N, int C, int HW, int layout, int dtype, void* workspace, float* stats, void* stream) {
    int rc = bn_check("pp_bn_stats", N, C, HW, layout, dtype);
    if (rc) return rc;
    PP_REQUIRE(x && workspace && stats, "pp_bn_stats: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BnShape sh{N, C, HW, layout == PP_LAYOUT_NHWC};
    const int nslab = bn_slabs(sh);
    float* partial = (float*)workspace;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + (int64_t)BN_MAX_SLABS * 2 * C);
    const dim3 grid = bn_grid(sh, nslab);
    if (dtype == PP_DTYPE_F32) {
        if (sh.nhwc) PP_LAUNCH("bn stats", st, (bn_stats_kernel<float, true><<<grid, BN_THREADS, 0, st>>>((const float*)x, sh, nslab, partial, ticket, stats)));
        else PP_LAUNCH("bn stats", st, (bn_stats_kernel<float, false><<<grid, BN_THREADS, 0, st>>>((const float*)x, sh, nslab, partial, ticket, stats)));
    } else {
        if (sh.nhwc) PP_LAUNCH("bn stats", st, (bn_stats_kernel<__nv_bfloat16, true><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)x, sh, nslab, partial, ticket, stats)));
        else PP_LAUNCH("bn stats", st, (bn_stats_kernel<__nv_bfloat16, false><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)x, sh, nslab, partial, ticket, stats)));
    }
    return check_launch("bn_stats_kernel");
}

int pp_bn_apply(const void* x, void* y, int64_t N, int C, int HW, int layout, int dtype, const float* stats, int nranks, const float* weight,
                const float* bias, float* running_mean, float* running_var, double eps, double momentum, float* save_mean,
                float* save_invstd, void* stream) {
    int rc = bn_check("pp_bn_apply", N, C, HW, layout, dtype);
    if (rc) return rc;
    PP_REQUIRE(x && y && stats && save_mean && save_invstd && nranks >= 1, "pp_bn_apply: null pointer / bad rank count");
    PP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "pp_bn_apply: running_mean and running_var go together");
    cudaStream_t st = (cudaStream_t)stream;
    BnShape sh{N, C, HW, layout == PP_LAYOUT_NHWC};
    const int nslab = bn_slabs(sh);
    BnApplyArgs a{stats, weight, bias, nranks, running_mean, running_var, save_mean, save_invstd, (float)eps, (float)momentum};
    const dim3 grid = bn_grid(sh, nslab);
    if (dtype == PP_DTYPE_F32) {
        if (sh.nhwc) PP_LAUNCH("bn apply", st, (bn_apply_kernel<float, true><<<grid, BN_THREADS, 0, st>>>((const float*)x, (float*)y, sh, nslab, a)));
        else PP_LAUNCH("bn apply", st, (bn_apply_kernel<float, false><<<grid, BN_THREADS, 0, st>>>((const float*)x, (float*)y, sh, nslab, a)));
    } else {
        if (sh.nhwc) PP_LAUNCH("bn apply", st, (bn_apply_kernel<__nv_bfloat16, true><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, sh, nslab, a)));
        else PP_LAUNCH("bn apply", st, (bn_apply_kernel<__nv_bfloat16, false><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, sh, nslab, a)));
    }
    return check_launch("bn_apply_kernel");
}

int pp_bn_bwd_stats(const void* dy, const void* x, int64_t N, int C, int HW, int layout, int dtype, const float* save_mean,
                    const float* save_invstd, void* workspace, float* sums, float* grad_weight, float* grad_bias, void* stream) {
    int rc = bn_check("pp_bn_bwd_stats", N, C, HW, layout, dtype);
    if (rc) return rc;
    PP_REQUIRE(dy && x && save_mean && save_invstd && workspace && sums, "pp_bn_bwd_stats: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BnShape sh{N, C, HW, layout == PP_LAYOUT_NHWC};
    const int nslab = bn_slabs(sh);
    float* partial = (float*)workspace;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + (int64_t)BN_MAX_SLABS * 2 * C);
    const dim3 grid = bn_grid(sh, nslab);
    if (dtype == PP_DTYPE_F32) {
        if (sh.nhwc) PP_LAUNCH("bn bwd stats", st, (bn_bwd_stats_kernel<float, true><<<grid, BN_THREADS, 0, st>>>((const float*)dy, (const float*)x, sh, nslab, save_mean, save_invstd, partial, ticket, sums, grad_weight, grad_bias)));
        else PP_LAUNCH("bn bwd stats", st, (bn_bwd_stats_kernel<float, false><<<grid, BN_THREADS, 0, st>>>((const float*)dy, (const float*)x, sh, nslab, save_mean, save_invstd, partial, ticket, sums, grad_weight, grad_bias)));
    } else {
        if (sh.nhwc) PP_LAUNCH("bn bwd stats", st, (bn_bwd_stats_kernel<__nv_bfloat16, true><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, sh, nslab, save_mean, save_invstd, partial, ticket, sums, grad_weight, grad_bias)));
        else PP_LAUNCH("bn bwd stats", st, (bn_bwd_stats_kernel<__nv_bfloat16, false><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, sh, nslab, save_mean, save_invstd, partial, ticket, sums, grad_weight, grad_bias)));
    }
    return check_launch("bn_bwd_stats_kernel");
}

int pp_bn_bwd_apply(const void* dy, const void* x, void* dx, int64_t N, int C, int HW, int layout, int dtype, const float* save_mean,
                    const float* save_invstd, const float* weight, const float* sums, const float* total_count, void* stream) {
    int rc = bn_check("pp_bn_bwd_apply", N, C, HW, layout, dtype);
    if (rc) return rc;
    PP_REQUIRE(dy && x && dx && save_mean && save_invstd && sums && total_count, "pp_bn_bwd_apply: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BnShape sh{N, C, HW, layout == PP_LAYOUT_NHWC};
    const int nslab = bn_slabs(sh);
    const dim3 grid = bn_grid(sh, nslab);
    if (dtype == PP_DTYPE_F32) {
        if (sh.nhwc) PP_LAUNCH("bn bwd apply", st, (bn_bwd_apply_kernel<float, true><<<grid, BN_THREADS, 0, st>>>((const float*)dy, (const float*)x, (float*)dx, sh, nslab, save_mean, save_invstd, weight, sums, total_count)));
        else PP_LAUNCH("bn bwd apply", st, (bn_bwd_apply_kernel<float, false><<<grid, BN_THREADS, 0, st>>>((const float*)dy, (const float*)x, (float*)dx, sh, nslab, save_mean, save_invstd, weight, sums, total_count)));
    } else {
        if (sh.nhwc) PP_LAUNCH("bn bwd apply", st, (bn_bwd_apply_kernel<__nv_bfloat16, true><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (__nv_bfloat16*)dx, sh, nslab, save_mean, save_invstd, weight, sums, total_count)));
        else PP_LAUNCH("bn bwd apply", st, (bn_bwd_apply_kernel<__nv_bfloat16, false><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (__nv_bfloat16*)dx, sh, nslab, save_mean, save_invstd, weight, sums, total_count)));
    }
    return check_launch("bn_bwd_apply_kernel");
}

}  // extern "C"
