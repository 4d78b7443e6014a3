// pp_optim.cu — SURVEY §8(f) rank 1: the two per-parameter loops that follow the pixel path in
// every training step, as multi-tensor kernels (one launch over all parameters instead of
// ~330 / ~1500 tiny launches and one host sync per parameter):
//   EMA momentum update of the key branch      contrast/models/PixPro.py:322-337
//   LARS adaptive scaling + the wrapped SGD     contrast/lars.py:109-152, torch.optim.SGD
//
// A parameter set is described by a device table of MtTensor entries plus a chunk map
// (chunk -> tensor, chunk index within the tensor); the host side builds both once per model.
// Arithmetic follows the reference op by op (the library is compiled with -fmad=false, so only
// the explicit fmaf()s below fuse — where ATen's add(alpha=...) kernels fuse):
//   EMA : k = RN(RN(k*m) + RN(q*(1-m)))
//   LARS: g' = fma(wd, p, g)                      p.grad.add(p, alpha=wd)         lars.py:120
//         a  = trust*|p| / (|g'| + eps) if |p|>0 and |g'|>0 else 1  (fp32 ops)     lars.py:125-133
//         g''= g' * a                                                              lars.py:136
//   SGD : buf = g'' (first step) | fma(1-damp, g'', RN(buf*mom));  p = fma(-lr, buf, p)
// The two norms are deterministic (fixed chunking and reduction order) but not torch's order:
// they agree to fp32 rounding of a sum, not bitwise.
#include <math.h>

#include "pp_common.cuh"

namespace pp {

constexpr int MT_CHUNK = 8192;   // elements per chunk (block of 256 threads x 8 float4)
constexpr int MT_THREADS = 256;

struct MtChunk {
    int tensor;  // index into the tensor table
    int idx;     // chunk index within the tensor
};

__global__ void __launch_bounds__(MT_THREADS) ema_kernel(const PpMtTensor* __restrict__ tab, const MtChunk* __restrict__ map, float m,
                                                        float om) {
    const MtChunk ch = map[blockIdx.x];
    const PpMtTensor t = tab[ch.tensor];
    const float* q = reinterpret_cast<const float*>(t.a);
    float* k = reinterpret_cast<float*>(t.b);
    const int64_t base = (int64_t)ch.idx * MT_CHUNK;
    const int64_t end = min(base + MT_CHUNK, t.numel);
    if ((((uintptr_t)q | (uintptr_t)k) & 15) == 0 && (base & 3) == 0) {
        for (int64_t i = base + 4 * threadIdx.x; i < end; i += 4 * MT_THREADS) {
            if (i + 3 < end) {
                const float4 qv = *reinterpret_cast<const float4*>(q + i);
                float4 kv = *reinterpret_cast<float4*>(k + i);
                kv.x = add(mul(kv.x, m), mul(qv.x, om)); kv.y = add(mul(kv.y, m), mul(qv.y, om));
                kv.z = add(mul(kv.z, m), mul(qv.z, om)); kv.w = add(mul(kv.w, m), mul(qv.w, om));
                *reinterpret_cast<float4*>(k + i) = kv;
            } else {
                for (int64_t j = i; j < end; j++) k[j] = add(mul(k[j], m), mul(q[j], om));
            }
        }
    } else {
        for (int64_t i = base + threadIdx.x; i < end; i += MT_THREADS) k[i] = add(mul(k[i], m), mul(q[i], om));
    }
}

// per chunk: sum p^2 and sum (fma(wd,p,g))^2, in double (one deterministic tree per block)
__global__ void __launch_bounds__(MT_THREADS) lars_norm_partial_kernel(const PpMtTensor* __restrict__ tab, const MtChunk* __restrict__ map,
                                                                      double* __restrict__ partial /*[nchunks][2]*/) {
    const MtChunk ch = map[blockIdx.x];
    const PpMtTensor t = tab[ch.tensor];
    const float* p = reinterpret_cast<const float*>(t.a);
    const float* g = reinterpret_cast<const float*>(t.b);
    const int64_t base = (int64_t)ch.idx * MT_CHUNK;
    const int64_t end = min(base + MT_CHUNK, t.numel);
    const float wd = t.s0;
    double sp = 0.0, sg = 0.0;
    for (int64_t i = base + threadIdx.x; i < end; i += MT_THREADS) {
        const float pv = p[i];
        const float gv = wd > 0.0f ? fma_(wd, pv, g[i]) : g[i];
        sp += (double)pv * pv;
        sg += (double)gv * gv;
    }
    __shared__ double sh[2][MT_THREADS];
    sh[0][threadIdx.x] = sp;
    sh[1][threadIdx.x] = sg;
    __syncthreads();
    for (int s = MT_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * (int64_t)blockIdx.x] = sh[0][0];
        partial[2 * (int64_t)blockIdx.x + 1] = sh[1][0];
    }
}

// per tensor: norms from its chunks' partials (chunks of one tensor are contiguous in the map), then the
// adaptive rate with the reference's fp32 operation sequence
__global__ void lars_rate_kernel(const PpMtTensor* __restrict__ tab, int ntensors, const int* __restrict__ first_chunk,
                                 const double* __restrict__ partial, float trust, float eps, float* __restrict__ rate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntensors) return;
    float a = 1.0f;
    if (tab[t].flags & PP_MT_LARS) {
        double sp = 0.0, sg = 0.0;
        for (int c = first_chunk[t]; c < first_chunk[t + 1]; c++) { sp += partial[2 * (int64_t)c]; sg += partial[2 * (int64_t)c + 1]; }
        const float pn = (float)sqrt(sp), gn = (float)sqrt(sg);
        if (pn > 0.0f && gn > 0.0f) a = __fdiv_rn(mul(trust, pn), add(gn, eps));  // lars.py:133
    }
    rate[t] = a;
}

__global__ void __launch_bounds__(MT_THREADS) lars_sgd_kernel(const PpMtTensor* __restrict__ tab, const MtChunk* __restrict__ map,
                                                             const float* __restrict__ rate) {
    const MtChunk ch = map[blockIdx.x];
    const PpMtTensor t = tab[ch.tensor];
    float* p = reinterpret_cast<float*>(t.a);
    const float* g = reinterpret_cast<const float*>(t.b);
    float* buf = reinterpret_cast<float*>(t.c);
    const int64_t base = (int64_t)ch.idx * MT_CHUNK;
    const int64_t end = min(base + MT_CHUNK, t.numel);
    const float wd = t.s0, nlr = -t.s1, mom = t.s2, omd = 1.0f - t.s3;
    const bool lars = (t.flags & PP_MT_LARS) != 0, first = (t.flags & PP_MT_FIRST_STEP) != 0, has_mom = mom != 0.0f;
    const float a = rate[ch.tensor];
    for (int64_t i = base + threadIdx.x; i < end; i += MT_THREADS) {
        const float pv = p[i];
        float gv = wd > 0.0f ? fma_(wd, pv, g[i]) : g[i];
        if (lars) gv = mul(gv, a);
        if (has_mom) {
            const float b = first ? gv : fma_(omd, gv, mul(buf[i], mom));  // buf.mul_(momentum).add_(grad, alpha=1-dampening)
            buf[i] = b;
            gv = b;
        }
        p[i] = fma_(nlr, gv, pv);  // p.add_(grad, alpha=-lr)
    }
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_mt_chunk_elems(void) { return MT_CHUNK; }

int pp_ema_update(const PpMtTensor* table_dev, const int* chunk_map_dev, int nchunks, double momentum, double one_minus_momentum,
                  void* stream) {
    PP_REQUIRE(nchunks >= 0, "pp_ema_update: bad chunk count");
    if (nchunks == 0) return PP_OK;
    PP_REQUIRE(table_dev && chunk_map_dev, "pp_ema_update: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    PP_LAUNCH("ema_update", st,
              ema_kernel<<<nchunks, MT_THREADS, 0, st>>>(table_dev, reinterpret_cast<const MtChunk*>(chunk_map_dev), (float)momentum,
                                                         (float)one_minus_momentum));
    return check_launch("ema_kernel");
}

int64_t pp_lars_workspace(int ntensors, int nchunks) { return (int64_t)nchunks * 2 * sizeof(double) + (int64_t)ntensors * sizeof(float); }

int pp_lars_sgd_step(const PpMtTensor* table_dev, int ntensors, const int* chunk_map_dev, const int* first_chunk_dev, int nchunks,
                     double trust_coef, double eps, void* workspace, void* stream) {
    PP_REQUIRE(ntensors >= 0 && nchunks >= 0, "pp_lars_sgd_step: bad counts");
    if (ntensors == 0 || nchunks == 0) return PP_OK;
    PP_REQUIRE(table_dev && chunk_map_dev && first_chunk_dev && workspace, "pp_lars_sgd_step: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double* partial = reinterpret_cast<double*>(workspace);
    float* rate = reinterpret_cast<float*>(partial + 2 * (int64_t)nchunks);
    const MtChunk* map = reinterpret_cast<const MtChunk*>(chunk_map_dev);
    PP_LAUNCH("lars_norms", st, lars_norm_partial_kernel<<<nchunks, MT_THREADS, 0, st>>>(table_dev, map, partial));
    int rc = check_launch("lars_norm_partial_kernel");
    if (rc) return rc;
    PP_LAUNCH("lars_rates", st,
              lars_rate_kernel<<<(ntensors + 127) / 128, 128, 0, st>>>(table_dev, ntensors, first_chunk_dev, partial, (float)trust_coef,
                                                                        (float)eps, rate));
    rc = check_launch("lars_rate_kernel");
    if (rc) return rc;
    PP_LAUNCH("lars_sgd", st, lars_sgd_kernel<<<nchunks, MT_THREADS, 0, st>>>(table_dev, map, rate));
    return check_launch("lars_sgd_kernel");
}

}  // extern "C"
