// pp_loss.cu — flow-guided correspondence, positive mask and masked cosine regression loss
// (sm_100a), forward and backward in one pass.
//
// Reference functions restated (paths relative to the reference repo):
//   add_optical_flow   contrast/models/PixPro.py:46-89
//   regression_loss    contrast/models/PixPro.py:92-247
//
// Formulation.  The reference materialises logit = qᵀk [B,P,P], multiplies by the positive
// mask and reduces.  Here   loss_b = Σ_i q_i · m_i / den_b   with   m_i = Σ_j pos_ij k_j ,
// and  d loss / d q_i = -2/(B den_b) m_i , so ONE masked contraction M = K·posᵀ gives both
// the loss and its gradient; the [B,P,P] logits and mask never touch HBM.  The mask is
// recomputed per tile from the 5·P per-sample centre/mask scalars (bit-exact arithmetic,
// see pp_common.cuh).
#include <math.h>

#include "pp_common.cuh"
#include "pp_small.cuh"
#include "pp_tc.cuh"
#include "pp_tc2.cuh"
#include "pp_warp.cuh"

namespace pp {

// ---- a7: warp one grid point through the flow ----------------------------------------------
// (WarpArgs: pp_warp.cuh)
__device__ __forceinline__ void warp_point(const float* __restrict__ flow, const uint8_t* __restrict__ mask, const WarpArgs& a,
                                           float xg, float yg, float& ox, float& oy, bool& mg) {
    // PixPro.py:61-62   2 * (x / (W_orig-1)) - 1
    float gx = sub(mul(2.0f, a.dwo(xg)), 1.0f);
    float gy = sub(mul(2.0f, a.dho(yg)), 1.0f);
    int64_t HW = (int64_t)a.Hin * a.Win;
    Taps t = make_taps(gx, gy, a.Win, a.Hin, a.half_w, a.half_h);
    float vx[4], vy[4];
    const bool in[4] = {t.inx0 && t.iny0, t.inx1 && t.iny0, t.inx0 && t.iny1, t.inx1 && t.iny1};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int64_t o = (int64_t)(t.y0 + (c >> 1)) * a.Win + (t.x0 + (c & 1));
        vx[c] = in[c] ? __ldg(flow + o) : 0.0f;
        vy[c] = in[c] ? __ldg(flow + HW + o) : 0.0f;
    }
    float fgx = combine(t, vx[0], vx[1], vx[2], vx[3]);  // PixPro.py:64
    float fgy = combine(t, vy[0], vy[1], vy[2], vy[3]);
    mg = true;
    if (mask) {  // PixPro.py:65-70 nearest lookup (nearbyint, zeros padding)
        float ix = mul(add(gx, 1.0f), a.half_w), iy = mul(add(gy, 1.0f), a.half_h);
        float xr = rintf(ix), yr = rintf(iy);
        bool inb = (xr > -1.0f) && (xr < (float)a.Win) && (yr > -1.0f) && (yr < (float)a.Hin);
        mg = inb ? (mask[(int64_t)yr * a.Win + (int64_t)xr] != 0) : false;
    }
    if (a.diff) {  // PixPro.py:76-80
        ox = a.drw(add(mul(xg, a.rw), fgx));
        oy = a.drh(add(mul(yg, a.rh), fgy));
    } else {  // PixPro.py:82-83
        ox = add(xg, fgx);
        oy = add(yg, fgy);
    }
}

__global__ void __launch_bounds__(128) add_flow_kernel(const float* __restrict__ flow, const uint8_t* __restrict__ mask,
                                                        const float* __restrict__ xg, const float* __restrict__ yg,
                                                        int64_t total, int P, WarpArgs a, float* __restrict__ out_x,
                                                        float* __restrict__ out_y, uint8_t* __restrict__ mask_grid) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t b = i / P;
    int64_t HW = (int64_t)a.Hin * a.Win;
    float ox, oy;
    bool mg;
    warp_point(flow + b * 2 * HW, mask ? mask + b * HW : nullptr, a, xg[i], yg[i], ox, oy, mg);
    out_x[i] = ox;
    out_y[i] = oy;
    if (mask_grid) mask_grid[i] = mg ? 1 : 0;
}

// ---- a8 stage 1: centres, flow warp, positive count ----------------------------------------
// Workspace (floats): [0] cqx[B,P] [1] cqy [2] ckx [3] cky [4] mg (0/1) ; then md[B], den[B],
// partial[B,ntile].
struct LossWs {
    float *cqx, *cqy, *ckx, *cky, *mg, *md, *den, *partial;
    int* rowcnt;     // [B,P] positives per query cell        (large-grid path)
    uint8_t* posb;   // [B,P,P] positive matrix as bytes       (large-grid path)
};
static int loss_ntile(int P) { return (P + 6) / 7; }
static LossWs carve_ws(void* ws, int64_t B, int P) {
    LossWs w;
    float* f = (float*)ws;
    w.cqx = f; f += B * P;
    w.cqy = f; f += B * P;
    w.ckx = f; f += B * P;
    w.cky = f; f += B * P;
    w.mg = f; f += B * P;
    w.md = f; f += B;
    w.den = f; f += B;
    w.partial = f; f += B * loss_ntile(P);
    w.rowcnt = reinterpret_cast<int*>(f); f += B * P;
    w.posb = reinterpret_cast<uint8_t*>(f);
    return w;
}

__device__ __forceinline__ bool pair_pos(float qx, float qy, float kx, float ky, float md, float pr) {
    float dx = sub(qx, kx), dy = sub(qy, ky);                       // PixPro.py:217
    float d = __fdiv_rn(__fsqrt_rn(add(mul(dx, dx), mul(dy, dy))), md);  // :217-218
    return d < pr;                                                  // :219
}

struct PrepArgs {
    const float *coord_q, *coord_k, *flow;
    const uint8_t* mask;
    const float* pre;   // [3,B,P] warped q centres (x, y) and mask bit from pp_sparse_corr, or NULL
    int G, P;
    float wo, ho;       // W_orig-1, H_orig-1
    ScalarDiv dG;       // / G
    float pr;
    WarpArgs warp;
    LossWs ws;
    float *pos_num, *pos_mean, *centres;
    uint8_t* pos_mask;
    int64_t B;
};

__global__ void __launch_bounds__(256) loss_prep_kernel(PrepArgs a) {
    extern __shared__ float sm[];
    const int P = a.P, G = a.G;
    float* qx = sm;
    float* qy = qx + P;
    float* kx = qy + P;
    float* ky = kx + P;
    float* mgs = ky + P;
    __shared__ float s_md;
    __shared__ int s_cnt[8];
    const int64_t b = blockIdx.x;
    const float* cq = a.coord_q + b * 10;
    const float* ck = a.coord_k + b * 10;
    // PixPro.py:140-143 bin sizes
    float qbw = a.dG(sub(cq[2], cq[0])), qbh = a.dG(sub(cq[3], cq[1]));
    float kbw = a.dG(sub(ck[2], ck[0])), kbh = a.dG(sub(ck[3], ck[1]));
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        int x = p % G, y = p / G;
        // :168-175 / :192-199   ((i+0.5)*bin + start) * (size-1)
        float fx = add((float)x, 0.5f), fy = add((float)y, 0.5f);
        float vqx = mul(add(mul(fx, qbw), cq[0]), a.wo);
        float vqy = mul(add(mul(fy, qbh), cq[1]), a.ho);
        float vkx = mul(add(mul(fx, kbw), ck[0]), a.wo);
        float vky = mul(add(mul(fy, kbh), ck[1]), a.ho);
        bool mg = true;
        if (a.pre) {  // :200, already evaluated by the sparse correspondence kernel
            const int64_t BP = a.B * P;
            vqx = a.pre[b * P + p]; vqy = a.pre[BP + b * P + p]; mg = a.pre[2 * BP + b * P + p] != 0.0f;
        } else if (a.flow) {  // :200
            int64_t HW = (int64_t)a.warp.Hin * a.warp.Win;
            float ox, oy;
            warp_point(a.flow + b * 2 * HW, a.mask ? a.mask + b * HW : nullptr, a.warp, vqx, vqy, ox, oy, mg);
            vqx = ox;
            vqy = oy;
        }
        qx[p] = vqx; qy[p] = vqy; kx[p] = vkx; ky[p] = vky; mgs[p] = mg ? 1.0f : 0.0f;
        a.ws.cqx[b * P + p] = vqx; a.ws.cqy[b * P + p] = vqy;
        a.ws.ckx[b * P + p] = vkx; a.ws.cky[b * P + p] = vky;
        a.ws.mg[b * P + p] = mg ? 1.0f : 0.0f;
        if (a.centres) {
            int64_t BP = a.B * P;
            a.centres[0 * BP + b * P + p] = vqx; a.centres[1 * BP + b * P + p] = vqy;
            a.centres[2 * BP + b * P + p] = vkx; a.centres[3 * BP + b * P + p] = vky;
        }
    }
    if (threadIdx.x == 0) {  // :155-157
        float qdw = mul(qbw, a.wo), qdh = mul(qbh, a.ho), kdw = mul(kbw, a.wo), kdh = mul(kbh, a.ho);
        float qd = __fsqrt_rn(add(mul(qdw, qdw), mul(qdh, qdh)));
        float kd = __fsqrt_rn(add(mul(kdw, kdw), mul(kdh, kdh)));
        s_md = fmaxf(qd, kd);
    }
    __syncthreads();
    const float md = s_md;
    int cnt = 0;
    for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
        int i = e / P, j = e - i * P;
        bool pos = pair_pos(qx[i], qy[i], kx[j], ky[j], md, a.pr) && (mgs[i] != 0.0f);  // :219-222
        cnt += pos;
        if (a.pos_mask) a.pos_mask[(b * P + i) * (int64_t)P + j] = pos ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_cnt[w];
        float fc = (float)tot;
        a.ws.md[b] = md;
        a.ws.den[b] = add(fc, 1e-6f);  // :241 fp32 denominator
        if (a.pos_num) a.pos_num[b] = fc;
        if (a.pos_mean) a.pos_mean[b] = __fdiv_rn(fc, (float)(P * P));
    }
}

// ---- a8 stage 2: masked contraction M = K posᵀ, loss partials and dq -----------------------
// grid (ntile, B); block 256.  Tile = TI query cells.  k is staged through shared memory in
// [256 channels][JT] chunks; thread c owns channel c and TI accumulators.
constexpr int TI = 7;
constexpr int JT = 32;

struct MainArgs {
    const float *q, *k;
    float* dq;
    LossWs ws;
    int C, P, ntile;
    float pr, scale;  // scale = -2/B
};

__global__ void __launch_bounds__(256) loss_main_kernel(MainArgs a) {
    __shared__ float ks[256][JT + 1];
    __shared__ uint8_t pos[TI][1024];  // P <= 1024
    __shared__ float red[8];
    const int P = a.P, C = a.C;
    const int64_t b = blockIdx.y;
    const int i0 = blockIdx.x * TI;
    const int ni = min(TI, P - i0);
    const float md = a.ws.md[b];
    const float* kx = a.ws.ckx + b * P;
    const float* ky = a.ws.cky + b * P;
    for (int e = threadIdx.x; e < TI * P; e += blockDim.x) {
        int t = e / P, j = e - t * P;
        bool p = false;
        if (t < ni) {
            int i = i0 + t;
            p = pair_pos(a.ws.cqx[b * P + i], a.ws.cqy[b * P + i], kx[j], ky[j], md, a.pr) && (a.ws.mg[b * P + i] != 0.0f);
        }
        pos[t][j] = p;
    }
    const float inv_den = a.scale / a.ws.den[b];  // -2/(B den)
    float lsum = 0.0f;
    for (int c0 = 0; c0 < C; c0 += 256) {
        const int c = c0 + threadIdx.x;
        float acc[TI];
#pragma unroll
        for (int t = 0; t < TI; t++) acc[t] = 0.0f;
        for (int j0 = 0; j0 < P; j0 += JT) {
            __syncthreads();  // pos ready / previous chunk consumed
            const int nj = min(JT, P - j0);
            for (int e = threadIdx.x; e < 256 * JT; e += blockDim.x) {
                int cc = e / JT, jj = e - cc * JT;
                float v = 0.0f;
                if (c0 + cc < C && jj < nj) v = __ldg(a.k + (b * C + c0 + cc) * (int64_t)P + j0 + jj);
                ks[cc][jj] = v;
            }
            __syncthreads();
            for (int jj = 0; jj < nj; jj++) {
                float kv = ks[threadIdx.x][jj];
#pragma unroll
                for (int t = 0; t < TI; t++) acc[t] += pos[t][j0 + jj] ? kv : 0.0f;
            }
        }
        if (c < C) {
            const float* qrow = a.q + (b * C + c) * (int64_t)P + i0;
            float* drow = a.dq ? a.dq + (b * C + c) * (int64_t)P + i0 : nullptr;
#pragma unroll
            for (int t = 0; t < TI; t++) {
                if (t < ni) {
                    lsum = fmaf(__ldg(qrow + t), acc[t], lsum);
                    if (drow) drow[t] = acc[t] * inv_den;
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < 8; w++) s += red[w];
        a.ws.partial[b * a.ntile + blockIdx.x] = s;
    }
}

// ---- a8 stage 3: deterministic final reduction: loss = -2 mean_b( Σ partial_b / den_b ) ----
struct FinalArgs {
    LossWs ws[2];
    float* loss[2];
};
__global__ void __launch_bounds__(256) loss_final_kernel(FinalArgs fa, int64_t B, int ntile) {
    const LossWs& ws = blockIdx.x ? fa.ws[1] : fa.ws[0];
    float* loss = blockIdx.x ? fa.loss[1] : fa.loss[0];
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
        float sb = 0.0f;
        for (int t = 0; t < ntile; t++) sb += ws.partial[b * ntile + t];
        s += (double)(sb / ws.den[b]);  // PixPro.py:241
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(-2.0 * red[0] / (double)B);  // :247
}

// ---- a8 for large grids (P > 64: 14x14, 28x28): tensor-core contraction ------------------------
// centres/warp per sample -> positive matrix as bytes + exact row counts (parallel over rows) ->
// denominators -> M = K posᵀ on the tcgen05 kernel (0/1 operand: exact in TF32) scaled to dq in
// the epilogue -> per-sample Σ q∘dq.  Same arithmetic for every boolean as the small path.
__global__ void __launch_bounds__(256) loss_centres_kernel(PrepArgs a) {
    const int P = a.P, G = a.G;
    const int64_t b = blockIdx.x;
    const float* cq = a.coord_q + b * 10;
    const float* ck = a.coord_k + b * 10;
    float qbw = a.dG(sub(cq[2], cq[0])), qbh = a.dG(sub(cq[3], cq[1]));
    float kbw = a.dG(sub(ck[2], ck[0])), kbh = a.dG(sub(ck[3], ck[1]));
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        int x = p % G, y = p / G;
        float fx = add((float)x, 0.5f), fy = add((float)y, 0.5f);
        float vqx = mul(add(mul(fx, qbw), cq[0]), a.wo), vqy = mul(add(mul(fy, qbh), cq[1]), a.ho);
        float vkx = mul(add(mul(fx, kbw), ck[0]), a.wo), vky = mul(add(mul(fy, kbh), ck[1]), a.ho);
        bool mg = true;
        if (a.pre) {
            const int64_t BP = a.B * P;
            vqx = a.pre[b * P + p]; vqy = a.pre[BP + b * P + p]; mg = a.pre[2 * BP + b * P + p] != 0.0f;
        } else if (a.flow) {
            int64_t HW = (int64_t)a.warp.Hin * a.warp.Win;
            float ox, oy;
            warp_point(a.flow + b * 2 * HW, a.mask ? a.mask + b * HW : nullptr, a.warp, vqx, vqy, ox, oy, mg);
            vqx = ox;
            vqy = oy;
        }
        a.ws.cqx[b * P + p] = vqx; a.ws.cqy[b * P + p] = vqy;
        a.ws.ckx[b * P + p] = vkx; a.ws.cky[b * P + p] = vky;
        a.ws.mg[b * P + p] = mg ? 1.0f : 0.0f;
        if (a.centres) {
            int64_t BP = a.B * P;
            a.centres[0 * BP + b * P + p] = vqx; a.centres[1 * BP + b * P + p] = vqy;
            a.centres[2 * BP + b * P + p] = vkx; a.centres[3 * BP + b * P + p] = vky;
        }
    }
    if (threadIdx.x == 0) {
        float qdw = mul(qbw, a.wo), qdh = mul(qbh, a.ho), kdw = mul(kbw, a.wo), kdh = mul(kbh, a.ho);
        float qd = __fsqrt_rn(add(mul(qdw, qdw), mul(qdh, qdh)));
        float kd = __fsqrt_rn(add(mul(kdw, kdw), mul(kdh, kdh)));
        a.ws.md[b] = fmaxf(qd, kd);
    }
}

// The positive test of one pair, d(s) = RN(RN(sqrt(s)) / md) < pr with s = RN(RN(dx^2) + RN(dy^2)) (PixPro.py:217-219), is a
// monotone function of s: both correctly rounded operations are non-decreasing.  So for a sample (md) there is ONE float s* with
// d(s) < pr  <=>  s < s*, and the P^2 pair tests of the large grids (614 656 per sample and direction at 28x28, each a square
// root and a division = ~20 of their ~30 instructions) become one comparison each — same bits, NaN and infinity included (both
// forms are false for them).  s* = the smallest float whose d is not below pr, found by stepping ulps from (pr md)^2, with the
// exact test itself as the judge; < 0 = not found within the step budget (the caller then tests every pair exactly).
__device__ __forceinline__ bool dist_below(float s, float md, float pr) { return __fdiv_rn(__fsqrt_rn(s), md) < pr; }
__device__ float pos_threshold(float md, float pr) {
    if (!(md > 0.0f) || !(pr > 0.0f) || !isfinite(md) || !isfinite(pr)) return -1.0f;
    float s = mul(mul(pr, md), mul(pr, md));
    if (!isfinite(s) || !(s > 1e-30f)) return -1.0f;
    int n = 0;
    for (; n < 64 && dist_below(s, md, pr); n++) s = __uint_as_float(__float_as_uint(s) + 1u);            // up to the first false
    if (n == 64) return -1.0f;
    for (n = 0; n < 64 && !dist_below(__uint_as_float(__float_as_uint(s) - 1u), md, pr); n++)              // down to the smallest false
        s = __uint_as_float(__float_as_uint(s) - 1u);
    return (n == 64 || !isfinite(s)) ? -1.0f : s;
}

// grid (ceil(P/8), B); warp w of the block owns query cell i = 8*blockIdx.x + w
// posf (optional): the same matrix as 0/1 floats — the B operand plane the TMA-fed contraction streams (exact in TF32)
__global__ void __launch_bounds__(256) loss_pos_kernel(LossWs ws, int P, float pr, uint8_t* posb, float* posf) {
    const int64_t b = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= P) return;
    const float qx = ws.cqx[b * P + i], qy = ws.cqy[b * P + i], md = ws.md[b];
    const bool mg = ws.mg[b * P + i] != 0.0f;
    const float* kx = ws.ckx + b * P;
    const float* ky = ws.cky + b * P;
    uint8_t* row = posb + (b * P + i) * (int64_t)P;
    int cnt = 0;
    const float sstar = mg ? pos_threshold(md, pr) : 0.0f;  // warp-uniform (a masked-out query cell has no positives: s < 0 never)
    if (sstar >= 0.0f && (P & 3) == 0) {  // 4 key cells per lane: one 16-byte and one 4-byte store per 4 pairs
        float* rowf = posf ? posf + (b * P + i) * (int64_t)P : nullptr;
        for (int j = 4 * lane; j < P; j += 128) {
            const float4 x4 = *reinterpret_cast<const float4*>(kx + j), y4 = *reinterpret_cast<const float4*>(ky + j);
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w};
            bool ps[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float dx = sub(qx, xs[u]), dy = sub(qy, ys[u]);
                ps[u] = add(mul(dx, dx), mul(dy, dy)) < sstar;
                cnt += ps[u];
            }
            *reinterpret_cast<uchar4*>(row + j) = make_uchar4(ps[0], ps[1], ps[2], ps[3]);
            if (rowf) *reinterpret_cast<float4*>(rowf + j) = make_float4(ps[0] ? 1.0f : 0.0f, ps[1] ? 1.0f : 0.0f, ps[2] ? 1.0f : 0.0f, ps[3] ? 1.0f : 0.0f);
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) ws.rowcnt[b * P + i] = cnt;
        return;
    }
    if (sstar >= 0.0f) {
        for (int j = lane; j < P; j += 32) {
            const float dx = sub(qx, kx[j]), dy = sub(qy, ky[j]);
            const bool pos = add(mul(dx, dx), mul(dy, dy)) < sstar;
            row[j] = pos ? 1 : 0;
            if (posf) posf[(b * P + i) * (int64_t)P + j] = pos ? 1.0f : 0.0f;
            cnt += pos;
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) ws.rowcnt[b * P + i] = cnt;
        return;
    }
    for (int j = lane; j < P; j += 32) {
        bool pos = mg && pair_pos(qx, qy, kx[j], ky[j], md, pr);
        row[j] = pos ? 1 : 0;
        if (posf) posf[(b * P + i) * (int64_t)P + j] = pos ? 1.0f : 0.0f;
        cnt += pos;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) ws.rowcnt[b * P + i] = cnt;
}

__global__ void __launch_bounds__(256) loss_cnt_kernel(LossWs ws, int P, float* pos_num, float* pos_mean) {
    __shared__ int red[256];
    const int64_t b = blockIdx.x;
    int c = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) c += ws.rowcnt[b * P + i];
    red[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float fc = (float)red[0];
        ws.den[b] = add(fc, 1e-6f);  // PixPro.py:241 fp32 denominator
        if (pos_num) pos_num[b] = fc;
        if (pos_mean) pos_mean[b] = __fdiv_rn(fc, (float)(P * P));
    }
}

struct TcLdPos {  // B operand: row = query cell i, k = key cell j; bytes -> 0/1 floats (exact in TF32)
    static constexpr bool kRowMajorK = true;
    const uint8_t* posb;
    int P;
    __device__ __forceinline__ float4 load4(int64_t b, int i, int j) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < P && j < P) {
            const uint8_t* q = posb + (b * P + i) * (int64_t)P + j;
            if ((P & 3) == 0) {  // j is a multiple of 4: one aligned 32-bit load instead of four byte loads
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(q));
                return make_float4((float)(w & 0xffu), (float)((w >> 8) & 0xffu), (float)((w >> 16) & 0xffu), (float)(w >> 24));
            }
            v.x = (float)q[0];
            if (j + 1 < P) v.y = (float)q[1];
            if (j + 2 < P) v.z = (float)q[2];
            if (j + 3 < P) v.w = (float)q[3];
        }
        return v;
    }
};
struct TcStDq {  // dq[b][c][i..i+15] = M * (-2 / (B den_b))
    static constexpr bool kAux = false;
    float* dq;
    const float* den;
    float scale;
    int C, P;
    __device__ __forceinline__ void store4(int64_t b, int c, int i, float4 v) const {
        const float s = scale / __ldg(den + b);
        st4_guard(dq + (b * C + c) * (int64_t)P + i, i, P, make_float4(v.x * s, v.y * s, v.z * s, v.w * s));
    }
    __device__ __forceinline__ void store16(int64_t b, int c, int i, const float v[16]) const {
        const float s = scale / __ldg(den + b);
        float* q = dq + (b * C + c) * (int64_t)P + i;
#pragma unroll
        for (int u = 0; u < 16; u++)
            if (i + u < P) q[u] = v[u] * s;
    }
};

// partial[b][t] = slice t of Σ q∘M = (Σ q∘dq) * den_b / scale: kDotSplit blocks per sample (one block per sample left
// 32-64 blocks on 148 SMs: 0.24 ms at 28x28), each a deterministic block reduction; loss_final adds the slices in order.
constexpr int kDotSplit = 8;
__global__ void __launch_bounds__(256) loss_dot_kernel(const float* __restrict__ q, const float* __restrict__ dq, LossWs ws,
                                                        int64_t CP, float scale) {
    __shared__ float red[256];
    const int64_t b = blockIdx.y;
    const int64_t per = (CP + kDotSplit - 1) / kDotSplit;
    const int64_t e0 = blockIdx.x * per, e1 = (e0 + per < CP) ? e0 + per : CP;
    const float* qb = q + b * CP;
    const float* db = dq + b * CP;
    float s = 0.0f;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) s = fmaf(__ldg(qb + e), __ldg(db + e), s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) ws.partial[b * kDotSplit + blockIdx.x] = red[0] * (ws.den[b] / scale);
}

// ---- a8 for small grids (P <= 64): ONE block per sample does stages 1 and 2 ------------------
// centres + flow warp + positive matrix (as 0/1 floats, transposed) in shared memory, then
// M = K·posᵀ with the thread-per-channel contraction of pp_small.cuh, the loss partial Σ q∘M and
// dq, all without leaving the SM.  Same arithmetic as loss_prep_kernel / loss_main_kernel.
struct SmallLossArgs {
    PrepArgs p;
    const float *q, *k;
    float* dq;
    int C;
    float scale;  // -2/B
};
// up to two independent calls (the two loss directions of PixPro.forward) in one launch:
// blocks [0,B) serve seg[0], blocks [B,2B) serve seg[1]
struct SmallLossArgs2 {
    SmallLossArgs seg[2];
    int B;
};

__host__ __device__ inline size_t small_loss_smem_bytes(int C, int P) {
    return ((size_t)2 * ((size_t)C * P + SLACK) + (size_t)PMAX * PS + 5 * PMAX + SM_THREADS + 16) * sizeof(float);
}

__global__ void __launch_bounds__(SM_THREADS) loss_small_kernel(SmallLossArgs2 sa2) {
    extern __shared__ __align__(16) float smem[];
    const SmallLossArgs& sa = ((int)blockIdx.x < sa2.B) ? sa2.seg[0] : sa2.seg[1];
    const PrepArgs& a = sa.p;
    const int C = sa.C, P = a.P, G = a.G, CP = C * P;
    float* ks = smem;                // k rows, later M
    float* qs = ks + CP + SLACK;     // q rows
    float* Pm = qs + CP + SLACK;     // Pm[j][i] = pos(i,j) as 0/1
    float* qx = Pm + PMAX * PS;
    float* qy = qx + PMAX;
    float* kx = qy + PMAX;
    float* ky = kx + PMAX;
    float* mgs = ky + PMAX;
    float* red = mgs + PMAX;       // SM_THREADS scratch
    float* sc = red + SM_THREADS;  // [0] md, [1] den
    const int64_t b = ((int)blockIdx.x < sa2.B) ? blockIdx.x : blockIdx.x - sa2.B;
    const int64_t off = b * (int64_t)C * P;
    stage_dense(ks, sa.k + off, CP);
    stage_dense(qs, sa.q + off, CP);
    for (int e = threadIdx.x; e < PMAX * PS; e += SM_THREADS) Pm[e] = 0.0f;
    const float* cq = a.coord_q + b * 10;
    const float* ck = a.coord_k + b * 10;
    float qbw = a.dG(sub(cq[2], cq[0])), qbh = a.dG(sub(cq[3], cq[1]));  // PixPro.py:140-143
    float kbw = a.dG(sub(ck[2], ck[0])), kbh = a.dG(sub(ck[3], ck[1]));
    if (threadIdx.x < P) {
        const int p = threadIdx.x, x = p % G, y = p / G;
        float fx = add((float)x, 0.5f), fy = add((float)y, 0.5f);        // :168-175 / :192-199
        float vqx = mul(add(mul(fx, qbw), cq[0]), a.wo), vqy = mul(add(mul(fy, qbh), cq[1]), a.ho);
        float vkx = mul(add(mul(fx, kbw), ck[0]), a.wo), vky = mul(add(mul(fy, kbh), ck[1]), a.ho);
        bool mg = true;
        if (a.pre) {  // :200, already evaluated by the sparse correspondence kernel
            const int64_t BP = a.B * P;
            vqx = a.pre[b * P + p]; vqy = a.pre[BP + b * P + p]; mg = a.pre[2 * BP + b * P + p] != 0.0f;
        } else if (a.flow) {  // :200
            int64_t HW = (int64_t)a.warp.Hin * a.warp.Win;
            float ox, oy;
            warp_point(a.flow + b * 2 * HW, a.mask ? a.mask + b * HW : nullptr, a.warp, vqx, vqy, ox, oy, mg);
            vqx = ox;
            vqy = oy;
        }
        qx[p] = vqx; qy[p] = vqy; kx[p] = vkx; ky[p] = vky; mgs[p] = mg ? 1.0f : 0.0f;
        if (a.centres) {
            int64_t BP = a.B * P;
            a.centres[0 * BP + b * P + p] = vqx; a.centres[1 * BP + b * P + p] = vqy;
            a.centres[2 * BP + b * P + p] = vkx; a.centres[3 * BP + b * P + p] = vky;
        }
    }
    if (threadIdx.x == 0) {  // :155-157
        float qdw = mul(qbw, a.wo), qdh = mul(qbh, a.ho), kdw = mul(kbw, a.wo), kdh = mul(kbh, a.ho);
        float qd = __fsqrt_rn(add(mul(qdw, qdw), mul(qdh, qdh)));
        float kd = __fsqrt_rn(add(mul(kdw, kdw), mul(kdh, kdh)));
        sc[0] = fmaxf(qd, kd);
    }
    __syncthreads();
    const float md = sc[0];
    float cnt = 0.0f;
    for (int e = threadIdx.x; e < P * P; e += SM_THREADS) {
        int i = e / P, j = e - i * P;
        bool pos = pair_pos(qx[i], qy[i], kx[j], ky[j], md, a.pr) && (mgs[i] != 0.0f);  // :219-222
        Pm[j * PS + i] = pos ? 1.0f : 0.0f;
        cnt += pos ? 1.0f : 0.0f;
        if (a.pos_mask) a.pos_mask[(b * P + i) * (int64_t)P + j] = pos ? 1 : 0;
    }
    red[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = 0.0f;
        for (int u = threadIdx.x; u < SM_THREADS; u += 32) t += red[u];  // small integers: exact in fp32
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            sc[1] = add(t, 1e-6f);  // :241 fp32 denominator
            if (a.pos_num) a.pos_num[b] = t;
            if (a.pos_mean) a.pos_mean[b] = __fdiv_rn(t, (float)(P * P));
            a.ws.md[b] = md;
            a.ws.den[b] = sc[1];
        }
    }
    __syncthreads();
    // M[c][i] = Σ_j k[c][j] pos(i,j)  (in place over the k rows)
    row_times_mat(ks, Pm, C, P, ks);
    __syncthreads();
    const float inv_den = sa.scale / sc[1];  // -2/(B den)
    float lsum = 0.0f;
    for (int e = threadIdx.x; e < CP; e += SM_THREADS) {
        float m = ks[e];
        lsum = fmaf(qs[e], m, lsum);
        if (sa.dq) sa.dq[off + e] = m * inv_den;
    }
    __syncthreads();
    red[threadIdx.x] = lsum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = 0.0f;
        for (int u = threadIdx.x; u < SM_THREADS; u += 32) t += red[u];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) a.ws.partial[b] = t;  // ntile = 1
    }
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_add_optical_flow(const float* flow, int64_t B, int Hin, int Win, const float* x_grid, const float* y_grid, int P,
                        int H_orig, int W_orig, const uint8_t* mask, int div_mode, float* out_x, float* out_y,
                        uint8_t* mask_grid, void* stream) {
    PP_REQUIRE(flow && x_grid && y_grid && out_x && out_y, "pp_add_optical_flow: null pointer");
    PP_REQUIRE(B >= 0 && Hin > 1 && Win > 1 && P > 0 && H_orig > 1 && W_orig > 1, "pp_add_optical_flow: bad shape");
    if (B == 0) return PP_OK;
    int64_t total = B * P;
    cudaStream_t st = (cudaStream_t)stream;
    PP_LAUNCH("add_flow", st,
              add_flow_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
                  flow, mask, x_grid, y_grid, total, P, make_warp_args(Hin, Win, H_orig, W_orig, div_mode), out_x, out_y,
                  mask_grid));
    return check_launch("add_flow_kernel");
}

// The TMA-fed contraction (pp_tc2.cuh) streams planes with 16-byte row strides: C % 4 == 0 and P % 4 == 0.
static bool loss_tc2(int C, int P) {
    static const int off = [] { const char* e = getenv("PIXPRO_B200_TC2"); return (e && e[0] == '0') ? 1 : 0; }();
    return !off && use_tensor_cores(P) && P > PMAX && C % 4 == 0 && P % 4 == 0;
}

int64_t pp_regression_loss_workspace(int64_t B, int C, int G) {
    int64_t P = (int64_t)G * G;
    int64_t bytes = (5 * B * P + 2 * B + B * loss_ntile((int)P)) * (int64_t)sizeof(float);
    if (P > PMAX) bytes += B * P * (int64_t)sizeof(int) + B * P * P;  // row counts + byte matrix of positives
    if (loss_tc2(C, (int)P)) bytes += 16 + B * P * P * (int64_t)sizeof(float);  // (16-byte aligned) 0/1 float plane
    return bytes;
}

struct LossCall {
    const float *q, *k, *coord_q, *coord_k, *flow;
    const uint8_t* mask;
    float *loss, *pos_num, *pos_mean, *dq;
    uint8_t* pos_mask;
    float* centres;
    void* workspace;
    const float* pre = nullptr;
};

static int regression_loss_impl(const LossCall* calls, int ncall, int64_t B, int C, int G, int Hin, int Win, int H_orig,
                                int W_orig, double pos_ratio, int div_mode, void* stream) {
    PP_REQUIRE(B > 0 && B <= 32767 && C > 0 && G > 0, "pp_regression_loss: bad shape B=%lld C=%d G=%d", (long long)B, C, G);
    PP_REQUIRE(G * G <= 1024, "pp_regression_loss: grid %dx%d exceeds 1024 cells", G, G);
    PP_REQUIRE(H_orig > 1 && W_orig > 1, "pp_regression_loss: bad original size %dx%d", H_orig, W_orig);
    cudaStream_t st = (cudaStream_t)stream;
    const int P = G * G;
    const float scale = (float)(-2.0 / (double)B);
    PrepArgs pa[2];
    for (int c = 0; c < ncall; c++) {
        const LossCall& L = calls[c];
        PP_REQUIRE(L.q && L.k && L.coord_q && L.coord_k && L.loss && L.workspace, "pp_regression_loss: null pointer");
        PP_REQUIRE(!L.flow || (Hin > 1 && Win > 1), "pp_regression_loss: bad flow size %dx%d", Hin, Win);
        PP_REQUIRE(!L.mask || L.flow, "pp_regression_loss: mask without flow");
        PrepArgs& a = pa[c];
        PP_REQUIRE(!L.pre || (!L.flow && !L.mask), "pp_regression_loss: warped centres and a dense flow are exclusive");
        a.coord_q = L.coord_q; a.coord_k = L.coord_k; a.flow = L.flow; a.mask = L.mask; a.pre = L.pre;
        a.G = G; a.P = P;
        a.wo = (float)(W_orig - 1); a.ho = (float)(H_orig - 1);
        a.dG = make_div((float)G, div_mode);
        a.pr = (float)pos_ratio;
        a.warp = make_warp_args(L.flow ? Hin : 2, L.flow ? Win : 2, H_orig, W_orig, div_mode);
        a.ws = carve_ws(L.workspace, B, P);
        a.pos_num = L.pos_num; a.pos_mean = L.pos_mean; a.centres = L.centres; a.pos_mask = L.pos_mask; a.B = B;
    }
    int ntile = loss_ntile(P);
    int rc;
    if (P <= PMAX && ((C * P) % 4 == 0) && small_loss_smem_bytes(C, P) <= 226 * 1024) {
        static unsigned long long opted = 0;  // one bit per device
        if (smem_opt_in(loss_small_kernel, 227 * 1024, opted) != cudaSuccess) {
            set_error("loss_small_kernel: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
            return PP_ERR_CUDA;
        }
        SmallLossArgs2 sa2;
        for (int c = 0; c < ncall; c++) {
            SmallLossArgs& sa = sa2.seg[c];
            sa.p = pa[c]; sa.q = calls[c].q; sa.k = calls[c].k; sa.dq = calls[c].dq; sa.C = C; sa.scale = scale;
        }
        if (ncall == 1) sa2.seg[1] = sa2.seg[0];
        sa2.B = (int)B;
        ntile = 1;
        PP_LAUNCH("loss_small", st,
                  loss_small_kernel<<<(unsigned)(B * ncall), SM_THREADS, small_loss_smem_bytes(C, P), st>>>(sa2));
        rc = check_launch("loss_small_kernel");
        if (rc) return rc;
    } else if (use_tensor_cores(P) && P > PMAX && calls[0].dq && (ncall == 1 || calls[1].dq)) {
        ntile = kDotSplit;  // <= loss_ntile(P) for every P > PMAX, so the workspace's partial[] region is large enough
        for (int c = 0; c < ncall; c++) {
            uint8_t* posb = calls[c].pos_mask ? calls[c].pos_mask : pa[c].ws.posb;
            PP_LAUNCH("loss_centres", st, loss_centres_kernel<<<(unsigned)B, 256, 0, st>>>(pa[c]));
            const bool tma = loss_tc2(C, P);
            float* posf = tma ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(pa[c].ws.posb + B * (int64_t)P * P) + 15) & ~(uintptr_t)15) : nullptr;
            PP_LAUNCH("loss_pos", st, loss_pos_kernel<<<dim3((P + 7) / 8, (unsigned)B), 256, 0, st>>>(pa[c].ws, P, pa[c].pr, posb, posf));
            PP_LAUNCH("loss_cnt", st, loss_cnt_kernel<<<(unsigned)B, 256, 0, st>>>(pa[c].ws, P, pa[c].pos_num, pa[c].pos_mean));
            rc = check_launch("loss prep (large grid)");
            if (rc) return rc;
            rc = -1;
            if (tma)  // A = k as it is (split by the kernel's converter warps), B = the 0/1 plane (exact: no lo part)
                rc = tc2::launch_tc2<false>("loss M=K*pos^T (tcgen05)", B, C, P, P, calls[c].k, nullptr, posf, posf,
                                            TcStDq{calls[c].dq, pa[c].ws.den, scale, C, P}, st);
            if (rc < 0)
                rc = launch_tc("loss M=K*pos^T (tcgen05)", B, C, P, P, TcLdN{calls[c].k, C, P}, TcLdPos{posb, P},
                               TcStDq{calls[c].dq, pa[c].ws.den, scale, C, P}, st);
            if (rc) return rc;
            PP_LAUNCH("loss_dot", st, loss_dot_kernel<<<dim3(kDotSplit, (unsigned)B), 256, 0, st>>>(calls[c].q, calls[c].dq, pa[c].ws, (int64_t)C * P, scale));
            rc = check_launch("loss_dot_kernel");
            if (rc) return rc;
        }
    } else {
        for (int c = 0; c < ncall; c++) {
            PP_LAUNCH("loss_prep", st, loss_prep_kernel<<<(unsigned)B, 256, 5 * P * sizeof(float), st>>>(pa[c]));
            rc = check_launch("loss_prep_kernel");
            if (rc) return rc;
            MainArgs ma;
            ma.q = calls[c].q; ma.k = calls[c].k; ma.dq = calls[c].dq; ma.ws = pa[c].ws; ma.C = C; ma.P = P; ma.ntile = ntile;
            ma.pr = pa[c].pr; ma.scale = scale;
            PP_LAUNCH("loss_main", st, loss_main_kernel<<<dim3(ma.ntile, (unsigned)B), 256, 0, st>>>(ma));
            rc = check_launch("loss_main_kernel");
            if (rc) return rc;
        }
    }
    FinalArgs fa;
    for (int c = 0; c < 2; c++) {
        fa.ws[c] = pa[c < ncall ? c : 0].ws;
        fa.loss[c] = calls[c < ncall ? c : 0].loss;
    }
    PP_LAUNCH("loss_final", st, loss_final_kernel<<<ncall, 256, 0, st>>>(fa, B, ntile));
    return check_launch("loss_final_kernel");
}

int pp_regression_loss(const float* q, const float* k, int64_t B, int C, int G, const float* coord_q, const float* coord_k,
                       const float* flow, int Hin, int Win, const uint8_t* mask, int H_orig, int W_orig, double pos_ratio,
                       int div_mode, float* loss, float* pos_num, float* pos_mean, float* dq, uint8_t* pos_mask,
                       float* centres, void* workspace, void* stream) {
    LossCall c{q, k, coord_q, coord_k, flow, mask, loss, pos_num, pos_mean, dq, pos_mask, centres, workspace};
    return regression_loss_impl(&c, 1, B, C, G, Hin, Win, H_orig, W_orig, pos_ratio, div_mode, stream);
}

int pp_regression_loss_pair(const float* const* q, const float* const* k, int64_t B, int C, int G,
                            const float* const* coord_q, const float* const* coord_k, const float* const* flow, int Hin,
                            int Win, const uint8_t* const* mask, int H_orig, int W_orig, double pos_ratio, int div_mode,
                            float* const* loss, float* const* pos_num, float* const* pos_mean, float* const* dq,
                            void* const* workspace, void* stream) {
    PP_REQUIRE(q && k && coord_q && coord_k && flow && mask && loss && pos_num && pos_mean && dq && workspace,
               "pp_regression_loss_pair: null pointer table");
    LossCall c[2];
    for (int i = 0; i < 2; i++)
        c[i] = LossCall{q[i], k[i], coord_q[i], coord_k[i], flow[i], mask[i], loss[i], pos_num[i], pos_mean[i], dq[i],
                        nullptr, nullptr, workspace[i]};
    return regression_loss_impl(c, 2, B, C, G, Hin, Win, H_orig, W_orig, pos_ratio, div_mode, stream);
}

int pp_regression_loss_warped(const float* q, const float* k, int64_t B, int C, int G, const float* coord_q, const float* coord_k,
                              const float* warped, int H_orig, int W_orig, double pos_ratio, int div_mode, float* loss,
                              float* pos_num, float* pos_mean, float* dq, uint8_t* pos_mask, float* centres, void* workspace,
                              void* stream) {
    PP_REQUIRE(warped, "pp_regression_loss_warped: null warped centres");
    LossCall c{q, k, coord_q, coord_k, nullptr, nullptr, loss, pos_num, pos_mean, dq, pos_mask, centres, workspace, warped};
    return regression_loss_impl(&c, 1, B, C, G, 0, 0, H_orig, W_orig, pos_ratio, div_mode, stream);
}

int pp_regression_loss_pair_warped(const float* const* q, const float* const* k, int64_t B, int C, int G,
                                   const float* const* coord_q, const float* const* coord_k, const float* const* warped,
                                   int H_orig, int W_orig, double pos_ratio, int div_mode, float* const* loss,
                                   float* const* pos_num, float* const* pos_mean, float* const* dq, void* const* workspace,
                                   void* stream) {
    PP_REQUIRE(q && k && coord_q && coord_k && warped && loss && pos_num && pos_mean && dq && workspace,
               "pp_regression_loss_pair_warped: null pointer table");
    PP_REQUIRE(warped[0] && warped[1], "pp_regression_loss_pair_warped: null warped centres");
    LossCall c[2];
    for (int i = 0; i < 2; i++)
        c[i] = LossCall{q[i], k[i], coord_q[i], coord_k[i], nullptr, nullptr, loss[i], pos_num[i], pos_mean[i], dq[i],
                        nullptr, nullptr, workspace[i], warped[i]};
    return regression_loss_impl(c, 2, B, C, G, 0, 0, H_orig, W_orig, pos_ratio, div_mode, stream);
}

}  // extern "C"
