// pp_flow.cu — flow-stage kernels (sm_100a): upflow8, (de)normalisation, flow chaining
// (dense and fused-with-upsampling), forward-backward consistency masks, mask ratio.
//
// Reference functions restated (paths relative to the reference repo):
//   upflow8                         contrast/flow/utils/utils.py:87-89
//   normalize_coord/flow, denorm    contrast/util.py:334-357
//   concat_flow                     contrast/util.py:301-330
//   forward_backward_consistency    contrast/util.py:253-297
//   apply_optical_flow (flow stage) contrast/util.py:175-248
//   calc_mask_ratio                 contrast/util.py:361-366
//
// All kernels are HBM-bound pointwise / gather kernels: one thread owns one output pixel
// (or four consecutive ones where a float4 store is possible), flow links are read through
// the read-only path, and the chain state lives in registers, so each output byte is
// written exactly once and no intermediate tensor is materialised.
#include <math.h>

#include "pp_common.cuh"

namespace pp {

// ------------------------------------------------------------------------------------------
// a1 upflow8: out[p,Y,X] = 8 * bilinear(in[p], Y, X).  One thread -> 4 consecutive X.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upflow8_kernel(const float* __restrict__ in, int64_t planes, int h, int w,
                                                       float rh, float rw, float* __restrict__ out) {
    const int H = 8 * h, W = 8 * w, W4 = W >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = planes * H * W4;
    if (idx >= total) return;
    int x4 = (int)(idx % W4);
    int64_t r = idx / W4;
    int Y = (int)(r % H);
    int64_t p = r / H;
    const float* src = in + p * (int64_t)h * w;
    AxisTap ty = axis_tap(Y, rh, h);
    const float* r0 = src + (int64_t)ty.i0 * w;
    const float* r1 = src + (int64_t)ty.i1 * w;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        AxisTap tx = axis_tap(x4 * 4 + j, rw, w);
        float a = __ldg(r0 + tx.i0), b = __ldg(r0 + tx.i1), c = __ldg(r1 + tx.i0), d = __ldg(r1 + tx.i1);
        v[j] = mul(8.0f, up_combine(ty, tx, a, b, c, d));
    }
    *reinterpret_cast<float4*>(out + (p * H + Y) * (int64_t)W + x4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------------
// a2 normalise kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_kernel(const float* x, int64_t total, int64_t HW, int kind,
                                                         ScalarDiv dw, ScalarDiv dh, float* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    bool is_y = ((i / HW) & 1) != 0;
    const ScalarDiv& d = is_y ? dh : dw;
    float v = x[i];
    float r;
    if (kind == PP_NORM_COORD) r = norm_coord(v, d);
    else if (kind == PP_NORM_FLOW) r = norm_flow(v, d);
    else r = denorm_flow(v, d.s);
    out[i] = r;
}

// ------------------------------------------------------------------------------------------
// Link accessors.  A "link" is one [2,H,W] flow field of the chain for one sample.
//   DenseLink  : values are read from a materialised dense field.
//   UpLink     : values are 8 * bilinear-upsample of a low-res [2,h,w] field, evaluated on
//                the fly with ATen's exact arithmetic (never materialised).
// value(c, y, x) returns the (optionally normalised) flow component c at integer pixel.
// ------------------------------------------------------------------------------------------
struct DenseLink {
    const float* base;  // [2,H,W]
    int64_t HW;
    int W;
    __device__ __forceinline__ float2 value(int y, int x) const {
        int64_t o = (int64_t)y * W + x;
        return make_float2(__ldg(base + o), __ldg(base + HW + o));
    }
};

struct UpLink {
    const float* base;  // [2,h,w]
    int h, w;
    float rh, rw;
    __device__ __forceinline__ float2 value(int y, int x) const {
        AxisTap ty = axis_tap(y, rh, h);
        AxisTap tx = axis_tap(x, rw, w);
        const float* r0 = base + ty.i0 * w;
        const float* r1 = base + ty.i1 * w;
        int hw = h * w;
        float ax = __ldg(r0 + tx.i0), bx = __ldg(r0 + tx.i1), cx = __ldg(r1 + tx.i0), dx = __ldg(r1 + tx.i1);
        float ay = __ldg(r0 + hw + tx.i0), by = __ldg(r0 + hw + tx.i1), cy = __ldg(r1 + hw + tx.i0), dy = __ldg(r1 + hw + tx.i1);
        return make_float2(mul(8.0f, up_combine(ty, tx, ax, bx, cx, dx)), mul(8.0f, up_combine(ty, tx, ay, by, cy, dy)));
    }
};

// grid_sample of a link at normalised (gx,gy); NORM: taps are normalize_flow()'d first
// (sampling the normalised field, util.py:316-318 / :278).
template <bool NORM, class Link>
__device__ __forceinline__ float2 sample_link(const Link& L, float gx, float gy, int W, int H, float half_w, float half_h,
                                              const ScalarDiv& dw, const ScalarDiv& dh) {
    Taps t = make_taps(gx, gy, W, H, half_w, half_h);
    float2 z = make_float2(0.0f, 0.0f);
    float2 vnw = (t.inx0 && t.iny0) ? L.value(t.y0, t.x0) : z;
    float2 vne = (t.inx1 && t.iny0) ? L.value(t.y0, t.x0 + 1) : z;
    float2 vsw = (t.inx0 && t.iny1) ? L.value(t.y0 + 1, t.x0) : z;
    float2 vse = (t.inx1 && t.iny1) ? L.value(t.y0 + 1, t.x0 + 1) : z;
    if (NORM) {
        vnw.x = norm_flow(vnw.x, dw); vne.x = norm_flow(vne.x, dw); vsw.x = norm_flow(vsw.x, dw); vse.x = norm_flow(vse.x, dw);
        vnw.y = norm_flow(vnw.y, dh); vne.y = norm_flow(vne.y, dh); vsw.y = norm_flow(vsw.y, dh); vse.y = norm_flow(vse.y, dh);
    }
    return make_float2(combine(t, vnw.x, vne.x, vsw.x, vse.x), combine(t, vnw.y, vne.y, vsw.y, vse.y));
}

// ------------------------------------------------------------------------------------------
// a3 / a6 chain kernel.  blockIdx.z = sample * ndir + dir; block = 32x8 pixels.
//   UP      : links are low-res [2,h,w] fields up-sampled x8 on the fly (flow_up path)
//   IS_NORM : --flow_cat_norm arithmetic (chain in normalised units)
// ------------------------------------------------------------------------------------------
struct ChainArgs {
    const float* links[2];  // per direction
    float* out[2];          // per direction, [B,2,H,W]
    int64_t stride_n, stride_b;
    int n, H, W, h, w;
    float rh, rw, half_w, half_h;
    ScalarDiv dw, dh;
    int ndir;
};

template <bool UP, bool IS_NORM>
__global__ void __launch_bounds__(256) chain_kernel(ChainArgs a) {
    int X = blockIdx.x * 32 + threadIdx.x;
    int Y = blockIdx.y * 8 + threadIdx.y;
    if (X >= a.W || Y >= a.H) return;
    int dir = blockIdx.z % a.ndir;
    int64_t b = blockIdx.z / a.ndir;
    const float* links = (dir ? a.links[1] : a.links[0]) + b * a.stride_b;
    int64_t HW = (int64_t)a.H * a.W;
    float ox, oy;
    if (a.n == 1) {  // util.py:303-308: clone (normalised when is_norm)
        float2 v;
        if (UP) {
            UpLink L{links, a.h, a.w, a.rh, a.rw};
            v = L.value(Y, X);
        } else {
            DenseLink L{links, HW, a.W};
            v = L.value(Y, X);
        }
        ox = IS_NORM ? norm_flow(v.x, a.dw) : v.x;
        oy = IS_NORM ? norm_flow(v.y, a.dh) : v.y;
    } else {
        float c0x = (float)X, c0y = (float)Y;
        if (IS_NORM) {
            c0x = norm_coord(c0x, a.dw);
            c0y = norm_coord(c0y, a.dh);
        }
        float cx = c0x, cy = c0y;
        for (int i = 0; i < a.n; i++) {  // util.py:315-323
            const float* lp = links + i * a.stride_n;
            float gx = IS_NORM ? cx : norm_coord(cx, a.dw);
            float gy = IS_NORM ? cy : norm_coord(cy, a.dh);
            float2 s;
            if (UP) {
                UpLink L{lp, a.h, a.w, a.rh, a.rw};
                s = sample_link<IS_NORM>(L, gx, gy, a.W, a.H, a.half_w, a.half_h, a.dw, a.dh);
            } else {
                DenseLink L{lp, HW, a.W};
                s = sample_link<IS_NORM>(L, gx, gy, a.W, a.H, a.half_w, a.half_h, a.dw, a.dh);
            }
            cx = add(cx, s.x);
            cy = add(cy, s.y);
        }
        ox = sub(cx, c0x);  // util.py:326,328
        oy = sub(cy, c0y);
    }
    float* o = (dir ? a.out[1] : a.out[0]) + b * 2 * HW + (int64_t)Y * a.W + X;
    o[0] = ox;
    o[HW] = oy;
}

// ------------------------------------------------------------------------------------------
// a5 forward-backward consistency.  blockIdx.z = sample * ndir + dir.  For dir 0 the pair is
// (fwd,bwd), for dir 1 it is (bwd,fwd) (util.py:212-213).
// ------------------------------------------------------------------------------------------
struct FbArgs {
    const float* flow[2];  // [B,2,H,W] each
    uint8_t* mask[2];      // [B,H,W]
    float* cycle;          // optional (dir 0 only), [B,2,H,W]
    float* coords1;        // optional (dir 0 only)
    int H, W;
    float half_w, half_h, a1, a2;
    ScalarDiv dw, dh;
    int ndir;
};

template <bool IS_NORM>
__global__ void __launch_bounds__(256) fb_kernel(FbArgs a) {
    int X = blockIdx.x * 32 + threadIdx.x;
    int Y = blockIdx.y * 8 + threadIdx.y;
    if (X >= a.W || Y >= a.H) return;
    int dir = blockIdx.z % a.ndir;
    int64_t b = blockIdx.z / a.ndir;
    int64_t HW = (int64_t)a.H * a.W;
    const float* f = (dir ? a.flow[1] : a.flow[0]) + b * 2 * HW;
    const float* g = (dir ? a.flow[0] : a.flow[1]) + b * 2 * HW;
    int64_t i = (int64_t)Y * a.W + X;
    float fx = __ldg(f + i), fy = __ldg(f + HW + i);
    float fnx = IS_NORM ? fx : norm_flow(fx, a.dw);  // :264
    float fny = IS_NORM ? fy : norm_flow(fy, a.dh);
    float c1x = add(norm_coord((float)X, a.dw), fnx);  // :271,275
    float c1y = add(norm_coord((float)Y, a.dh), fny);
    bool inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);  // :276
    DenseLink L{g, HW, a.W};
    float2 bi;
    if (IS_NORM) bi = sample_link<false>(L, c1x, c1y, a.W, a.H, a.half_w, a.half_h, a.dw, a.dh);
    else bi = sample_link<true>(L, c1x, c1y, a.W, a.H, a.half_w, a.half_h, a.dw, a.dh);  // :278 on the normalised field
    float cyx = add(fnx, bi.x), cyy = add(fny, bi.y);                 // :279
    float cyc2 = add(mul(cyx, cyx), mul(cyy, cyy));                   // :293
    float f2 = add(mul(fnx, fnx), mul(fny, fny));
    float b2 = add(mul(bi.x, bi.x), mul(bi.y, bi.y));
    float eps = add(mul(a.a1, add(f2, b2)), a.a2);                    // :294
    bool ok = inb && (sub(cyc2, eps) <= 0.0f);                        // :296
    (dir ? a.mask[1] : a.mask[0])[b * HW + i] = ok ? 1 : 0;
    if (dir == 0) {
        if (a.cycle) { a.cycle[b * 2 * HW + i] = cyx; a.cycle[b * 2 * HW + HW + i] = cyy; }
        if (a.coords1) { a.coords1[b * 2 * HW + i] = c1x; a.coords1[b * 2 * HW + HW + i] = c1y; }
    }
}

// a11 calc_mask_ratio: one block per sample, integer count of zeros (exact), one division.
__global__ void __launch_bounds__(256) mask_ratio_kernel(const uint8_t* __restrict__ mask, int64_t HW, float* __restrict__ ratio) {
    const uint8_t* m = mask + blockIdx.x * HW;
    unsigned long long cnt = 0;
    for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) cnt += (m[i] == 0);
    __shared__ unsigned long long sm[256];
    sm[threadIdx.x] = cnt;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) ratio[blockIdx.x] = (float)((double)sm[0] / (double)HW);
}

static float fb_alpha2_eff(double alpha_2, int H, int W) {
    // util.py:289-291: sqrt of an int64 tensor -> fp32; .item() -> double; python division
    float r = sqrtf((float)((int64_t)H * H + (int64_t)W * W));
    return (float)(alpha_2 / (double)r);
}

static int launch_chain(const float* l0, const float* l1, float* o0, float* o1, int ndir, int n, int64_t B, int H, int W,
                        int h, int w, bool up, int64_t stride_n, int64_t stride_b, int is_norm, int div_mode,
                        cudaStream_t st) {
    ChainArgs a;
    a.links[0] = l0; a.links[1] = l1; a.out[0] = o0; a.out[1] = o1;
    a.stride_n = stride_n; a.stride_b = stride_b;
    a.n = n; a.H = H; a.W = W; a.h = h; a.w = w;
    a.rh = up ? up_scale(h, H) : 0.f; a.rw = up ? up_scale(w, W) : 0.f;
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.dw = make_div((float)(W - 1), div_mode); a.dh = make_div((float)(H - 1), div_mode);
    a.ndir = ndir;
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, (unsigned)(B * ndir));
    if (up) {
        if (is_norm) PP_LAUNCH("chain_up_norm", st, chain_kernel<true, true><<<grid, block, 0, st>>>(a));
        else PP_LAUNCH("chain_up", st, chain_kernel<true, false><<<grid, block, 0, st>>>(a));
    } else {
        if (is_norm) PP_LAUNCH("chain_dense_norm", st, chain_kernel<false, true><<<grid, block, 0, st>>>(a));
        else PP_LAUNCH("chain_dense", st, chain_kernel<false, false><<<grid, block, 0, st>>>(a));
    }
    return check_launch("chain_kernel");
}

static int launch_fb(const float* f0, const float* f1, uint8_t* m0, uint8_t* m1, float* cycle, float* coords1, int ndir,
                     int64_t B, int H, int W, double alpha_1, double alpha_2, int is_norm, int div_mode, cudaStream_t st) {
    FbArgs a;
    a.flow[0] = f0; a.flow[1] = f1; a.mask[0] = m0; a.mask[1] = m1;
    a.cycle = cycle; a.coords1 = coords1; a.H = H; a.W = W;
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.a1 = (float)alpha_1; a.a2 = fb_alpha2_eff(alpha_2, H, W);
    a.dw = make_div((float)(W - 1), div_mode); a.dh = make_div((float)(H - 1), div_mode);
    a.ndir = ndir;
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, (unsigned)(B * ndir));
    if (is_norm) PP_LAUNCH("fb_norm", st, fb_kernel<true><<<grid, block, 0, st>>>(a));
    else PP_LAUNCH("fb", st, fb_kernel<false><<<grid, block, 0, st>>>(a));
    return check_launch("fb_kernel");
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_upflow8(const float* in, int64_t planes, int h, int w, float* out, void* stream) {
    PP_REQUIRE(planes >= 0 && h > 0 && w > 0, "pp_upflow8: bad shape planes=%lld h=%d w=%d", (long long)planes, h, w);
    if (planes == 0) return PP_OK;  // empty batch: nothing to do (pointers may be null)
    PP_REQUIRE(in && out, "pp_upflow8: null pointer");
    int64_t total = planes * 8 * h * (2 * w);
    cudaStream_t st = (cudaStream_t)stream;
    PP_LAUNCH("upflow8", st,
              upflow8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, planes, h, w, up_scale(h, 8 * h),
                                                                               up_scale(w, 8 * w), out));
    return check_launch("upflow8_kernel");
}

int pp_normalize(const float* x, int64_t B, int H, int W, int kind, int div_mode, float* out, void* stream) {
    PP_REQUIRE(B >= 0 && H > 1 && W > 1, "pp_normalize: bad shape B=%lld H=%d W=%d", (long long)B, H, W);
    PP_REQUIRE(kind >= 0 && kind <= 2, "pp_normalize: bad kind %d", kind);
    if (B == 0) return PP_OK;
    PP_REQUIRE(x && out, "pp_normalize: null pointer");
    int64_t HW = (int64_t)H * W, total = B * 2 * HW;
    cudaStream_t st = (cudaStream_t)stream;
    PP_LAUNCH("normalize", st,
              normalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                  x, total, HW, kind, make_div((float)(W - 1), div_mode), make_div((float)(H - 1), div_mode), out));
    return check_launch("normalize_kernel");
}

int pp_concat_flow(const float* flows, int n, int64_t B, int H, int W, int64_t stride_n, int64_t stride_b, int is_norm,
                   int div_mode, float* out, void* stream) {
    PP_REQUIRE(n >= 1 && B >= 0 && H > 1 && W > 1, "pp_concat_flow: bad shape n=%d B=%lld H=%d W=%d", n, (long long)B, H, W);
    PP_REQUIRE(B <= 65535, "pp_concat_flow: B=%lld exceeds 65535", (long long)B);
    if (B == 0) return PP_OK;
    PP_REQUIRE(flows && out, "pp_concat_flow: null pointer");
    return launch_chain(flows, nullptr, out, nullptr, 1, n, B, H, W, 0, 0, false, stride_n, stride_b, is_norm, div_mode,
                        (cudaStream_t)stream);
}

int pp_fb_consistency(const float* fwd, const float* bwd, int64_t B, int H, int W, double alpha_1, double alpha_2,
                      int is_norm, int div_mode, uint8_t* mask, float* cycle, float* coords1_norm, void* stream) {
    PP_REQUIRE(B >= 0 && H > 1 && W > 1, "pp_fb_consistency: bad shape B=%lld H=%d W=%d", (long long)B, H, W);
    PP_REQUIRE(B <= 65535, "pp_fb_consistency: B=%lld exceeds 65535", (long long)B);
    if (B == 0) return PP_OK;
    PP_REQUIRE(fwd && bwd && mask, "pp_fb_consistency: null pointer");
    return launch_fb(fwd, bwd, mask, nullptr, cycle, coords1_norm, 1, B, H, W, alpha_1, alpha_2, is_norm, div_mode,
                     (cudaStream_t)stream);
}

int pp_flow_stage(const float* lo_fwd, const float* lo_bwd, int64_t B, int n, int h, int w, int flow_up, int use_mask,
                  double alpha_1, double alpha_2, int is_norm, int div_mode, float* flow_fwd, float* flow_bwd,
                  uint8_t* mask_fwd, uint8_t* mask_bwd, void* stream) {
    PP_REQUIRE(n >= 1 && B >= 0 && h > 1 && w > 1, "pp_flow_stage: bad shape B=%lld n=%d h=%d w=%d", (long long)B, n, h, w);
    PP_REQUIRE(B * 2 <= 65535, "pp_flow_stage: B=%lld exceeds 32767", (long long)B);
    if (B == 0) return PP_OK;
    PP_REQUIRE(lo_fwd && lo_bwd && flow_fwd && flow_bwd, "pp_flow_stage: null pointer");
    PP_REQUIRE(!use_mask || (mask_fwd && mask_bwd), "pp_flow_stage: use_mask set but mask outputs are null");
    cudaStream_t st = (cudaStream_t)stream;
    int H = flow_up ? 8 * h : h, W = flow_up ? 8 * w : w;
    int64_t link = 2 * (int64_t)h * w;  // loader layout [B,n,2,h,w]
    int rc = launch_chain(lo_fwd, lo_bwd, flow_fwd, flow_bwd, 2, n, B, H, W, h, w, flow_up != 0, link, (int64_t)n * link,
                          is_norm, div_mode, st);
    if (rc) return rc;
    if (use_mask) {
        rc = launch_fb(flow_fwd, flow_bwd, mask_fwd, mask_bwd, nullptr, nullptr, 2, B, H, W, alpha_1, alpha_2, is_norm,
                       div_mode, st);
        if (rc) return rc;
    }
    if (is_norm) {  // util.py:229-231
        rc = pp_normalize(flow_fwd, B, H, W, PP_DENORM_FLOW, div_mode, flow_fwd, stream);
        if (rc) return rc;
        rc = pp_normalize(flow_bwd, B, H, W, PP_DENORM_FLOW, div_mode, flow_bwd, stream);
        if (rc) return rc;
    }
    return PP_OK;
}

int pp_calc_mask_ratio(const uint8_t* mask, int64_t B, int H, int W, float* ratio, void* stream) {
    PP_REQUIRE(B >= 0 && H > 0 && W > 0, "pp_calc_mask_ratio: bad shape");
    if (B == 0) return PP_OK;
    PP_REQUIRE(mask && ratio, "pp_calc_mask_ratio: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    PP_LAUNCH("mask_ratio", st, mask_ratio_kernel<<<(unsigned)B, 256, 0, st>>>(mask, (int64_t)H * W, ratio));
    return check_launch("mask_ratio_kernel");
}

}  // extern "C"
