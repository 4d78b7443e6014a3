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
// All kernels are pointwise / gather kernels whose DRAM traffic equals the algorithmic bytes
// (ncu: profiles/r01_a_flow_stage_ncu_summary.txt); what limits them is instruction issue, so
// the hot variants (a) own four consecutive pixels per thread (float4 / uchar4 accesses, row
// terms amortised), (b) address with 32-bit offsets from per-thread base pointers, and (c) use
// the certified 3-instruction exact division of pp_common.cuh instead of the IEEE sequence.
// The chain state lives in registers; no intermediate tensor is materialised.
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pp_common.cuh"
#include "pp_warp.cuh"

namespace pp {

// ------------------------------------------------------------------------------------------
// a1 upflow8 (stand-alone): out[p,Y,X] = 8 * bilinear(in[p], Y, X).  One thread -> 4 X.
// ------------------------------------------------------------------------------------------
// Evaluates 4 consecutive output columns X4..X4+3 of one output row for one low-res plane.
// The 4 columns touch at most 3 consecutive low-res columns (scale < 1/8), which are loaded
// once per row and selected per pixel — same values as 16 separate loads.
struct Up4 {
    AxisTap tx[4];
    int c0;
    bool has1, has2;  // columns c0+1 / c0+2 exist (else the value is the clamped neighbour)
    __device__ __forceinline__ void init(int X4, float rw, int w) {
#pragma unroll
        for (int j = 0; j < 4; j++) tx[j] = axis_tap(X4 + j, rw, w);
        c0 = tx[0].i0;
        has1 = c0 + 1 <= w - 1;
        has2 = c0 + 2 <= w - 1;
    }
    // r0/r1: pointers to column c0 of the two low-res rows of one plane
    __device__ __forceinline__ void eval(const float* __restrict__ r0, const float* __restrict__ r1, const AxisTap& ty,
                                         float out[4]) const {
        float a0 = __ldg(r0), b0 = __ldg(r1);
        float a1 = a0, b1 = b0;
        if (has1) { a1 = __ldg(r0 + 1); b1 = __ldg(r1 + 1); }
        float a2 = a1, b2 = b1;
        if (has2) { a2 = __ldg(r0 + 2); b2 = __ldg(r1 + 2); }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            bool d = tx[j].i0 != c0;  // 0 or 1 column to the right of c0
            float a = d ? a1 : a0, b = d ? a2 : a1, c = d ? b1 : b0, e = d ? b2 : b1;
            out[j] = mul(8.0f, up_combine(ty, tx[j], a, b, c, e));
        }
    }
};

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
    Up4 u;
    u.init(x4 * 4, rw, w);
    float v[4];
    u.eval(ptr_at(src, ty.i0 * w + u.c0), ptr_at(src, ty.i1 * w + u.c0), ty, v);
    *reinterpret_cast<float4*>(out + (p * H + Y) * (int64_t)W + x4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------------
// a2 normalise kernels (stand-alone API; guarded division: exact for every input)
// ------------------------------------------------------------------------------------------
template <int DM>
__global__ void __launch_bounds__(256) normalize_kernel(const float* x, int64_t total, int64_t HW, int kind, Div<DM> dw,
                                                         Div<DM> dh, float* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    bool is_y = ((i / HW) & 1) != 0;
    float v = x[i];
    float r;
    if (kind == PP_NORM_COORD) r = is_y ? norm_coord(v, dh) : norm_coord(v, dw);
    else if (kind == PP_NORM_FLOW) r = is_y ? norm_flow(v, dh) : norm_flow(v, dw);
    else r = denorm_flow(v, is_y ? dh.s : dw.s);
    out[i] = r;
}

// ------------------------------------------------------------------------------------------
// Link accessors.  A "link" is one [2,H,W] flow field of the chain for one sample.
//   DenseLink : values are read from a materialised dense field.
//   UpLink    : values are 8 * bilinear-upsample of a low-res [2,h,w] field, evaluated on the
//               fly with ATen's exact arithmetic (never materialised).
// ------------------------------------------------------------------------------------------
struct DenseLink {
    const float* base;  // [2,H,W]
    int HW, W;
    __device__ __forceinline__ float2 value(int y, int x) const {
        int o = y * W + x;
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
        int o0 = ty.i0 * w, o1 = ty.i1 * w, hw = h * w;
        const float* p = base;
        float ax = __ldg(p + o0 + tx.i0), bx = __ldg(p + o0 + tx.i1), cx = __ldg(p + o1 + tx.i0), dx = __ldg(p + o1 + tx.i1);
        p += hw;
        float ay = __ldg(p + o0 + tx.i0), by = __ldg(p + o0 + tx.i1), cy = __ldg(p + o1 + tx.i0), dy = __ldg(p + o1 + tx.i1);
        return make_float2(mul(8.0f, up_combine(ty, tx, ax, bx, cx, dx)), mul(8.0f, up_combine(ty, tx, ay, by, cy, dy)));
    }
};

// grid_sample of a link at normalised (gx,gy); NORM: taps are normalize_flow()'d first
// (sampling the normalised field, util.py:316-318 / :278).
template <bool NORM, int DM, class Link>
__device__ __forceinline__ float2 sample_link(const Link& L, float gx, float gy, int W, int H, float half_w, float half_h,
                                              const Div<DM>& dw, const Div<DM>& dh) {
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
// a3 / a6 chain kernels.  blockIdx.z = sample * ndir + dir.
// ------------------------------------------------------------------------------------------
template <int DM>
struct ChainArgs {
    const float* links[2];  // per direction
    float* out[2];          // per direction, [B,2,H,W]
    int64_t stride_n, stride_b;
    int n, H, W, h, w;
    float rh, rw, half_w, half_h;
    Div<DM> dw, dh;
    int ndir;
};

// One pixel of the composite flow (util.py:301-330): n == 1 clones link 0, else the chain of n
// grid_samples starting at the pixel's own coordinate.
//   UP: links are low-res fields up-sampled x8 on the fly;  IS_NORM: --flow_cat_norm arithmetic
template <bool UP, bool IS_NORM, int DM>
__device__ __forceinline__ float2 chain_pixel(const ChainArgs<DM>& a, const float* links, int X, int Y) {
    const int HW = a.H * a.W;
    if (a.n == 1) {  // util.py:303-308: clone (normalised when is_norm)
        float2 v;
        if (UP) {
            UpLink L{links, a.h, a.w, a.rh, a.rw};
            v = L.value(Y, X);
        } else {
            DenseLink L{links, HW, a.W};
            v = L.value(Y, X);
        }
        return make_float2(IS_NORM ? norm_flow(v.x, a.dw) : v.x, IS_NORM ? norm_flow(v.y, a.dh) : v.y);
    }
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
    return make_float2(sub(cx, c0x), sub(cy, c0y));  // util.py:326,328
}

// generic chain: one thread per pixel; block 32x8.
template <bool UP, bool IS_NORM, int DM>
__global__ void __launch_bounds__(256) chain_kernel(ChainArgs<DM> a) {
    int X = blockIdx.x * 32 + threadIdx.x;
    int Y = blockIdx.y * 8 + threadIdx.y;
    if (X >= a.W || Y >= a.H) return;
    int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    int64_t b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const float* links = (dir ? a.links[1] : a.links[0]) + b * a.stride_b;
    int HW = a.H * a.W;
    const float2 v = chain_pixel<UP, IS_NORM, DM>(a, links, X, Y);
    float* o = (dir ? a.out[1] : a.out[0]) + b * 2 * (int64_t)HW + Y * a.W + X;
    o[0] = v.x;
    o[HW] = v.y;
}

// Dense-link chain, 4 pixels per thread (X0 + lane + 32*j: every gather of a warp covers 32
// consecutive pixels), not normalised, W % 128 == 0.  Interior samples (all four taps inside
// the frame) take a branch-free path: clamped coordinates, 8 loads off one pointer (immediate
// offsets when the frame size is a compile-time constant), the four chains of a thread
// interleaved by the scheduler.  Samples touching the border are recomputed by the fully
// predicated sample_link (rare, divergent).  Same arithmetic as chain_kernel<false,false,DM>.
template <int DM, int WC, int HC>
__global__ void __launch_bounds__(256) chain_dense4_kernel(ChainArgs<DM> a) {
    const int X0 = blockIdx.x * 128 + threadIdx.x;
    const int Y = blockIdx.y * 8 + threadIdx.y;
    const int W = WC ? WC : a.W, H = HC ? HC : a.H, HW = H * W;
    if (WC) {
        a.half_w = (float)(WC - 1) / 2.0f; a.half_h = (float)(HC - 1) / 2.0f;
        a.dw = const_div<DM>((float)(WC - 1));
        a.dh = const_div<DM>((float)(HC - 1));
    }
    if (Y >= H) return;
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int64_t b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const float* links = (dir ? a.links[1] : a.links[0]) + b * a.stride_b;
    const unsigned xlim = W - 2, ylim = H - 2;
    float cx[4], cy[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { cx[j] = (float)(X0 + 32 * j); cy[j] = (float)Y; }
    for (int i = 0; i < a.n; i++) {  // util.py:315-323
        const float* lp = links + i * a.stride_n;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float gx = norm_coord(cx[j], a.dw), gy = norm_coord(cy[j], a.dh);
            const float ix = mul(add(gx, 1.0f), a.half_w), iy = mul(add(gy, 1.0f), a.half_h);
            const float xw = floorf(ix), yn = floorf(iy);
            const float wgt = sub(ix, xw), e = sub(add(xw, 1.0f), ix), n_ = sub(iy, yn), s_ = sub(add(yn, 1.0f), iy);
            const float nw = mul(s_, e), ne = mul(s_, wgt), sw = mul(n_, e), se = mul(n_, wgt);
            const int x0 = (int)xw, y0 = (int)yn;
            // interior <=> 0 <= x0 <= W-2 and 0 <= y0 <= H-2 (float compares first: huge / NaN coordinates are not interior)
            const bool interior = (xw >= 0.0f) && (yn >= 0.0f) && ((unsigned)x0 <= xlim) && ((unsigned)y0 <= ylim);
            const unsigned xc = min((unsigned)x0, xlim), yc = min((unsigned)y0, ylim);
            const float* p0 = ptr_at(lp, (int)(yc * W + xc));
            float x00, x01, x10, x11, y00, y01, y10, y11;
            if (WC) {
                x00 = __ldg(p0); x01 = __ldg(p0 + 1); x10 = __ldg(p0 + WC); x11 = __ldg(p0 + WC + 1);
                y00 = __ldg(p0 + WC * HC); y01 = __ldg(p0 + WC * HC + 1); y10 = __ldg(p0 + WC * HC + WC); y11 = __ldg(p0 + WC * HC + WC + 1);
            } else {
                const float* p1 = ptr_at(p0, W);
                const float* q0 = ptr_at(p0, HW);
                const float* q1 = ptr_at(q0, W);
                x00 = __ldg(p0); x01 = __ldg(p0 + 1); x10 = __ldg(p1); x11 = __ldg(p1 + 1);
                y00 = __ldg(q0); y01 = __ldg(q0 + 1); y10 = __ldg(q1); y11 = __ldg(q1 + 1);
            }
            float sx = combine4(x00, x01, x10, x11, nw, ne, sw, se);
            float sy = combine4(y00, y01, y10, y11, nw, ne, sw, se);
            if (!interior) {
                DenseLink L{lp, HW, W};
                float2 sl = sample_link<false>(L, gx, gy, W, H, a.half_w, a.half_h, a.dw, a.dh);
                sx = sl.x;
                sy = sl.y;
            }
            cx[j] = add(cx[j], sx);
            cy[j] = add(cy[j], sy);
        }
    }
    float* o = ptr_at((dir ? a.out[1] : a.out[0]) + b * 2 * (int64_t)HW, Y * W + X0);
    float* oy = ptr_at(o, HW);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        o[32 * j] = sub(cx[j], (float)(X0 + 32 * j));  // util.py:328
        oy[32 * j] = sub(cy[j], (float)Y);
    }
}

// n == 1 with flow_up (the published n_frames=2 setting): the composite IS the up-sampled
// link.  ATen's bilinear kernel interpolates horizontally first:
//     val = fma(l0y, T(i0y,X), l1y * T(i1y,X)),   T(r,X) = fma(l0x, L[r][i0x], l1x * L[r][i1x])
// and T(r,X) depends only on the low-res ROW r, which 8 consecutive output rows share.  One
// thread therefore owns a 4-pixel-wide, 8-row-tall strip of both channels: it evaluates T
// for the (at most) 3 low-res rows the strip touches once — bit-identical to recomputing it
// per pixel — and each output value then costs 3 flops instead of 7; column taps are computed
// once per strip.  Block = 32x8 threads = a 128x64-pixel tile; every store is a float4 and a
// warp writes 512 contiguous bytes.  ~15 instructions per output pixel: HBM-write-bound.
template <int DM>
__global__ void __launch_bounds__(256) upchain1_kernel(ChainArgs<DM> a) {
    const int X4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int Y0 = (blockIdx.y * 8 + threadIdx.y) * 8;
    if (X4 >= a.W || Y0 >= a.H) return;
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int64_t b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const float* lo = (dir ? a.links[1] : a.links[0]) + b * a.stride_b;
    const int w = a.w, hw = a.h * a.w;
    // column taps of the 4 pixels; they touch low-res columns c0 .. c0+2 (clamped)
    AxisTap tx[4];
#pragma unroll
    for (int j = 0; j < 4; j++) tx[j] = axis_tap(X4 + j, a.rw, w);
    const int c0 = tx[0].i0;
    const bool has1 = c0 + 1 <= w - 1, has2 = c0 + 2 <= w - 1;
    // low-res rows r0 .. r0+2 (clamped) cover the 8 output rows
    const int r0 = axis_tap(Y0, a.rh, a.h).i0;
    float T[2][3][4];  // [channel][low-res row][pixel]
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int rr = min(r0 + r, a.h - 1);
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            const float* p = ptr_at(lo, ch * hw + rr * w + c0);
            float v0 = __ldg(p), v1 = v0, v2;
            if (has1) v1 = __ldg(p + 1);
            v2 = v1;
            if (has2) v2 = __ldg(p + 2);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool d = tx[j].i0 != c0;  // 0 or 1 column to the right of c0
                T[ch][r][j] = fma_(tx[j].l0, d ? v1 : v0, mul(tx[j].l1, d ? v2 : v1));
            }
        }
    }
    const int HW = a.H * a.W;
    float* o = ptr_at((dir ? a.out[1] : a.out[0]) + b * 2 * (int64_t)HW, Y0 * a.W + X4);
    // T[.][dy] / T[.][dy+1] are the rows i0y / i1y of output row k (the clamp of i1y at the last
    // low-res row is already in T).  dy is uniform across the block row, so the two cases are a
    // uniform branch around straight-line code, not per-value selects.
    auto emit_row = [&](int k, const AxisTap& ty, const float (&A0)[4], const float (&A1)[4], const float (&B0)[4],
                        const float (&B1)[4]) {
        float vx[4], vy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            vx[j] = mul(8.0f, fma_(ty.l0, A0[j], mul(ty.l1, A1[j])));
            vy[j] = mul(8.0f, fma_(ty.l0, B0[j], mul(ty.l1, B1[j])));
        }
        float* ok = ptr_at(o, k * a.W);
        *reinterpret_cast<float4*>(ok) = make_float4(vx[0], vx[1], vx[2], vx[3]);
        *reinterpret_cast<float4*>(ptr_at(ok, HW)) = make_float4(vy[0], vy[1], vy[2], vy[3]);
    };
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (Y0 + k >= a.H) break;
        const AxisTap ty = axis_tap(Y0 + k, a.rh, a.h);
        if (ty.i0 != r0) emit_row(k, ty, T[0][1], T[0][2], T[1][1], T[1][2]);
        else emit_row(k, ty, T[0][0], T[0][1], T[1][0], T[1][1]);
    }
}

// ------------------------------------------------------------------------------------------
// a5 forward-backward consistency.  blockIdx.z = sample * ndir + dir.  For dir 0 the pair is
// (fwd,bwd), for dir 1 it is (bwd,fwd) (util.py:212-213).
// ------------------------------------------------------------------------------------------
template <int DM>
struct FbArgs {
    const float* flow[2];  // [B,2,H,W] each
    uint8_t* mask[2];      // [B,H,W]
    float* cycle;          // optional (dir 0 only), [B,2,H,W]
    float* coords1;        // optional (dir 0 only)
    int H, W;
    float half_w, half_h, a1, a2;
    Div<DM> dw, dh;    // / (W-1), / (H-1)
    Div<DM> dw2, dh2;  // / ((W-1)/2), / ((H-1)/2): fused 2*v/s (fbmask4_kernel only)
    int ndir;
};

// generic: one thread per pixel, every optional output, any W.
template <bool IS_NORM, int DM>
__global__ void __launch_bounds__(256) fb_kernel(FbArgs<DM> a) {
    int X = blockIdx.x * 32 + threadIdx.x;
    int Y = blockIdx.y * 8 + threadIdx.y;
    if (X >= a.W || Y >= a.H) return;
    int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    int64_t b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    int HW = a.H * a.W;
    const float* f = (dir ? a.flow[1] : a.flow[0]) + b * 2 * (int64_t)HW;
    const float* g = (dir ? a.flow[0] : a.flow[1]) + b * 2 * (int64_t)HW;
    int i = Y * a.W + X;
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
        if (a.cycle) { a.cycle[b * 2 * (int64_t)HW + i] = cyx; a.cycle[b * 2 * (int64_t)HW + HW + i] = cyy; }
        if (a.coords1) { a.coords1[b * 2 * (int64_t)HW + i] = c1x; a.coords1[b * 2 * (int64_t)HW + HW + i] = c1y; }
    }
}

// One pixel of the FB test with fully predicated taps (frame-edge cases of fbmask4_kernel).
template <int DM>
__device__ __noinline__ bool fb_pixel_edge(const float* g, int W, int H, int HW, float c1x, float c1y, float fnx, float fny,
                                           float half_w, float half_h, float a1, float a2, Div<DM> dw, Div<DM> dh) {
    DenseLink L{g, HW, W};
    float2 bi = sample_link<true>(L, c1x, c1y, W, H, half_w, half_h, dw, dh);
    float cyx = add(fnx, bi.x), cyy = add(fny, bi.y);
    float cyc2 = add(mul(cyx, cyx), mul(cyy, cyy));
    float f2 = add(mul(fnx, fnx), mul(fny, fny));
    float b2 = add(mul(bi.x, bi.x), mul(bi.y, bi.y));
    float eps = add(mul(a1, add(f2, b2)), a2);
    return sub(cyc2, eps) <= 0.0f;
}

// Mask-only, W % 128 == 0, unnormalised inputs (the flow-stage path).  Block = 32x8 threads =
// a 128x8-pixel tile; one thread -> the 4 pixels X0 + lane + 32*j of its row, so that every
// load instruction of a warp covers 32 CONSECUTIVE pixels: own flow and mask accesses are
// fully coalesced, and each tap gather touches 1-2 lines instead of the 4+ of a 4-px-per-lane
// layout (profiles/r01_c_*: that layout was LSU-bound, lg_throttle 7.5).  Only the boolean
// leaves this kernel, so
//  * a pixel whose warped position is outside (-1,1)^2 is 0 (util.py:276,296); its taps are
//    still fetched, from a clamped (valid) address, and ignored: no divergent control flow;
//  * inside, ix in (0, W-1] and iy in (0, H-1]: all four taps exist unless ix == W-1 or
//    iy == H-1 exactly, which takes the predicated edge routine;
//  * DM_FAST (unguarded exact division) is admissible: inputs outside its certified range are
//    non-finite / < 2^-100 and cannot change the comparison.
//  * WC/HC > 0: the frame size is a compile-time constant (the 1280x720 frames of the published
//    BDD100K runs), so all eight tap addresses are immediates off one pointer and the division
//    constants are instruction immediates (no per-pixel constant-bank loads).
template <int DM, int WC, int HC>
__global__ void __launch_bounds__(256) fbmask4_kernel(FbArgs<DM> a) {
    const int X0 = blockIdx.x * 128 + threadIdx.x;
    const int Y = blockIdx.y * 8 + threadIdx.y;
    const int W = WC ? WC : a.W, H = HC ? HC : a.H, HW = H * W;
    if (WC) {  // fold every derived constant
        a.half_w = (float)(WC - 1) / 2.0f; a.half_h = (float)(HC - 1) / 2.0f;
        a.dw2 = const_div<DM>((float)(WC - 1) / 2.0f);
        a.dh2 = const_div<DM>((float)(HC - 1) / 2.0f);
        a.dw = const_div<DM>((float)(WC - 1));
        a.dh = const_div<DM>((float)(HC - 1));
    }
    if (Y >= H) return;
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int64_t b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const float* f = (dir ? a.flow[1] : a.flow[0]) + b * 2 * (int64_t)HW;
    const float* g = (dir ? a.flow[0] : a.flow[1]) + b * 2 * (int64_t)HW;
    const int i = Y * W + X0;
    const float* fpx = ptr_at(f, i);
    const float* fpy = ptr_at(fpx, HW);
    float fxs[4], fys[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { fxs[j] = __ldg(fpx + 32 * j); fys[j] = __ldg(fpy + 32 * j); }
    const float yn = norm_coord_h((float)Y, a.dh2);
    const unsigned xlim = W - 2, ylim = H - 2;
    uint8_t* mp = (dir ? a.mask[1] : a.mask[0]) + b * HW + i;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float fnx = norm_flow_h(fxs[j], a.dw2), fny = norm_flow_h(fys[j], a.dh2);   // :264
        float c1x = add(norm_coord_h((float)(X0 + 32 * j), a.dw2), fnx);           // :271,275
        float c1y = add(yn, fny);
        bool inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);                     // :276
        float ix = mul(add(c1x, 1.0f), a.half_w), iy = mul(add(c1y, 1.0f), a.half_h);
        float xw = floorf(ix), yn_ = floorf(iy);
        float wgt = sub(ix, xw), e = sub(add(xw, 1.0f), ix), n = sub(iy, yn_), s_ = sub(add(yn_, 1.0f), iy);
        float nw = mul(s_, e), ne = mul(s_, wgt), sw = mul(n, e), se = mul(n, wgt);
        int x0 = (int)xw, y0 = (int)yn_;
        bool edge = inb && ((unsigned)x0 > xlim || (unsigned)y0 > ylim);
        // clamp: garbage coordinates of out-of-frame pixels still address valid memory
        unsigned xc = min((unsigned)x0, xlim), yc = min((unsigned)y0, ylim);
        const float* p0 = ptr_at(g, (int)(yc * W + xc));  // row y0, x channel
        float x00, x01, x10, x11, y00, y01, y10, y11;
        if (WC) {  // immediate offsets off one pointer
            x00 = __ldg(p0); x01 = __ldg(p0 + 1); x10 = __ldg(p0 + WC); x11 = __ldg(p0 + WC + 1);
            y00 = __ldg(p0 + WC * HC); y01 = __ldg(p0 + WC * HC + 1); y10 = __ldg(p0 + WC * HC + WC); y11 = __ldg(p0 + WC * HC + WC + 1);
        } else {
            const float* p1 = ptr_at(p0, W);   // row y0+1
            const float* q0 = ptr_at(p0, HW);  // y channel
            const float* q1 = ptr_at(q0, W);
            x00 = __ldg(p0); x01 = __ldg(p0 + 1); x10 = __ldg(p1); x11 = __ldg(p1 + 1);
            y00 = __ldg(q0); y01 = __ldg(q0 + 1); y10 = __ldg(q1); y11 = __ldg(q1 + 1);
        }
        float bx = combine4(norm_flow_h(x00, a.dw2), norm_flow_h(x01, a.dw2), norm_flow_h(x10, a.dw2),
                                          norm_flow_h(x11, a.dw2), nw, ne, sw, se);
        float by = combine4(norm_flow_h(y00, a.dh2), norm_flow_h(y01, a.dh2), norm_flow_h(y10, a.dh2),
                                          norm_flow_h(y11, a.dh2), nw, ne, sw, se);
        float cyx = add(fnx, bx), cyy = add(fny, by);                     // :279
        float cyc2 = add(mul(cyx, cyx), mul(cyy, cyy));                   // :293
        float f2 = add(mul(fnx, fnx), mul(fny, fny));
        float b2 = add(mul(bx, bx), mul(by, by));
        float eps = add(mul(a.a1, add(f2, b2)), a.a2);                    // :294
        bool ok = inb && (sub(cyc2, eps) <= 0.0f);                        // :296
        if (edge) ok = fb_pixel_edge<DM>(g, W, H, HW, c1x, c1y, fnx, fny, a.half_w, a.half_h, a.a1, a.a2, a.dw, a.dh);
        mp[32 * j] = ok ? 1 : 0;
    }
}

// One whole pixel of the mask from scratch, generic predicated taps (rare fix-up path of the packed
// kernels: pixels that land exactly on the last row / column).  Same arithmetic as fb_kernel.
template <int DM>
__device__ __noinline__ void fb_pixel_redo(const float* f, const float* g, uint8_t* m, int X, int Y, int W, int H, float half_w,
                                           float half_h, float a1, float a2, Div<DM> dw, Div<DM> dh) {
    const int HW = H * W, i = Y * W + X;
    const float fnx = norm_flow(__ldg(f + i), dw), fny = norm_flow(__ldg(f + HW + i), dh);
    const float c1x = add(norm_coord((float)X, dw), fnx), c1y = add(norm_coord((float)Y, dh), fny);
    const bool inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);
    m[i] = (inb && fb_pixel_edge<DM>(g, W, H, HW, c1x, c1y, fnx, fny, half_w, half_h, a1, a2, dw, dh)) ? 1 : 0;
}

// fbmask4_kernel on packed fp32 pairs (certified-division mode only).  Same pixel layout (one
// thread -> pixels X0 + lane + 32*j of its row); pixels (j, j+1) form the two halves of every FFMA2 /
// FMUL2 / FADD2, which halves the issue slots of the ~60 exactly-rounded operations per pixel
// (profiles/r01_n_flow_gather_ncu_summary.txt: the scalar kernel is issue-bound, 119 instructions per
// pixel).  Three more instruction savings, all value-preserving:
//  * the 2-instruction certified division (pp_common.cuh);
//  * floor as F2I.FLOOR + I2FP (one conversion-unit op instead of FRND + F2I);
//  * east/south weights as 1 - w instead of (floor+1) - i: for an in-frame pixel i >= 0, so
//    w = i - floor(i) is exact and both forms round the same real number 1 - w (a pixel outside the
//    frame is masked off whatever its weights are).
// Products that the reference rounds before adding (squares, alpha_1 * sum) are added with scalar
// adds: see the contraction caveat at F2 in pp_common.cuh.
template <int WC, int HC>
__global__ void __launch_bounds__(256) fbmask4p_kernel(FbArgs<DM_FAST> a) {
    constexpr int DM = DM_FAST;
    const int X0 = blockIdx.x * 128 + threadIdx.x;
    const int Y = blockIdx.y * 8 + threadIdx.y;
    const int W = WC ? WC : a.W, H = HC ? HC : a.H, HW = H * W;
    if (WC) {  // fold every derived constant
        a.half_w = (float)(WC - 1) / 2.0f; a.half_h = (float)(HC - 1) / 2.0f;
        a.dw2 = const_div<DM>((float)(WC - 1) / 2.0f);
        a.dh2 = const_div<DM>((float)(HC - 1) / 2.0f);
        a.dw = const_div<DM>((float)(WC - 1));
        a.dh = const_div<DM>((float)(HC - 1));
    }
    if (Y >= H) return;
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int64_t b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const float* f = (dir ? a.flow[1] : a.flow[0]) + b * 2 * (int64_t)HW;
    const float* g = (dir ? a.flow[0] : a.flow[1]) + b * 2 * (int64_t)HW;
    const int i = Y * W + X0;
    const float* fpx = ptr_at(f, i);
    const float* fpy = ptr_at(fpx, HW);
    float fxs[4], fys[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { fxs[j] = __ldg(fpx + 32 * j); fys[j] = __ldg(fpy + 32 * j); }
    const Div2 dW = make_div2(a.dw2), dH = make_div2(a.dh2);
    const F2 one2 = pk1(1.0f), hw2 = pk1(a.half_w), hh2 = pk1(a.half_h);
    const F2 yn2 = pk1(norm_coord_h((float)Y, a.dh2));
    const unsigned xlim = W - 2, ylim = H - 2;
    uint8_t* mp = (dir ? a.mask[1] : a.mask[0]) + b * HW + i;
    unsigned edges = 0;
#pragma unroll
    for (int jp = 0; jp < 4; jp += 2) {
        const F2 fnx = dW(pk(fxs[jp], fxs[jp + 1])), fny = dH(pk(fys[jp], fys[jp + 1]));          // :264
        const F2 xn = sub2(dW(pk((float)(X0 + 32 * jp), (float)(X0 + 32 * jp + 32))), one2);      // :271
        const F2 c1x = add2(xn, fnx), c1y = add2(yn2, fny);                                       // :275
        const F2 ix = mul2(add2(c1x, one2), hw2), iy = mul2(add2(c1y, one2), hh2);
        float c1xs[2], c1ys[2], ixs[2], iys[2], wxs[2], wys[2];
        unpk(c1x, c1xs[0], c1xs[1]); unpk(c1y, c1ys[0], c1ys[1]);
        unpk(ix, ixs[0], ixs[1]); unpk(iy, iys[0], iys[1]);
        bool inb[2];
        float t[2][8];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            inb[p] = (fabsf(c1xs[p]) < 1.0f) && (fabsf(c1ys[p]) < 1.0f);                          // :276
            const int x0 = __float2int_rd(ixs[p]), y0 = __float2int_rd(iys[p]);
            wxs[p] = sub(ixs[p], __int2float_rn(x0));
            wys[p] = sub(iys[p], __int2float_rn(y0));
            if (inb[p] && ((unsigned)x0 > xlim || (unsigned)y0 > ylim)) edges |= 1u << (jp + p);
            // clamp: garbage coordinates of out-of-frame pixels still address valid memory
            const unsigned xc = min((unsigned)x0, xlim), yc = min((unsigned)y0, ylim);
            const float* p0 = ptr_at(g, (int)(yc * W + xc));  // row y0, x channel
            if (WC) {  // immediate offsets off one pointer
                t[p][0] = __ldg(p0); t[p][1] = __ldg(p0 + 1); t[p][2] = __ldg(p0 + WC); t[p][3] = __ldg(p0 + WC + 1);
                t[p][4] = __ldg(p0 + WC * HC); t[p][5] = __ldg(p0 + WC * HC + 1);
                t[p][6] = __ldg(p0 + WC * HC + WC); t[p][7] = __ldg(p0 + WC * HC + WC + 1);
            } else {
                const float* p1 = ptr_at(p0, W);   // row y0+1
                const float* q0 = ptr_at(p0, HW);  // y channel
                const float* q1 = ptr_at(q0, W);
                t[p][0] = __ldg(p0); t[p][1] = __ldg(p0 + 1); t[p][2] = __ldg(p1); t[p][3] = __ldg(p1 + 1);
                t[p][4] = __ldg(q0); t[p][5] = __ldg(q0 + 1); t[p][6] = __ldg(q1); t[p][7] = __ldg(q1 + 1);
            }
        }
        const F2 wx = pk(wxs[0], wxs[1]), wy = pk(wys[0], wys[1]);
        const F2 e = sub2(one2, wx), s_ = sub2(one2, wy);
        const F2 nw = mul2(s_, e), ne = mul2(s_, wx), sw = mul2(wy, e), se = mul2(wy, wx);
        const F2 bx = combine4_2(dW(pk(t[0][0], t[1][0])), dW(pk(t[0][1], t[1][1])), dW(pk(t[0][2], t[1][2])),
                                 dW(pk(t[0][3], t[1][3])), nw, ne, sw, se);
        const F2 by = combine4_2(dH(pk(t[0][4], t[1][4])), dH(pk(t[0][5], t[1][5])), dH(pk(t[0][6], t[1][6])),
                                 dH(pk(t[0][7], t[1][7])), nw, ne, sw, se);
        const F2 cyx = add2(fnx, bx), cyy = add2(fny, by);                                        // :279
        float cx2[2], cy2[2], fx2[2], fy2[2], bx2[2], by2[2];
        unpk(mul2(cyx, cyx), cx2[0], cx2[1]); unpk(mul2(cyy, cyy), cy2[0], cy2[1]);
        unpk(mul2(fnx, fnx), fx2[0], fx2[1]); unpk(mul2(fny, fny), fy2[0], fy2[1]);
        unpk(mul2(bx, bx), bx2[0], bx2[1]); unpk(mul2(by, by), by2[0], by2[1]);
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const float cyc2 = add(cx2[p], cy2[p]);                                               // :293
            const float eps = add(mul(a.a1, add(add(fx2[p], fy2[p]), add(bx2[p], by2[p]))), a.a2);  // :294
            mp[32 * (jp + p)] = (inb[p] && (sub(cyc2, eps) <= 0.0f)) ? 1 : 0;                     // :296
        }
    }
    if (edges) {  // pixels warped exactly onto the last row / column: redo with predicated taps
#pragma unroll 1
        for (int j = 0; j < 4; j++)
            if (edges >> j & 1)
                fb_pixel_redo<DM>(f, g, (dir ? a.mask[1] : a.mask[0]) + b * HW, X0 + 32 * j, Y, W, H, a.half_w, a.half_h, a.a1, a.a2, a.dw, a.dh);
    }
}

}  // namespace pp
#include "pp_fbtile.cuh"
#include "pp_chainup.cuh"
namespace pp {

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

template <int DM>
static int launch_chain_dm(const float* l0, const float* l1, float* o0, float* o1, int ndir, int n, int64_t B, int H, int W,
                           int h, int w, bool up, int64_t stride_n, int64_t stride_b, int is_norm, cudaStream_t st) {
    ChainArgs<DM> a;
    a.links[0] = l0; a.links[1] = l1; a.out[0] = o0; a.out[1] = o1;
    a.stride_n = stride_n; a.stride_b = stride_b;
    a.n = n; a.H = H; a.W = W; a.h = h; a.w = w;
    a.rh = up ? up_scale(h, H) : 0.f; a.rw = up ? up_scale(w, W) : 0.f;
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.dw = make_div<DM>((float)(W - 1)); a.dh = make_div<DM>((float)(H - 1));
    a.ndir = ndir;
    dim3 block(32, 8);
    if (up && n == 1 && !is_norm) {  // W = 8w is a multiple of 4; H = 8h a multiple of 8
        dim3 grid((W / 4 + 31) / 32, (H / 8 + 7) / 8, (unsigned)(B * ndir));
        PP_LAUNCH(ndir == 1 ? "chain_up1" : "chain_up", st, upchain1_kernel<DM><<<grid, block, 0, st>>>(a));  // chain_up1: one direction
        return check_launch("upchain1_kernel");
    }
    if constexpr (DM == DM_FAST) {
        if (up && n > 1 && !is_norm) {  // x8 up-sampling fused into the chain, no scratch (pp_chainup.cuh); -1 = not applicable
            const int rc = cup::launch(l0, l1, o0, o1, ndir, n, B, h, w, stride_n, stride_b, st);
            if (rc != -1) return rc;
        }
        if (!up && n > 1 && !is_norm) {  // TMA-staged dense chain (pp_fbtile.cuh); -1 = not applicable
            const int rc = fbt::launch_chain_box(l0, l1, o0, o1, ndir, n, B, H, W, stride_n, stride_b, st);
            if (rc != -1) return rc;
        }
    }
    if (!up && n > 1 && !is_norm && (W % 128 == 0) && H >= 2) {
        dim3 grid4(W / 128, (H + 7) / 8, (unsigned)(B * ndir));
        if (W == 1280 && H == 720) PP_LAUNCH("chain_dense", st, (chain_dense4_kernel<DM, 1280, 720><<<grid4, block, 0, st>>>(a)));
        else PP_LAUNCH("chain_dense", st, (chain_dense4_kernel<DM, 0, 0><<<grid4, block, 0, st>>>(a)));
        return check_launch("chain_dense4_kernel");
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, (unsigned)(B * ndir));
    if (up) {
        if (is_norm) PP_LAUNCH("chain_up_norm", st, (chain_kernel<true, true, DM><<<grid, block, 0, st>>>(a)));
        else PP_LAUNCH("chain_up", st, (chain_kernel<true, false, DM><<<grid, block, 0, st>>>(a)));
    } else {
        if (is_norm) PP_LAUNCH("chain_dense_norm", st, (chain_kernel<false, true, DM><<<grid, block, 0, st>>>(a)));
        else PP_LAUNCH("chain_dense", st, (chain_kernel<false, false, DM><<<grid, block, 0, st>>>(a)));
    }
    return check_launch("chain_kernel");
}

static int launch_chain(const float* l0, const float* l1, float* o0, float* o1, int ndir, int n, int64_t B, int H, int W,
                        int h, int w, bool up, int64_t stride_n, int64_t stride_b, int is_norm, int div_mode,
                        cudaStream_t st) {
    // The chain's only divisions normalise running coordinates (is_norm: also tap values; that
    // rarely used mode keeps the IEEE sequence).  DM_FAST is exact there, see pp_common.cuh.
    if (div_mode == PP_DIV_RCP)
        return launch_chain_dm<DM_RCP>(l0, l1, o0, o1, ndir, n, B, H, W, h, w, up, stride_n, stride_b, is_norm, st);
    if (!is_norm && div_certified((float)(W - 1)) && div_certified((float)(H - 1)))
        return launch_chain_dm<DM_FAST>(l0, l1, o0, o1, ndir, n, B, H, W, h, w, up, stride_n, stride_b, is_norm, st);
    return launch_chain_dm<DM_IEEE>(l0, l1, o0, o1, ndir, n, B, H, W, h, w, up, stride_n, stride_b, is_norm, st);
}

template <int DM>
static int launch_fb_dm(const float* f0, const float* f1, uint8_t* m0, uint8_t* m1, float* cycle, float* coords1, int ndir,
                        int64_t B, int H, int W, double alpha_1, double alpha_2, int is_norm, bool mask_only4,
                        cudaStream_t st) {
    FbArgs<DM> a;
    a.flow[0] = f0; a.flow[1] = f1; a.mask[0] = m0; a.mask[1] = m1;
    a.cycle = cycle; a.coords1 = coords1; a.H = H; a.W = W;
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.a1 = (float)alpha_1; a.a2 = fb_alpha2_eff(alpha_2, H, W);
    a.dw = make_div<DM>((float)(W - 1)); a.dh = make_div<DM>((float)(H - 1));
    a.dw2 = make_div<DM>((float)(W - 1) / 2.0f); a.dh2 = make_div<DM>((float)(H - 1) / 2.0f);
    a.ndir = ndir;
    dim3 block(32, 8);
    if (mask_only4) {
        dim3 grid(W / 128, (H + 7) / 8, (unsigned)(B * ndir));
        if constexpr (DM == DM_FAST) {
            if (W == 1280 && H == 720) PP_LAUNCH("fb", st, (fbmask4p_kernel<1280, 720><<<grid, block, 0, st>>>(a)));
            else PP_LAUNCH("fb", st, (fbmask4p_kernel<0, 0><<<grid, block, 0, st>>>(a)));
            return check_launch("fbmask4p_kernel");
        }
        if (W == 1280 && H == 720) PP_LAUNCH("fb", st, (fbmask4_kernel<DM, 1280, 720><<<grid, block, 0, st>>>(a)));
        else PP_LAUNCH("fb", st, (fbmask4_kernel<DM, 0, 0><<<grid, block, 0, st>>>(a)));
        return check_launch("fbmask4_kernel");
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, (unsigned)(B * ndir));
    if (is_norm) PP_LAUNCH("fb_norm", st, (fb_kernel<true, DM><<<grid, block, 0, st>>>(a)));
    else PP_LAUNCH("fb", st, (fb_kernel<false, DM><<<grid, block, 0, st>>>(a)));
    return check_launch("fb_kernel");
}

static int launch_fb(const float* f0, const float* f1, uint8_t* m0, uint8_t* m1, float* cycle, float* coords1, int ndir,
                     int64_t B, int H, int W, double alpha_1, double alpha_2, int is_norm, int div_mode, cudaStream_t st) {
    const bool mask_only4 = !cycle && !coords1 && !is_norm && (W % 128 == 0) && H >= 2;
    if (div_mode == PP_DIV_RCP)
        return launch_fb_dm<DM_RCP>(f0, f1, m0, m1, cycle, coords1, ndir, B, H, W, alpha_1, alpha_2, is_norm, mask_only4, st);
    const bool cert = div_certified((float)(W - 1)) && div_certified((float)(H - 1));
    const bool cert_half = div_certified((float)(W - 1) / 2.0f) && div_certified((float)(H - 1) / 2.0f);
    if (cert && cert_half && mask_only4 && fb_alpha2_eff(alpha_2, H, W) >= 1e-12f) {
        // TMA-staged tile kernel (pp_fbtile.cuh); -1 = not applicable to this shape / alignment
        const int rc = fbt::launch(f0, f1, m0, m1, ndir, B, H, W, (float)alpha_1, fb_alpha2_eff(alpha_2, H, W), st);
        if (rc != -1) return rc;
    }
    if (cert && cert_half && mask_only4)
        return launch_fb_dm<DM_FAST>(f0, f1, m0, m1, cycle, coords1, ndir, B, H, W, alpha_1, alpha_2, is_norm, true, st);
    if (cert)  // values leave the kernel (cycle / coords1): guarded variant, exact for every input
        return launch_fb_dm<DM_FASTG>(f0, f1, m0, m1, cycle, coords1, ndir, B, H, W, alpha_1, alpha_2, is_norm, false, st);
    return launch_fb_dm<DM_IEEE>(f0, f1, m0, m1, cycle, coords1, ndir, B, H, W, alpha_1, alpha_2, is_norm, mask_only4, st);
}


// ------------------------------------------------------------------------------------------
// Sparse correspondence: the flow stage evaluated only where the loss looks at it.
//
// regression_loss consumes the dense composite flow and FB mask of apply_optical_flow at the P = G*G
// grid centres of each sample only (add_optical_flow, PixPro.py:46-89: one bilinear grid_sample of the
// flow, one nearest lookup of the mask).  Every op of the flow stage is point-wise, so the composite at
// the 4 integer neighbours of a centre, and the FB test at its nearest pixel (which needs the opposite
// composite at the 4 neighbours of that pixel's warped position), evaluated on the fly with the same
// arithmetic, are bit-identical to sampling the dense tensors — at ~10^-3 of the work and no HBM
// traffic beyond the low-res links.  8 lanes per grid point: lanes 0-3 evaluate the composite at the 4
// taps, lane 4 at the nearest pixel (phase 1, concurrently); lanes 0-3 then evaluate the opposite
// composite at the 4 taps of the FB sample (phase 2); lane 0 combines.
// out[dir] = [3,B,P]: warped centre x, y (PixPro.py:76-83) and the mask bit (PixPro.py:65-70) as 0/1.
// ------------------------------------------------------------------------------------------
template <int DM>
struct SparseArgs {
    ChainArgs<DM> ch;        // links[0] = forward links, links[1] = backward links; out unused
    const float* coord[2];   // crop descriptors [B,10] whose centres are warped by direction d (NULL: skip)
    float* out[2];           // [3,B,P] per direction
    int64_t B;
    int G, P, use_mask;
    float wo, ho;            // W_orig-1, H_orig-1
    ScalarDiv dG;
    WarpArgs warp;
    float a1, a2;
};

template <bool UP, int DM>
__global__ void __launch_bounds__(256) sparse_corr_kernel(SparseArgs<DM> a) {
    const int64_t gid = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3;
    const int sl = threadIdx.x & 7, lane = threadIdx.x & 31, base = lane & ~7;
    const unsigned gmask = 0xffu << base;
    const int64_t BP = a.B * a.P;
    if (gid >= 2 * BP) return;  // whole 8-lane groups leave together
    const int dir = gid >= BP ? 1 : 0;
    if (!a.coord[dir]) return;
    const int64_t r = gid - (int64_t)dir * BP;
    const int64_t b = r / a.P;
    const int p = (int)(r - b * a.P);
    const ChainArgs<DM>& ch = a.ch;
    const int W = ch.W, H = ch.H;
    const float* lf = (dir ? ch.links[1] : ch.links[0]) + b * ch.stride_b;  // this direction's links
    const float* lg = (dir ? ch.links[0] : ch.links[1]) + b * ch.stride_b;  // the opposite direction's
    float vqx, vqy;
    grid_centre(a.coord[dir] + b * 10, p % a.G, p / a.G, a.dG, a.wo, a.ho, vqx, vqy);
    // PixPro.py:61-62   2 * (x / (W_orig-1)) - 1
    const float gx = sub(mul(2.0f, a.warp.dwo(vqx)), 1.0f), gy = sub(mul(2.0f, a.warp.dho(vqy)), 1.0f);
    const Taps t = make_taps(gx, gy, W, H, a.warp.half_w, a.warp.half_h);
    // ---- phase 1: composite at the 4 taps (lanes 0-3) and at the nearest pixel (lane 4) ----
    int X = 0, Y = 0;
    bool act = false;
    if (sl < 4) {
        X = t.x0 + (sl & 1); Y = t.y0 + (sl >> 1);
        act = ((sl & 1) ? t.inx1 : t.inx0) && ((sl >> 1) ? t.iny1 : t.iny0);
    } else if (sl == 4 && a.use_mask) {  // PixPro.py:65-70 nearest lookup (nearbyint, zeros padding)
        const float ix = mul(add(gx, 1.0f), a.warp.half_w), iy = mul(add(gy, 1.0f), a.warp.half_h);
        const float xr = rintf(ix), yr = rintf(iy);
        act = (xr > -1.0f) && (xr < (float)W) && (yr > -1.0f) && (yr < (float)H);
        X = act ? (int)xr : 0; Y = act ? (int)yr : 0;
    }
    float2 v = make_float2(0.0f, 0.0f);
    if (act) v = chain_pixel<UP, false, DM>(ch, lf, X, Y);
    float tx[4], ty[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        tx[c] = __shfl_sync(gmask, v.x, base + c);
        ty[c] = __shfl_sync(gmask, v.y, base + c);
    }
    const float fgx = combine(t, tx[0], tx[1], tx[2], tx[3]);  // PixPro.py:64
    const float fgy = combine(t, ty[0], ty[1], ty[2], ty[3]);
    // ---- phase 2: forward-backward test at the nearest pixel (util.py:253-297, fb_kernel<false>) ----
    bool mg = true;
    if (a.use_mask) {
        const float fx = __shfl_sync(gmask, v.x, base + 4), fy = __shfl_sync(gmask, v.y, base + 4);
        const int Xn = __shfl_sync(gmask, X, base + 4), Yn = __shfl_sync(gmask, Y, base + 4);
        const bool inb = __shfl_sync(gmask, (int)act, base + 4) != 0;
        mg = false;
        if (inb) {  // uniform over the 8-lane group
            const float fnx = norm_flow(fx, ch.dw), fny = norm_flow(fy, ch.dh);                    // :264
            const float c1x = add(norm_coord((float)Xn, ch.dw), fnx), c1y = add(norm_coord((float)Yn, ch.dh), fny);  // :271,275
            const bool in1 = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);                           // :276
            const Taps t2 = make_taps(c1x, c1y, W, H, ch.half_w, ch.half_h);
            float2 gv = make_float2(0.0f, 0.0f);
            if (sl < 4) {
                const bool act2 = ((sl & 1) ? t2.inx1 : t2.inx0) && ((sl >> 1) ? t2.iny1 : t2.iny0);
                if (act2) gv = chain_pixel<UP, false, DM>(ch, lg, t2.x0 + (sl & 1), t2.y0 + (sl >> 1));
                gv.x = norm_flow(gv.x, ch.dw);  // :265 the sampled field is the normalised one
                gv.y = norm_flow(gv.y, ch.dh);
            }
            float ux[4], uy[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                ux[c] = __shfl_sync(gmask, gv.x, base + c);
                uy[c] = __shfl_sync(gmask, gv.y, base + c);
            }
            const float bix = combine(t2, ux[0], ux[1], ux[2], ux[3]), biy = combine(t2, uy[0], uy[1], uy[2], uy[3]);  // :278
            const float cyx = add(fnx, bix), cyy = add(fny, biy);                                  // :279
            const float cyc2 = add(mul(cyx, cyx), mul(cyy, cyy));                                  // :293
            const float f2 = add(mul(fnx, fnx), mul(fny, fny)), b2 = add(mul(bix, bix), mul(biy, biy));
            const float eps = add(mul(a.a1, add(f2, b2)), a.a2);                                   // :294
            mg = in1 && (sub(cyc2, eps) <= 0.0f);                                                  // :296
        }
    }
    if (sl == 0) {
        float ox, oy;
        if (a.warp.diff) {  // PixPro.py:76-80
            ox = a.warp.drw(add(mul(vqx, a.warp.rw), fgx));
            oy = a.warp.drh(add(mul(vqy, a.warp.rh), fgy));
        } else {  // PixPro.py:82-83
            ox = add(vqx, fgx);
            oy = add(vqy, fgy);
        }
        float* o = a.out[dir] + b * a.P + p;
        o[0] = ox;
        o[BP] = oy;
        o[2 * BP] = mg ? 1.0f : 0.0f;
    }
}

template <int DM>
static int launch_sparse_dm(const float* lo_fwd, const float* lo_bwd, int64_t B, int n, int h, int w, bool up, bool use_mask,
                            double alpha_1, double alpha_2, const float* coord_fwd, const float* coord_bwd, int G, int H_orig,
                            int W_orig, int div_mode, float* warped_fwd, float* warped_bwd, cudaStream_t st) {
    const int H = up ? 8 * h : h, W = up ? 8 * w : w;
    SparseArgs<DM> a;
    ChainArgs<DM>& c = a.ch;
    c.links[0] = lo_fwd; c.links[1] = lo_bwd; c.out[0] = c.out[1] = nullptr;
    c.stride_n = 2 * (int64_t)h * w; c.stride_b = (int64_t)n * c.stride_n;  // loader layout [B,n,2,h,w]
    c.n = n; c.H = H; c.W = W; c.h = h; c.w = w;
    c.rh = up ? up_scale(h, H) : 0.f; c.rw = up ? up_scale(w, W) : 0.f;
    c.half_w = (float)(W - 1) / 2.0f; c.half_h = (float)(H - 1) / 2.0f;
    c.dw = make_div<DM>((float)(W - 1)); c.dh = make_div<DM>((float)(H - 1));
    c.ndir = 2;
    a.coord[0] = coord_fwd; a.coord[1] = coord_bwd; a.out[0] = warped_fwd; a.out[1] = warped_bwd;
    a.B = B; a.G = G; a.P = G * G; a.use_mask = use_mask ? 1 : 0;
    a.wo = (float)(W_orig - 1); a.ho = (float)(H_orig - 1);
    a.dG = make_div((float)G, div_mode);
    a.warp = make_warp_args(H, W, H_orig, W_orig, div_mode);
    a.a1 = (float)alpha_1; a.a2 = use_mask ? fb_alpha2_eff(alpha_2, H, W) : 0.0f;
    const int64_t threads = 2 * B * a.P * 8;
    const unsigned nb = (unsigned)((threads + 255) / 256);
    if (up) PP_LAUNCH("sparse_corr", st, (sparse_corr_kernel<true, DM><<<nb, 256, 0, st>>>(a)));
    else PP_LAUNCH("sparse_corr", st, (sparse_corr_kernel<false, DM><<<nb, 256, 0, st>>>(a)));
    return check_launch("sparse_corr_kernel");
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_upflow8(const float* in, int64_t planes, int h, int w, float* out, void* stream) {
    PP_REQUIRE(planes >= 0 && h > 0 && w > 0, "pp_upflow8: bad shape planes=%lld h=%d w=%d", (long long)planes, h, w);
    if (planes == 0) return PP_OK;  // empty batch: nothing to do (pointers may be null)
    PP_REQUIRE(in && out, "pp_upflow8: null pointer");
    PP_REQUIRE((int64_t)h * w * 64 < (1ll << 31), "pp_upflow8: plane too large");
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
    unsigned nb = (unsigned)((total + 255) / 256);
    const float sw = (float)(W - 1), sh = (float)(H - 1);
    if (div_mode == PP_DIV_RCP)
        PP_LAUNCH("normalize", st, normalize_kernel<DM_RCP><<<nb, 256, 0, st>>>(x, total, HW, kind, make_div<DM_RCP>(sw), make_div<DM_RCP>(sh), out));
    else if (div_certified(sw) && div_certified(sh))
        PP_LAUNCH("normalize", st, normalize_kernel<DM_FASTG><<<nb, 256, 0, st>>>(x, total, HW, kind, make_div<DM_FASTG>(sw), make_div<DM_FASTG>(sh), out));
    else
        PP_LAUNCH("normalize", st, normalize_kernel<DM_IEEE><<<nb, 256, 0, st>>>(x, total, HW, kind, make_div<DM_IEEE>(sw), make_div<DM_IEEE>(sh), out));
    return check_launch("normalize_kernel");
}

int pp_concat_flow(const float* flows, int n, int64_t B, int H, int W, int64_t stride_n, int64_t stride_b, int is_norm,
                   int div_mode, float* out, void* stream) {
    PP_REQUIRE(n >= 1 && B >= 0 && H > 1 && W > 1, "pp_concat_flow: bad shape n=%d B=%lld H=%d W=%d", n, (long long)B, H, W);
    PP_REQUIRE(B <= 65535, "pp_concat_flow: B=%lld exceeds 65535", (long long)B);
    PP_REQUIRE((int64_t)H * W * 2 < (1ll << 31), "pp_concat_flow: frame too large");
    if (B == 0) return PP_OK;
    PP_REQUIRE(flows && out, "pp_concat_flow: null pointer");
    return launch_chain(flows, nullptr, out, nullptr, 1, n, B, H, W, 0, 0, false, stride_n, stride_b, is_norm, div_mode,
                        (cudaStream_t)stream);
}

int pp_fb_consistency(const float* fwd, const float* bwd, int64_t B, int H, int W, double alpha_1, double alpha_2,
                      int is_norm, int div_mode, uint8_t* mask, float* cycle, float* coords1_norm, void* stream) {
    PP_REQUIRE(B >= 0 && H > 1 && W > 1, "pp_fb_consistency: bad shape B=%lld H=%d W=%d", (long long)B, H, W);
    PP_REQUIRE(B <= 65535, "pp_fb_consistency: B=%lld exceeds 65535", (long long)B);
    PP_REQUIRE((int64_t)H * W * 2 < (1ll << 31), "pp_fb_consistency: frame too large");
    if (B == 0) return PP_OK;
    PP_REQUIRE(fwd && bwd && mask, "pp_fb_consistency: null pointer");
    return launch_fb(fwd, bwd, mask, nullptr, cycle, coords1_norm, 1, B, H, W, alpha_1, alpha_2, is_norm, div_mode,
                     (cudaStream_t)stream);
}

// Both FB masks of apply_optical_flow (contrast/util.py:211-213: the two forward_backward_consistency calls, arguments
// swapped) in ONE launch — the second half of pp_flow_stage as its own entry, so a caller can run it on another stream
// than the chain (sample chunks: the masks of chunk i under the up-sampling of chunk i+1; host_step.py).
int pp_fb_masks(const float* fwd, const float* bwd, int64_t B, int H, int W, double alpha_1, double alpha_2, int is_norm,
                int div_mode, uint8_t* mask_fwd, uint8_t* mask_bwd, void* stream) {
    PP_REQUIRE(B >= 0 && H > 1 && W > 1, "pp_fb_masks: bad shape B=%lld H=%d W=%d", (long long)B, H, W);
    PP_REQUIRE(B * 2 <= 65535, "pp_fb_masks: B=%lld exceeds 32767", (long long)B);
    PP_REQUIRE((int64_t)H * W * 2 < (1ll << 31), "pp_fb_masks: frame too large");
    if (B == 0) return PP_OK;
    PP_REQUIRE(fwd && bwd && mask_fwd && mask_bwd, "pp_fb_masks: null pointer");
    return launch_fb(fwd, bwd, mask_fwd, mask_bwd, nullptr, nullptr, 2, B, H, W, alpha_1, alpha_2, is_norm, div_mode,
                     (cudaStream_t)stream);
}

// The n > 1 flow_up path needs no scratch since round 2 (the x8 up-sampling is fused into the chain kernel,
// pp_chainup.cuh); the entry stays in the ABI and returns 0.
int64_t pp_flow_stage_workspace(int64_t B, int n, int h, int w, int flow_up) {
    (void)B; (void)n; (void)h; (void)w; (void)flow_up;
    return 0;
}

int pp_flow_stage(const float* lo_fwd, const float* lo_bwd, int64_t B, int n, int h, int w, int flow_up, int use_mask,
                  double alpha_1, double alpha_2, int is_norm, int div_mode, float* flow_fwd, float* flow_bwd,
                  uint8_t* mask_fwd, uint8_t* mask_bwd, void* workspace, int64_t workspace_bytes, void* stream) {
    PP_REQUIRE(n >= 1 && B >= 0 && h > 1 && w > 1, "pp_flow_stage: bad shape B=%lld n=%d h=%d w=%d", (long long)B, n, h, w);
    PP_REQUIRE(B * 2 <= 65535, "pp_flow_stage: B=%lld exceeds 32767", (long long)B);
    PP_REQUIRE((int64_t)h * w * 128 < (1ll << 31), "pp_flow_stage: frame too large");
    if (B == 0) return PP_OK;
    PP_REQUIRE(lo_fwd && lo_bwd && flow_fwd && flow_bwd, "pp_flow_stage: null pointer");
    PP_REQUIRE(!use_mask || (mask_fwd && mask_bwd), "pp_flow_stage: use_mask set but mask outputs are null");
    cudaStream_t st = (cudaStream_t)stream;
    int H = flow_up ? 8 * h : h, W = flow_up ? 8 * w : w;
    int64_t link = 2 * (int64_t)h * w;  // loader layout [B,n,2,h,w]
    int rc;
    (void)workspace; (void)workspace_bytes;
    // n == 1 with up-sampling and masks (the published n_frames = 2 setting): the backward composite is produced inside its
    // mask kernel and the forward mask kernel recomputes its own composite (fbt::fbbox_up_kernel, pp_fbtile.cuh) — same bits,
    // one HBM-write-bound half of the up-sampling launch gone.  Conditions = those of the TMA-staged FB kernel (launch_fb).
    if (n == 1 && flow_up && use_mask && !is_norm && div_mode != PP_DIV_RCP && div_certified((float)(W - 1)) &&
        div_certified((float)(H - 1)) && div_certified((float)(W - 1) / 2.0f) && div_certified((float)(H - 1) / 2.0f) &&
        fb_alpha2_eff(alpha_2, H, W) >= 1e-12f && fbt::up_applicable(flow_fwd, flow_bwd, B, H, W, h, w, (float)alpha_1)) {
        rc = launch_chain(lo_fwd, nullptr, flow_fwd, nullptr, 1, 1, B, H, W, h, w, true, link, link, 0, div_mode, st);
        if (rc) return rc;
        const float a2e = fb_alpha2_eff(alpha_2, H, W);
        rc = fbt::launch_up(lo_bwd, link, flow_bwd, flow_fwd, mask_bwd, B, H, W, h, w, (float)alpha_1, a2e, st);
        if (rc == 0 && fbt::up_mode(B) == 2)  // the forward mask on the plain kernel (its own composite loaded)
            return launch_fb(flow_fwd, flow_bwd, mask_fwd, nullptr, nullptr, nullptr, 1, B, H, W, alpha_1, alpha_2, 0, div_mode, st);
        if (rc == 0) rc = fbt::launch_up(lo_fwd, link, nullptr, flow_bwd, mask_fwd, B, H, W, h, w, (float)alpha_1, a2e, st);
        if (rc >= 0) return rc;
        // a tensor map could not be encoded: the forward composite exists, finish on the two-kernel route below
        rc = launch_chain(lo_bwd, nullptr, flow_bwd, nullptr, 1, 1, B, H, W, h, w, true, link, link, 0, div_mode, st);
        if (rc) return rc;
        return launch_fb(flow_fwd, flow_bwd, mask_fwd, mask_bwd, nullptr, nullptr, 2, B, H, W, alpha_1, alpha_2, 0, div_mode, st);
    }
    rc = launch_chain(lo_fwd, lo_bwd, flow_fwd, flow_bwd, 2, n, B, H, W, h, w, flow_up != 0, link, (int64_t)n * link, is_norm,
                      div_mode, st);
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

int pp_sparse_corr(const float* lo_fwd, const float* lo_bwd, int64_t B, int n, int h, int w, int flow_up, int use_mask,
                   double alpha_1, double alpha_2, const float* coord_fwd, const float* coord_bwd, int G, int H_orig, int W_orig,
                   int div_mode, float* warped_fwd, float* warped_bwd, void* stream) {
    PP_REQUIRE(n >= 1 && B >= 0 && h > 1 && w > 1, "pp_sparse_corr: bad shape B=%lld n=%d h=%d w=%d", (long long)B, n, h, w);
    PP_REQUIRE(G > 0 && G * G <= 1024 && H_orig > 1 && W_orig > 1, "pp_sparse_corr: bad grid %d or original size %dx%d", G, H_orig, W_orig);
    PP_REQUIRE((int64_t)h * w * 128 < (1ll << 31) && B * (int64_t)G * G < (1ll << 26), "pp_sparse_corr: problem too large");
    if (B == 0) return PP_OK;
    PP_REQUIRE(lo_fwd && lo_bwd, "pp_sparse_corr: null links");
    PP_REQUIRE((coord_fwd && warped_fwd) || (coord_bwd && warped_bwd), "pp_sparse_corr: no direction requested");
    PP_REQUIRE((!coord_fwd) == (!warped_fwd) && (!coord_bwd) == (!warped_bwd), "pp_sparse_corr: coord / warped pointers must come in pairs");
    // few points, latency-bound: plain IEEE division (what the certified fast forms of the dense kernels equal)
    if (div_mode == PP_DIV_RCP)
        return launch_sparse_dm<DM_RCP>(lo_fwd, lo_bwd, B, n, h, w, flow_up != 0, use_mask != 0, alpha_1, alpha_2, coord_fwd, coord_bwd,
                                        G, H_orig, W_orig, div_mode, warped_fwd, warped_bwd, (cudaStream_t)stream);
    return launch_sparse_dm<DM_IEEE>(lo_fwd, lo_bwd, B, n, h, w, flow_up != 0, use_mask != 0, alpha_1, alpha_2, coord_fwd, coord_bwd,
                                     G, H_orig, W_orig, div_mode, warped_fwd, warped_bwd, (cudaStream_t)stream);
}

int64_t pp_fb_redo_count(int reset) {
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, fbt::g_redo_pixels, sizeof(v)) != cudaSuccess) return -1;
    unsigned int to = 0;
    if (cudaMemcpyFromSymbol(&to, fbt::g_wait_timeouts, sizeof(to)) != cudaSuccess) return -1;
    if (to) {
        set_error("pp_fb_redo_count: %u mbarrier waits of the tile kernel timed out (results invalid)", to);
        return -2;
    }
    if (reset) {
        const unsigned long long z = 0;
        cudaMemcpyToSymbol(fbt::g_redo_pixels, &z, sizeof(z));
    }
    return (int64_t)v;
}

int64_t pp_chain_slow_count(int reset) {
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, cup::g_slow_pixels, sizeof(v)) != cudaSuccess) return -1;
    if (reset) {
        const unsigned long long z = 0;
        cudaMemcpyToSymbol(cup::g_slow_pixels, &z, sizeof(z));
    }
    return (int64_t)v;
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
